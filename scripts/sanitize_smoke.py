"""Tiny decodes of every decoder family for `compute-sanitizer --tool memcheck` (run one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import gpt2_decoder, legacy_weights, lstm_decoder, transformer_decoder

dev = torch.device("cuda:0")
torch.set_grad_enabled(False)
g = torch.Generator().manual_seed(0)
for prec in ("bf16x3", "fp32"):
    m, _ = legacy_weights(600, 0); m.precision = prec; m = m.to(dev)
    enc = torch.relu(torch.randn(5, 196, 2048, generator=g)).to(dev)
    m.beam_search(enc, beam_size=3, max_length=6)
    m.greedy(enc, max_length=5)
    m.sample(enc, num_samples=2, with_greedy=True, max_length=5)
    for kind, heads in (("soft", 8), ("multi_head", 8), ("aoa", 8), ("adaptive", 1)):
        d, _ = lstm_decoder(kind, H=128, layers=2, heads=heads, V=300); d.precision = prec; d = d.to(dev)
        ef = {"features": torch.randn(3, 37, 128, generator=g).to(dev), "pooled_features": torch.randn(3, 128, generator=g).to(dev)}
        d.generate(ef, 5, num_beams=3)
        d.generate(ef, 5)
    t, _ = transformer_decoder(H=128, layers=2, heads=4, V=300); t.precision = prec; t = t.to(dev)
    t.generate({"features": torch.randn(3, 49, 128, generator=g).to(dev)}, 6, num_beams=3)
    p, _ = gpt2_decoder(H=64, layers=2, heads=4, V=300); p.precision = prec; p = p.to(dev)
    p.generate({"pooled_features": torch.randn(3, 64, generator=g).to(dev)}, 6, num_beams=3)
    p.generate({"pooled_features": torch.randn(3, 64, generator=g).to(dev)}, 6, do_sample=True, num_samples=2, with_greedy=True)
torch.cuda.synchronize()
print("sanitize_smoke ok")
