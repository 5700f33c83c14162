"""Decode time of configs[1] at small per-GPU batches (the strong-scaling shards), ms per decode."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import legacy_weights
dev = torch.device("cuda:0")
torch.set_grad_enabled(False)
m, _ = legacy_weights(10000, 0); m.precision = "bf16x3"; m = m.to(dev)
res = {}
for B in (64, 256, 512, 1024, 2048):
    enc = torch.randn(B, 196, 2048, device=dev).relu_()
    for _ in range(3):
        m.beam_search(enc, beam_size=5, max_length=20, crop=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        m.beam_search(enc, beam_size=5, max_length=20, crop=False)
    e1.record(); torch.cuda.synchronize()
    res[B] = round(e0.elapsed_time(e1) / 10, 3)
print(json.dumps({"ms_per_decode": res}))
