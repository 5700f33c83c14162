import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import gpt2_decoder
dev = torch.device("cuda:0"); torch.set_grad_enabled(False)
m, _ = gpt2_decoder(H=768, layers=12, heads=12, V=50257, max_length=64); m.precision = os.environ.get("PREC", "bf16"); m = m.to(dev)
ef = {"pooled_features": torch.randn(1024, 768, device=dev)}
for _ in range(2):
    m.generate(ef, 20, num_beams=5)
torch.cuda.synchronize()
