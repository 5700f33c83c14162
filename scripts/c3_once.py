import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import transformer_decoder
dev = torch.device("cuda:0"); torch.set_grad_enabled(False)
m, _ = transformer_decoder(H=768, layers=6, heads=8, V=10000, max_length=50); m.precision = os.environ.get("PREC", "bf16x3"); m = m.to(dev)
ef = {"features": torch.randn(2048, 196, 768, device=dev)}
for _ in range(2):
    m.generate(ef, 20, num_beams=3)
torch.cuda.synchronize()
