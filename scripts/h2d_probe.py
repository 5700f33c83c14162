"""Probe: pinned host->device copy bandwidth on this box, and e2e decode time vs pipeline chunk size."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import capdec_b200 as cd
from tests.helpers import legacy_weights

dev = torch.device("cuda:0")
B, L, D = int(os.environ.get("B", 4096)), 196, 2048
host = torch.empty(B, L, D, dtype=torch.float32).pin_memory()
host.normal_().relu_()
devbuf = torch.empty_like(host, device=dev)
for n in (256, 512, 4096):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(0, B, n):
        devbuf[i:i + n].copy_(host[i:i + n], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D {host.numel() * 4 / 1e9:.2f} GB in chunks of {n} images: {dt * 1e3:.1f} ms = {host.numel() * 4 / dt / 1e9:.1f} GB/s")
model, _ = legacy_weights(10000, 0)
model.precision = os.environ.get("PREC", "tf32x3")
model = model.to(dev)
eng = model._engine(dev)
out = {"tokens": torch.empty(B, 20, dtype=torch.int32).pin_memory(), "lengths": torch.empty(B, dtype=torch.int32).pin_memory(),
       "scores": torch.empty(B, dtype=torch.float32).pin_memory()}
for chunk in [int(x) for x in os.environ.get("CHUNKS", "296,444,512,592,740").split(",")]:
    for _ in range(2):
        eng.decode_beam_host(host, None, 5, 20, chunk_images=chunk, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.decode_beam_host(host, None, 5, 20, chunk_images=chunk, out=out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"e2e chunk {chunk}: {dt * 1e3:.1f} ms = {B / dt:.0f} img/s")
