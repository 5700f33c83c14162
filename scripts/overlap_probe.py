"""Probe: two half-batch decodes on two streams (SM-partitioned persistent kernels) vs one full-batch decode."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import capdec_b200 as cd
from tests.helpers import legacy_weights

dev = torch.device("cuda:0")
B, L, D = 4096, 196, 2048
feats = torch.randn(B, L, D, device=dev).relu_()
prec = os.environ.get("PREC", "bf16x3")
engs = []
for _ in range(2):
    m, _ = legacy_weights(10000, 0)
    m.precision = prec
    engs.append(m.to(dev)._engine(dev))
halves = [feats[: B // 2], feats[B // 2:]]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def full():
    return engs[0].decode_beam(feats, None, None, 5, 20)

def split():
    outs = []
    for e, h, s in zip(engs, halves, streams):
        with torch.cuda.stream(s):
            outs.append(e.decode_beam(h, None, None, 5, 20))
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    return outs

for name, fn in (("full", full), ("split2", split)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name}: {ms:.1f} ms per {B} images = {B / ms * 1e3:.0f} img/s")
