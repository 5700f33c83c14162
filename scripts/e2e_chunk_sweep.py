"""e2e (host features -> captions) against the H2D pipeline chunk size, per host feature format (one GPU)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from capdec_b200 import engine as eng_mod
from tests.helpers import legacy_weights

dev = torch.device("cuda:0")
B = int(os.environ.get("N", 4096))
m, _ = legacy_weights(10000, 0); m.precision = "bf16x3"; m = m.to(dev)
eng = m._engine(dev)
f32 = bench.global_features(0, B, pin=True).reshape(B, 196, 2048)
srcs = {"f32": (f32, {}), "bf16": (f32.bfloat16().pin_memory(), {}), "p24": (eng_mod.pack_p24_host(f32).pin_memory(), {"dtype": "p24", "num_regions": 196})}
for name, (x, kw) in srcs.items():
    for chunk in (296, 592, 888, 1184, 1480, 2048):
        for _ in range(2):
            eng.decode_beam_host(x, None, 5, 20, chunk_images=chunk, **kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            eng.decode_beam_host(x, None, 5, 20, chunk_images=chunk, **kw)
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 3 * 1e3
        print(json.dumps({"fmt": name, "chunk": chunk, "ms": round(ms, 2), "images_per_s": round(B / ms * 1e3)}))
