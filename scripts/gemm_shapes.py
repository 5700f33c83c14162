"""Time the decode-step GEMM shapes in isolation through a legacy engine's forward (stage timers), with and without
stream-K (CAPDEC_NO_STREAMK=1), at the per-GPU batch sizes of the strong-scaling run."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import legacy_weights
dev = torch.device("cuda:0")
torch.set_grad_enabled(False)
m, _ = legacy_weights(10000, 0); m.precision = "bf16x3"; m = m.to(dev)
eng = m._engine(dev)
for B in (512, 1024, 2048):
    enc = torch.randn(B, 196, 2048, device=dev).relu_()
    for _ in range(3):
        m.beam_search(enc, beam_size=5, max_length=20, crop=False)
    eng.stage_timing(True)
    for _ in range(5):
        m.beam_search(enc, beam_size=5, max_length=20, crop=False)
    st = eng.stage_times(); eng.stage_timing(False)
    print(json.dumps({"images": B, "no_streamk": bool(os.environ.get("CAPDEC_NO_STREAMK")),
                      "gate_us": round(st["gate_gemm"][0] / st["gate_gemm"][1] * 1e3, 1), "vocab_us": round(st["vocab_gemm"][0] / st["vocab_gemm"][1] * 1e3, 1),
                      "attn_us": round(st["attention"][0] / st["attention"][1] * 1e3, 1)}))
