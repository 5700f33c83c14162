"""Diagnostic: per-beam-slot accuracy of the legacy step (teacher-forced rows, rows_per_image = k) vs the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import legacy as olegacy, teacher as oteach
from tests.helpers import legacy_features, legacy_weights

torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
B, T, V = 96, 12, 10000
m, sd = legacy_weights(V, 0)
enc = legacy_features(B, seed=4242)
for k in (3, 5, 6):
    caps = torch.randint(3, V, (B * k, T), generator=torch.Generator().manual_seed(k))
    caps[:, 0] = 1
    st = olegacy.LegacyStepper(sd, enc, k)
    ref = torch.stack([st(caps[:, t]) for t in range(T)], dim=1)
    ref_lp = oteach.token_logprobs(ref, caps)
    for prec in ("fp32", "bf16x3"):
        m.precision = prec
        mg = m.to(dev)
        eng = mg._engine(dev)
        logits, lp, _ = eng.forward_tokens(enc.reshape(B, 196, 2048).to(dev), None, None, caps.to(dev), k, want_logits=True, want_logprob=True)
        err = (logits.cpu() - ref).abs().amax(dim=2)          # [R, T]
        per_slot = err.view(B, k, T).amax(dim=(0,))            # [k, T]
        print(f"k={k} {prec}: max |dlogit| per beam slot (rows) x step (cols)")
        for b in range(k):
            print("  slot", b, " ".join(f"{float(x):.1e}" for x in per_slot[b]))
