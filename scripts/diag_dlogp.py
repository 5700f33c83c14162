"""Diagnostic: locate the largest per-candidate |dlogp| between the GPU beam trace and the CPU oracle at C2 shape and show
who is right (fp64 oracle on that image)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import beam as obeam, legacy as olegacy
from tests.helpers import legacy_features, legacy_weights

torch.set_grad_enabled(False)
B, k, T = int(os.environ.get("N", 1024)), 5, 20
dev = torch.device("cuda:0")
m, sd = legacy_weights(10000, 0)
enc = legacy_features(B, seed=4242)
m.precision = os.environ.get("PREC", "fp32")
out = m.to(dev).beam_search(enc.to(dev), beam_size=k, max_length=T, trace=True)
parts = [obeam.beam_search(olegacy.LegacyStepper(sd, enc[i:i + 256], k), min(256, B - i), k, T, record_steps=True) for i in range(0, B, 256)]
ref_lp = torch.cat([torch.stack([s["top_lp"] for s in p["steps"]]) for p in parts], dim=1)
ref_tok = torch.cat([torch.stack([s["top_tok"] for s in p["steps"]]) for p in parts], dim=1)
ref_beam = torch.cat([torch.stack([s["top_beam"] for s in p["steps"]]) for p in parts], dim=1)
lp, tok, beam = out["top_logprob"].cpu(), out["top_token"].cpu().long(), out["top_beam"].cpu().long()
live = ref_lp > -1e8
agree = ((tok == ref_tok) & (beam == ref_beam)) | ~live
consistent = torch.cumprod(agree.all(dim=2).long(), dim=0).bool()
cmp = live & consistent[:, :, None]
d = (lp - ref_lp).abs() * cmp
flat = d.flatten().topk(12)
for v, idx in zip(flat.values.tolist(), flat.indices.tolist()):
    s, rem = divmod(idx, B * 2 * k)
    i, j = divmod(rem, 2 * k)
    print(f"|dlogp| {v:.3e} step {s} image {i} cand {j} tok {int(tok[s,i,j])} beam {int(beam[s,i,j])} gpu {float(lp[s,i,j]):.6f} ref {float(ref_lp[s,i,j]):.6f}")
s, rem = divmod(int(flat.indices[0]), B * 2 * k)
i, j = divmod(rem, 2 * k)
print("per-step diffs of image", i, "candidate 0..9:")
for t in range(s + 1):
    print(t, [f"{float(x):+.2e}" for x in (lp[t, i] - ref_lp[t, i])], "tok", tok[t, i].tolist(), "rtok", ref_tok[t, i].tolist())
# fp64 oracle on that image alone
sd64 = {n: v.double() for n, v in sd.items()}
r64 = obeam.beam_search(olegacy.LegacyStepper(sd64, enc[i:i + 1].double(), k), 1, k, T, record_steps=True)
r32 = obeam.beam_search(olegacy.LegacyStepper(sd, enc[i:i + 1], k), 1, k, T, record_steps=True)
for t in range(s + 1):
    a = r64["steps"][t]["top_lp"][0].float(); b = r32["steps"][t]["top_lp"][0]
    print("fp64-vs-gpu", t, [f"{float(x):+.2e}" for x in (lp[t, i] - a)], " fp64-vs-cpu32", [f"{float(x):+.2e}" for x in (b - a)],
          " cpu32(alone)-vs-cpu32(batch)", [f"{float(x):+.2e}" for x in (b - ref_lp[t, i])])
