"""Diagnostic (GPU box): where does the CUDA path lose precision?  Compares CUDA fp32 and the fp32 CPU oracle
against a float64 CPU oracle, stage by stage."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import capdec_b200 as cd
from capdec_b200 import engine as E
from oracle import attention as oatt, legacy as olegacy, lstm as olstm
from tests.helpers import *
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")

def err(a, b): return float((a.double().cpu() - b.double()).abs().max())

# 1. plain linear
g = torch.Generator().manual_seed(0)
a, w, b = torch.randn(300, 512, generator=g), torch.randn(256, 512, generator=g) * 0.05, torch.randn(256, generator=g)
ref64 = a.double() @ w.double().t() + b.double()
print("linear  cuda-fp32 vs f64 %.2e | cpu-f32 vs f64 %.2e" % (err(E.linear(a.to(dev), w.to(dev), b.to(dev)), ref64), err(torch.nn.functional.linear(a, w, b), ref64)))

# 2. attention modules
for kind, heads in (("soft", 1), ("multi_head", 8), ("aoa", 8), ("adaptive", 8)):
    H, L, B, rpi = 256, 49, 4, 3
    torch.manual_seed(11)
    mod = cd.build_attention(cd.AttentionConfig(attention_type=cd.AttentionType(kind), num_heads=heads, hidden_dim=H)).eval()
    sd = {"attention." + k: v.detach().clone() for k, v in mod.state_dict().items()}
    sd64 = {k: v.double() for k, v in sd.items()}
    g = torch.Generator().manual_seed(12)
    q, feats = torch.randn(B * rpi, H, generator=g), torch.randn(B, L, H, generator=g)
    mem, cell = torch.randn(B * rpi, H, generator=g), torch.randn(B * rpi, H, generator=g)
    img = torch.arange(B).repeat_interleave(rpi)
    c32, w32 = oatt.attend(kind, sd, "attention.", q, feats, heads, img, None, 1.0, mem, cell)
    c64, w64 = oatt.attend(kind, sd64, "attention.", q.double(), feats.double(), heads, img, None, 1.0, mem.double(), cell.double())
    mod = mod.to(dev); f = feats.to(dev)
    cc, wc = mod(q.to(dev), f, f, None, memory_state=mem.to(dev), cell_state=cell.to(dev), rows_per_image=rpi)
    print("%-10s ctx: cuda %.2e cpu32 %.2e | weights: cuda %.2e cpu32 %.2e (|w|max %.3f, |ctx|max %.3f)" % (
        kind, err(cc, c64), err(c32, c64), err(wc, w64), err(w32, w64), float(w64.abs().max()), float(c64.abs().max())))

# 3. legacy teacher-forced: per-step errors
V = 1000
m, sd = legacy_weights(V, 0)
sd64 = {k: v.double() for k, v in sd.items()}
enc = legacy_features(4)
caps = torch.randint(0, V, (4, 9), generator=torch.Generator().manual_seed(7)); lens = [9, 9, 9, 9]
p32, a32, _ = olegacy.forward_teacher_forced(sd, enc, caps, lens)
h, c = olegacy.init_state(sd64, enc.double().reshape(4, -1, 2048))
p64, a64 = [], []
for t in range(8):
    lg, h, c, al = olegacy.step(sd64, enc.double().reshape(4, -1, 2048), h, c, caps[:, t])
    p64.append(lg); a64.append(al)
p64, a64 = torch.stack(p64, 1), torch.stack(a64, 1)
pc, _, _, ac = m.to(dev)(enc.to(dev), caps.to(dev), lens)
for t in range(8):
    print("legacy t=%d logits: cuda %.2e cpu32 %.2e | alpha: cuda %.2e cpu32 %.2e" % (
        t, err(pc[:, t], p64[:, t]), err(p32[:, t], p64[:, t]), err(ac[:, t], a64[:, t]), err(a32[:, t], a64[:, t])))

# 4. lstm arch greedy: per-step alpha error vs f64 along the f64 path is not available through the API; compare
#    first-step alphas (state identical by construction)
for kind, heads, layers in (("soft", 8, 1), ("multi_head", 8, 2), ("aoa", 8, 1)):
    mm, sd = lstm_decoder(kind, H=256, layers=layers, heads=heads, V=2000, seed=1)
    sd64 = {k: v.double() for k, v in sd.items()}
    feats, pooled, _ = lstm_inputs(6, 49, 256, seed=21)
    ids32, al32 = olstm.generate_greedy(sd, feats, pooled, kind, layers, 8, num_heads=heads)
    ids64, al64 = olstm.generate_greedy(sd64, feats.double(), pooled.double(), kind, layers, 8, num_heads=heads)
    ids, info = mm.to(dev).generate({"features": feats.to(dev), "pooled_features": pooled.to(dev)}, 8)
    alc = info["attention_weights"]
    print(kind, "alpha err per step cuda:", ["%.1e" % err(alc[:, t], al64[:, t]) for t in range(8)])
    print(kind, "alpha err per step cpu32:", ["%.1e" % err(al32[:, t], al64[:, t]) for t in range(8)], "tokens equal", bool((ids.cpu() == ids64).all()))
