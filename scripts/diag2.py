import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import capdec_b200 as cd
from oracle import lstm as olstm
from tests.helpers import *
torch.set_grad_enabled(False)
dev = torch.device("cuda:0")
def err(a, b): return float((a.double().cpu() - b.double()).abs().max())
for kind, heads, layers, ragged in (("soft", 8, 2, False), ("multi_head", 8, 1, False), ("multi_head", 8, 2, False), ("multi_head", 8, 2, True), ("soft", 8, 3, False), ("multi_head", 2, 1, False)):
    mm, sd = lstm_decoder(kind, H=256, layers=layers, heads=heads, V=2000, seed=1)
    sd64 = {k: v.double() for k, v in sd.items()}
    feats, pooled, mask = lstm_inputs(6, 49, 256, seed=21, ragged=ragged)
    pm = None if mask is None else ~mask
    ids64, al64 = olstm.generate_greedy(sd64, feats.double(), pooled.double(), kind, layers, 6, num_heads=heads, mask=pm)
    ef = {"features": feats.to(dev), "pooled_features": pooled.to(dev)}
    if mask is not None: ef["attention_mask"] = mask.to(dev)
    ids, info = mm.to(dev).generate(ef, 6)
    alc = info["attention_weights"]
    print(kind, heads, layers, ragged, "alpha err per step cuda:", ["%.1e" % err(alc[:, t], al64[:, t]) for t in range(6)], bool((ids.cpu() == ids64).all()))
