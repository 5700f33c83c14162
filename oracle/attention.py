"""CPU restatement of the src/ attention mechanisms (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/models/attention.py for a 2-D query [R,H] (the only form the decoders
use, src/models/decoders.py:287-294):
  soft        SoftAttention.forward           :57-118
  multi_head  MultiHeadAttention.forward      :142-218
  adaptive    AdaptiveAttention.forward       :242-294
  aoa         AttentionOnAttention.forward    :322-360
Functional over a state_dict whose keys carry `prefix` (e.g. "attention.") -- same parameter names as
the reference modules.  key == value == image features [B,L,H]; `img_of_row` maps query rows to
images so beams of one image share its features.  `mask` is key_padding_mask (True = padding).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def soft(sd, p, q, feats, img_of_row=None, mask=None, temperature=1.0):
    kv = feats if img_of_row is None else feats[img_of_row]
    qp = _lin(sd, p + "query_proj", q).unsqueeze(1)                 # :76
    kp = _lin(sd, p + "key_proj", kv)                               # :77
    s = _lin(sd, p + "energy", torch.tanh(qp + kp)).squeeze(-1)     # :87-91
    s = s / temperature                                             # :94
    if mask is not None:
        m = mask if img_of_row is None else mask[img_of_row]
        s = s.masked_fill(m, -1e9)                                  # :97-100
    w = torch.softmax(s, dim=-1)                                    # :104
    ctx = torch.matmul(w.unsqueeze(1), kv).squeeze(1)               # :109-111
    return ctx, w


def multi_head(sd, p, q, feats, num_heads, img_of_row=None, mask=None, temperature=1.0):
    kv = feats if img_of_row is None else feats[img_of_row]
    R, H = q.shape
    d = H // num_heads
    qh = _lin(sd, p + "query_proj", q).view(R, 1, num_heads, d).transpose(1, 2)      # :161-170
    kh = _lin(sd, p + "key_proj", kv).view(R, -1, num_heads, d).transpose(1, 2)      # :172
    vh = _lin(sd, p + "value_proj", kv).view(R, -1, num_heads, d).transpose(1, 2)    # :174
    s = torch.matmul(qh, kh.transpose(-1, -2)) / (temperature * (d ** 0.5))          # :179-180
    if mask is not None:
        m = mask if img_of_row is None else mask[img_of_row]
        s = s.masked_fill(m[:, None, None, :], -1e9)                                 # :183-186
    w = torch.softmax(s, dim=-1)                                                     # :190
    a = torch.matmul(w, vh).transpose(1, 2).contiguous().view(R, 1, H)               # :195-202
    ctx = _lin(sd, p + "output_proj", a).squeeze(1)                                  # :205
    return ctx, w.mean(dim=1).squeeze(1)                                             # :211


def _base(sd, p, q, feats, num_heads, img_of_row, mask, temperature):
    # :229-230 / :308-309 -- MultiHeadAttention if num_heads > 1 else SoftAttention
    if num_heads > 1:
        return multi_head(sd, p + "base_attention.", q, feats, num_heads, img_of_row, mask, temperature)
    return soft(sd, p + "base_attention.", q, feats, img_of_row, mask, temperature)


def adaptive(sd, p, q, feats, num_heads, memory_state, cell_state, img_of_row=None, mask=None, temperature=1.0):
    g = torch.sigmoid(_lin(sd, p + "sentinel_gate", torch.cat([q, memory_state], dim=-1)))   # :266-269
    s = _lin(sd, p + "sentinel_proj", g * torch.tanh(cell_state))                            # :270-272
    ctx, w = _base(sd, p, q, feats, num_heads, img_of_row, mask, temperature)                # :275-277
    beta = torch.sigmoid(_lin(sd, p + "adaptive_weight", torch.cat([ctx, s], dim=-1)))       # :280-283
    return beta * ctx + (1 - beta) * s, w                                                    # :286-287


def aoa(sd, p, q, feats, num_heads, img_of_row=None, mask=None, temperature=1.0):
    ctx, w = _base(sd, p, q, feats, num_heads, img_of_row, mask, temperature)                # :338-340
    qt = _lin(sd, p + "query_proj", q)                                                       # :343
    cat = torch.cat([ctx, qt], dim=-1)                                                       # :346
    info = torch.tanh(_lin(sd, p + "info_vector_proj.0", cat))                               # :349
    gate = torch.sigmoid(_lin(sd, p + "info_gate_proj.0", cat))                              # :350
    return info * gate, w                                                                    # :353


def attend(kind: str, sd, p, q, feats, num_heads=8, img_of_row=None, mask=None, temperature=1.0,
           memory_state=None, cell_state=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """build_attention dispatch, src/models/attention.py:363-375."""
    if kind == "soft":
        return soft(sd, p, q, feats, img_of_row, mask, temperature)
    if kind == "multi_head":
        return multi_head(sd, p, q, feats, num_heads, img_of_row, mask, temperature)
    if kind == "adaptive":
        return adaptive(sd, p, q, feats, num_heads, memory_state, cell_state, img_of_row, mask, temperature)
    if kind == "aoa":
        return aoa(sd, p, q, feats, num_heads, img_of_row, mask, temperature)
    raise ValueError(f"Unsupported attention type: {kind}")
