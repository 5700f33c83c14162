"""CPU restatement of the legacy Show-Attend-Tell decoder step (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/models/decoder.py:
  init state        :127,137-139   enc.view(B,-1,D); h = h_lin(mean_L enc); c = c_lin(mean_L enc)
  step body         :152-171       additive ReLU attention, softmax over regions, f_beta gate,
                                   LSTMCell([emb ; gated ctx]), fc(h)   (dropout = identity in eval)
  teacher forcing   :120-176       `forward` with shrinking batch_size_t (captions sorted by length)

Everything is torch fp32 on CPU, written over a plain state_dict (same parameter names as the
reference module) so the oracle needs neither the reference tree nor its imports at run time.
Pinned against the reference module itself in tests/test_oracle_pin.py and tests/golden/.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

PAD, START, END, UNK = 0, 1, 2, 3  # models/constants.py:1-4


def init_state(sd: Dict[str, torch.Tensor], enc: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """models/decoder.py:127,137-139.  enc [B,L,D] (or [B,14,14,D]) -> h,c [B,512]."""
    enc = enc.reshape(enc.size(0), -1, enc.size(-1))
    avg = enc.mean(dim=1)
    h = F.linear(avg, sd["h_lin.weight"], sd["h_lin.bias"])
    c = F.linear(avg, sd["c_lin.weight"], sd["c_lin.bias"])
    return h, c


def lstm_cell(sd, x, h, c, prefix="decode_step"):
    """torch.nn.LSTMCell semantics (gate order i,f,g,o), models/decoder.py:168."""
    gates = (F.linear(x, sd[f"{prefix}.weight_ih"], sd[f"{prefix}.bias_ih"])
             + F.linear(h, sd[f"{prefix}.weight_hh"], sd[f"{prefix}.bias_hh"]))
    i, f, g, o = gates.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def step(sd, enc, h, c, tokens, img_of_row=None):
    """One decode step for R rows.  models/decoder.py:152-171.

    enc [B,L,D]; h,c [R,512]; tokens int64 [R]; img_of_row int64 [R] maps rows to images
    (None => R == B, identity).  Returns logits [R,V], h', c', alpha [R,L].
    The reference recomputes enc_att(enc) every step (:152); so does this restatement, row-wise,
    so the arithmetic (operand order, reductions inside F.linear) is the reference's.
    """
    e = enc if img_of_row is None else enc[img_of_row]
    enc_att = F.linear(e, sd["enc_att.weight"], sd["enc_att.bias"])            # :152
    dec_att = F.linear(h, sd["dec_att.weight"], sd["dec_att.bias"])            # :153
    att = F.linear(torch.relu(enc_att + dec_att.unsqueeze(1)),
                   sd["att.weight"], sd["att.bias"]).squeeze(2)                # :154-155
    alpha = torch.softmax(att, dim=1)                                          # :156
    awe = (e * alpha.unsqueeze(2)).sum(dim=1)                                  # :157-158
    gate = torch.sigmoid(F.linear(h, sd["f_beta.weight"], sd["f_beta.bias"]))  # :160
    awe = gate * awe                                                           # :161
    emb = F.embedding(tokens, sd["embedding.weight"])                          # :132,163
    x = torch.cat([emb, awe], dim=1)                                           # :164 (.double().float() is value-preserving)
    h2, c2 = lstm_cell(sd, x, h, c)                                            # :168
    logits = F.linear(h2, sd["fc.weight"], sd["fc.bias"])                      # :171 (dropout identity in eval)
    return logits, h2, c2, alpha


def forward_teacher_forced(sd, enc, captions, caption_lengths):
    """models/decoder.py:120-176 -> (predictions [B,T,V], alphas [B,T,L], dec_len list)."""
    B = enc.size(0)
    enc = enc.reshape(B, -1, enc.size(-1))
    L = enc.size(1)
    V = sd["fc.weight"].size(0)
    dec_len = [int(x) - 1 for x in caption_lengths]
    T = max(dec_len)
    h, c = init_state(sd, enc)
    preds = torch.zeros(B, T, V)
    alphas = torch.zeros(B, T, L)
    for t in range(T):
        bt = sum(l > t for l in dec_len)
        logits, h, c, alpha = step(sd, enc[:bt], h[:bt], c[:bt], captions[:bt, t])
        preds[:bt, t] = logits
        alphas[:bt, t] = alpha
    return preds, alphas, dec_len


class LegacyStepper:
    """Adapter used by oracle.beam / oracle.sample: state = (h, c) per row, rows grouped by image."""

    def __init__(self, sd, enc, rows_per_image: int):
        self.sd = sd
        self.enc = enc.reshape(enc.size(0), -1, enc.size(-1))
        self.k = rows_per_image
        B = self.enc.size(0)
        self.img_of_row = torch.arange(B).repeat_interleave(rows_per_image)
        h, c = init_state(sd, self.enc)
        self.h = h.repeat_interleave(rows_per_image, 0)
        self.c = c.repeat_interleave(rows_per_image, 0)
        self.vocab_size = sd["fc.weight"].size(0)
        self.last_alpha = None

    def reorder(self, row_index: torch.Tensor):
        self.h = self.h[row_index]
        self.c = self.c[row_index]

    def __call__(self, tokens: torch.Tensor) -> torch.Tensor:
        logits, self.h, self.c, self.last_alpha = step(self.sd, self.enc, self.h, self.c, tokens, self.img_of_row)
        return logits
