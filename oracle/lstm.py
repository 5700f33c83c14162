"""CPU restatement of src LSTMDecoder decode (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/models/decoders.py:
  _init_hidden_states   :122-135   h0 = init_h(pooled).view(B,layers,H).transpose(0,1)  (same for c0)
  generate (greedy)     :236-314   out[:,t]=tok; x=[emb(tok); prev_ctx]; nn.LSTM one step;
                                   ctx,a = attention(q=top h, feats, feats, mask, h[-1], c[-1]);
                                   logits = output_layer(ctx); tok = argmax.  Runs exactly max_length
                                   steps, position 0 holds the start token, the last argmax is
                                   discarded, there is no EOS handling (reference behaviour, kept).
nn.LSTM (eval mode, inter-layer dropout inactive) is restated as stacked cells with torch's gate
order i,f,g,o.  Functional over the decoder's state_dict (reference parameter names).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import attention as A


def init_hidden(sd, pooled, num_layers):
    """decoders.py:122-135 -> h0,c0 [layers,B,H]."""
    B = pooled.size(0)
    H = sd["init_h.weight"].size(1)
    h0 = F.linear(pooled, sd["init_h.weight"], sd["init_h.bias"]).view(B, num_layers, H).transpose(0, 1).contiguous()
    c0 = F.linear(pooled, sd["init_c.weight"], sd["init_c.bias"]).view(B, num_layers, H).transpose(0, 1).contiguous()
    return h0, c0


def lstm_step(sd, x, h, c, num_layers):
    """One time step of nn.LSTM(batch_first) in eval mode.  x [R,E+H]; h,c [layers,R,H]."""
    hs, cs = [], []
    inp = x
    for l in range(num_layers):
        g = (F.linear(inp, sd[f"lstm.weight_ih_l{l}"], sd[f"lstm.bias_ih_l{l}"])
             + F.linear(h[l], sd[f"lstm.weight_hh_l{l}"], sd[f"lstm.bias_hh_l{l}"]))
        i, f, gg, o = g.chunk(4, dim=1)
        c2 = torch.sigmoid(f) * c[l] + torch.sigmoid(i) * torch.tanh(gg)
        h2 = torch.sigmoid(o) * torch.tanh(c2)
        hs.append(h2); cs.append(c2)
        inp = h2
    return inp, torch.stack(hs), torch.stack(cs)


class LSTMStepper:
    """decoders.py:274-303 as a stepper: state = (h, c, prev_ctx); rows grouped by image."""

    def __init__(self, sd, feats, pooled, kind, num_layers, num_heads=8, rows_per_image=1, mask=None,
                 temperature=1.0):
        self.sd, self.feats, self.kind = sd, feats, kind
        self.num_layers, self.num_heads, self.temperature = num_layers, num_heads, temperature
        self.mask = mask                                   # key_padding_mask, True = padding
        B = feats.size(0)
        self.img_of_row = torch.arange(B).repeat_interleave(rows_per_image)
        h0, c0 = init_hidden(sd, pooled, num_layers)
        self.h = h0.repeat_interleave(rows_per_image, 1)
        self.c = c0.repeat_interleave(rows_per_image, 1)
        self.prev_ctx = torch.zeros(B * rows_per_image, sd["init_h.weight"].size(1))
        self.vocab_size = sd["output_layer.weight"].size(0)
        self.last_alpha = None

    def reorder(self, idx):
        self.h = self.h[:, idx]
        self.c = self.c[:, idx]
        self.prev_ctx = self.prev_ctx[idx]

    def __call__(self, tokens):
        sd = self.sd
        emb = F.embedding(tokens, sd["embedding.weight"])                               # :274
        x = torch.cat([emb, self.prev_ctx], dim=1)                                      # :277
        q, self.h, self.c = lstm_step(sd, x, self.h, self.c, self.num_layers)           # :281-284
        ctx, alpha = A.attend(self.kind, sd, "attention.", q, self.feats, self.num_heads,
                              self.img_of_row, self.mask, self.temperature,
                              memory_state=self.h[-1], cell_state=self.c[-1])           # :287-294
        self.prev_ctx = ctx                                                             # :297
        self.last_alpha = alpha
        return F.linear(ctx, sd["output_layer.weight"], sd["output_layer.bias"])        # :303


@torch.no_grad()
def generate_greedy(sd, feats, pooled, kind, num_layers, max_length, num_heads=8, start_token_id=1,
                    mask=None, temperature=1.0, return_margins=False):
    """decoders.py:236-314.  Returns (output_ids [B,T] int64, attention_weights [B,T,L][, margins [B,T]])."""
    st = LSTMStepper(sd, feats, pooled, kind, num_layers, num_heads, 1, mask, temperature)
    B = feats.size(0)
    cur = torch.full((B,), start_token_id, dtype=torch.long)
    out = torch.zeros(B, max_length, dtype=torch.long)
    alphas, margins = [], []
    for t in range(max_length):
        out[:, t] = cur
        logits = st(cur)
        alphas.append(st.last_alpha)
        top2 = logits.topk(2, dim=1).values
        margins.append((top2[:, 0] - top2[:, 1]) / logits.std(dim=1))   # top-1/top-2 gap in units of the logit spread
        cur = logits.argmax(dim=1)
    res = (out, torch.stack(alphas, dim=1))
    if return_margins:
        res = res + (torch.stack(margins, dim=1),)
    return res
