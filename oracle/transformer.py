"""CPU restatement of src TransformerDecoder decode (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/models/decoders.py:
  __init__   :343-375   nn.Embedding x2, nn.TransformerDecoder(nn.TransformerDecoderLayer(d_model, nhead, 4*d_model,
                        dropout, activation="gelu", batch_first=True), num_layers), output_layer, visual_projection
  generate   :439-493   memory = visual_projection(features); per step the WHOLE prefix is re-run through the decoder
                        with a causal mask (no KV cache), logits of the last position, argmax, cat, break only when
                        every row emitted EOS at the same step.
The reference composes stock torch.nn modules, so the restatement instantiates the same torch.nn modules from the
decoder's state_dict (reference parameter names) and re-runs the prefix exactly like the reference does.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def build_modules(sd, num_layers, num_heads):
    H = sd["embedding.weight"].size(1)
    layer = nn.TransformerDecoderLayer(d_model=H, nhead=num_heads, dim_feedforward=sd["transformer_decoder.layers.0.linear1.weight"].size(0),
                                       dropout=0.0, activation="gelu", batch_first=True)
    dec = nn.TransformerDecoder(layer, num_layers=num_layers)
    dec.load_state_dict({k[len("transformer_decoder."):]: v for k, v in sd.items() if k.startswith("transformer_decoder.")})
    dec = dec.to(sd["embedding.weight"].dtype)
    return dec.eval()


class TransformerStepper:
    """decoders.py:463-483 as a stepper: state = token prefix per row (recomputed every step); rows grouped by image."""

    def __init__(self, sd, feats, num_layers, num_heads, rows_per_image=1):
        self.sd = sd
        self.dec = build_modules(sd, num_layers, num_heads)
        mem = F.linear(feats, sd["visual_projection.weight"], sd["visual_projection.bias"])      # :453
        self.mem = mem.repeat_interleave(rows_per_image, 0)
        self.prefix = None
        self.vocab_size = sd["output_layer.weight"].size(0)

    def reorder(self, idx):
        self.prefix = self.prefix[idx]

    @torch.no_grad()
    def __call__(self, tokens):
        self.prefix = tokens[:, None] if self.prefix is None else torch.cat([self.prefix, tokens[:, None]], 1)
        ids = self.prefix
        n = ids.size(1)
        x = F.embedding(ids, self.sd["embedding.weight"]) + F.embedding(torch.arange(n), self.sd["position_encoding.weight"])[None]  # :468-469
        mask = nn.Transformer.generate_square_subsequent_mask(n).to(x.dtype)                                                          # :472
        out = self.dec(tgt=x, memory=self.mem, tgt_mask=mask)                                                                          # :476-480
        return F.linear(out[:, -1], self.sd["output_layer.weight"], self.sd["output_layer.bias"])                                     # :483


@torch.no_grad()
def generate_greedy(sd, feats, num_layers, num_heads, max_length, bos_token_id=1, eos_token_id=2, return_margins=False):
    """decoders.py:439-493 -> input_ids [B, <= max_length] (int64)."""
    st = TransformerStepper(sd, feats, num_layers, num_heads)
    B = feats.size(0)
    ids = torch.full((B, 1), bos_token_id, dtype=torch.long)
    margins = []
    for _ in range(max_length - 1):
        logits = st(ids[:, -1])
        top2 = logits.topk(2, dim=1).values
        margins.append((top2[:, 0] - top2[:, 1]) / logits.std(dim=1))
        nxt = logits.argmax(dim=-1, keepdim=True)
        ids = torch.cat([ids, nxt], dim=1)
        if bool((nxt == eos_token_id).all()):
            break
    if return_margins:
        return ids, torch.stack(margins, dim=1)
    return ids
