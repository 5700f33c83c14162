"""Teacher-forced passes over given captions (TEST INFRASTRUCTURE ONLY): CPU restatement of what the reference's
`decoder(encoder_features, captions=ids)["logits"]` returns, the call CaptioningTrainer._sample_captions makes once
per generated token (/root/reference/src/train/trainer.py:413-420).

  lstm_logits          src/models/decoders.py:137-234   step t consumes captions[:, t] and the previous context;
                                                        logits[:, t] = output_layer(context_t)  (dropout = identity in eval;
                                                        the caption_lengths sort / un-sort has no effect on the result)
  transformer_logits   src/models/decoders.py:377-438   ONE full-prefix pass of nn.TransformerDecoder with the causal
                                                        tgt_mask (:400) and tgt_key_padding_mask = captions == pad (:405);
                                                        region mask as additive -1e9 on the memory keys (:393-398 intent)
  gpt2_logits          src/models/decoders.py:563-596   GPT2LMHeadModel over captions behind the image prefix bound as
                                                        oracle/gpt2.py pins it; attention_mask = captions != pad (:581)
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import gpt2 as G, lstm as L, transformer as T


@torch.no_grad()
def lstm_logits(sd, feats, pooled, kind, num_layers, num_heads, captions, mask=None, temperature=1.0, rows_per_image=1):
    """-> (logits [R,t,V], attention_weights [R,t,L]);  mask = key_padding_mask (True = padding)."""
    st = L.LSTMStepper(sd, feats, pooled, kind, num_layers, num_heads, rows_per_image, mask, temperature)
    logits, alphas = [], []
    for t in range(captions.size(1)):
        logits.append(st(captions[:, t]))
        alphas.append(st.last_alpha)
    return torch.stack(logits, dim=1), torch.stack(alphas, dim=1)


@torch.no_grad()
def transformer_logits(sd, feats, num_layers, num_heads, captions, pad_token_id=0, region_padding_mask=None,
                       rows_per_image=1):
    """-> logits [R,t,V], one pass over the whole prefix exactly like decoders.py:377-438."""
    dec = T.build_modules(sd, num_layers, num_heads)
    mem = F.linear(feats, sd["visual_projection.weight"], sd["visual_projection.bias"]).repeat_interleave(rows_per_image, 0)
    n = captions.size(1)
    x = F.embedding(captions, sd["embedding.weight"]) + F.embedding(torch.arange(n), sd["position_encoding.weight"])[None]
    tgt_mask = nn.Transformer.generate_square_subsequent_mask(n).to(x.dtype)
    mkpm = None
    if region_padding_mask is not None:      # additive -1e9 on padded region keys
        mkpm = (region_padding_mask.to(x.dtype) * -1e9).repeat_interleave(rows_per_image, 0)
    kpm = (captions == pad_token_id)
    kpm_f = torch.zeros(kpm.shape, dtype=x.dtype).masked_fill(kpm, float("-inf"))   # same type as tgt_mask (no torch warning)
    out = dec(tgt=x, memory=mem, tgt_mask=tgt_mask, tgt_key_padding_mask=kpm_f, memory_key_padding_mask=mkpm)
    return F.linear(out, sd["output_layer.weight"], sd["output_layer.bias"])


@torch.no_grad()
def gpt2_logits(model, sd, pooled, captions, pad_token_id=0, rows_per_image=1):
    """-> (logits [R,t,V] fp32, loss) of transformers' GPT2LMHeadModel itself."""
    prefix = G.image_prefix(sd, pooled, model.config.n_embd)
    P = prefix.size(1)
    R, n = captions.shape
    am = torch.cat([torch.ones(R, P, dtype=torch.long), (captions != pad_token_id).long()], dim=1)
    out = model(input_ids=captions, past_key_values=G._prefix_cache(model, prefix, rows_per_image), attention_mask=am,
                position_ids=(P + torch.arange(n))[None].expand(R, n), labels=captions, use_cache=True)
    return out.logits.float(), out.loss.float()


@torch.no_grad()
def token_logprobs(logits, captions):
    """log p(captions[:, t+1]) under logits[:, t]  -> [R, t-1]"""
    lp = torch.log_softmax(logits[:, :-1].float(), dim=-1)
    return lp.gather(2, captions[:, 1:, None]).squeeze(2)
