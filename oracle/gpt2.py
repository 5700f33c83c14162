"""Oracle for the GPT-2 decoder path (TEST INFRASTRUCTURE ONLY).

The arithmetic of /root/reference/src/models/decoders.py::GPT2Decoder lives in the third-party package
`transformers` (requirements.txt:3 pins only `>=4.20.0`; 5.5.0 is installed here and on the GPU box): the GPT-2 block
(`GPT2LMHeadModel`, pre-LN, gelu_new, tied lm_head) and `GenerationMixin._beam_search`.  Call sites in the reference:
model construction decoders.py:513,531; `self.model.generate(num_beams=..., past_key_values=prefix)` :645.
As shipped that call raises on every transformers version (the prefix is passed as a list of [B,P,hidden] tensors,
SURVEY.md section 0.4), so parity is anchored on the intended computation SURVEY section 8(c) pinned by probe:

    prefix = image_to_prefix(pooled).view(B, P, n_embd)                      decoders.py:634-637
    past K == past V == prefix.view(B, P, n_head, d).transpose(1, 2)         for EVERY layer (decoders.py:608-615)
    generate(input_ids=[bos], num_beams=k, max_length=T, pad/bos/eos ids, attention_mask=ones[B, P+1])

`hf_generate` runs exactly that through transformers itself; `HFStepper` exposes the same model as a stepper so the
pinned beam driver (oracle/beam.py) can record per-step candidates for the 1e-3 log-prob comparison.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _prefix_cache(model, prefix, rows_per_image):
    from transformers import DynamicCache
    cfg = model.config
    B, P, H = prefix.shape
    kv = prefix.view(B, P, cfg.n_head, H // cfg.n_head).transpose(1, 2).repeat_interleave(rows_per_image, 0)
    return DynamicCache(ddp_cache_data=[(kv.clone(), kv.clone()) for _ in range(cfg.n_layer)])


def image_prefix(sd, pooled, n_embd):
    out = F.linear(pooled, sd["image_to_prefix.weight"], sd["image_to_prefix.bias"])
    return out.view(pooled.size(0), -1, n_embd)


@torch.no_grad()
def hf_generate(model, sd, pooled, num_beams, max_length, pad=0, bos=1, eos=2, length_penalty=1.0):
    """transformers' own beam search on the prefix-conditioned model -> (sequences, sequences_scores)."""
    prefix = image_prefix(sd, pooled, model.config.n_embd)
    B, P, _ = prefix.shape
    out = model.generate(input_ids=torch.full((B, 1), bos), max_length=max_length, num_beams=num_beams, do_sample=False,
                         pad_token_id=pad, bos_token_id=bos, eos_token_id=eos, length_penalty=length_penalty,
                         past_key_values=_prefix_cache(model, prefix, num_beams),
                         attention_mask=torch.ones(B, P + 1, dtype=torch.long), return_dict_in_generate=True,
                         output_scores=True)
    return out.sequences, out.sequences_scores


class HFStepper:
    """GPT2LMHeadModel + DynamicCache as a stepper for oracle.beam / oracle.sample (rows grouped by image)."""

    def __init__(self, model, sd, pooled, rows_per_image=1):
        self.model = model
        prefix = image_prefix(sd, pooled, model.config.n_embd)
        self.P = prefix.size(1)
        self.cache = _prefix_cache(model, prefix, rows_per_image)
        self.n = 0
        self.vocab_size = model.config.vocab_size

    def reorder(self, idx):
        self.cache.reorder_cache(idx)

    @torch.no_grad()
    def __call__(self, tokens):
        R = tokens.size(0)
        out = self.model(input_ids=tokens[:, None], past_key_values=self.cache, use_cache=True,
                         attention_mask=torch.ones(R, self.P + self.n + 1, dtype=torch.long),
                         position_ids=torch.full((R, 1), self.P + self.n))
        self.n += 1
        return out.logits[:, -1].float()
