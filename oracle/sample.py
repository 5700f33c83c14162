"""Greedy / ancestral-sampling rollouts over a stepper (TEST INFRASTRUCTURE ONLY).

sample_rollout restates /root/reference/src/train/trainer.py:383-438 (`_sample_captions`):
  per step  logits -> softmax (:423) -> Categorical.sample (:424-425) -> log_prob of the draw (:428)
            -> append (:432) -> break when every row drew EOS (:435).
torch's CPU RNG stream cannot be reproduced by a CUDA kernel, so -- as SURVEY.md section 8(c) fixes --
both sides consume the same pre-generated uniforms u[R,steps] and draw by inverse CDF over the
vocabulary in index order: token = #{v : cdf[v] <= u}.  The CDF here is accumulated in float64 so a
disagreement can only come from the kernel's rounding; `sample_rollout` also returns the distance of u
to the nearest CDF edge so tests can excuse boundary cases explicitly.

greedy_rollout is the legacy path's free-running argmax decode with LSTMDecoder.generate's
conventions (src/models/decoders.py:269-306): position 0 = start token, exactly max_length steps
evaluated, last argmax discarded, no EOS stop.

sample_captions_loop is the SAME loop driven through a `decoder_forward(captions) -> {"logits": [R,t,V]}` callable,
i.e. literally the call structure of trainer.py:413-436 (full teacher-forced forward on the growing prefix, last
position's logits); it is what runs against the drop-in modules' `forward(captions=...)`.
"""
from __future__ import annotations

import torch


@torch.no_grad()
def greedy_rollout(stepper, num_rows, max_length, start_token_id=1):
    cur = torch.full((num_rows,), start_token_id, dtype=torch.long)
    out = torch.zeros(num_rows, max_length, dtype=torch.long)
    margins = []
    for t in range(max_length):
        out[:, t] = cur
        logits = stepper(cur)
        top2 = logits.topk(2, dim=1).values
        margins.append((top2[:, 0] - top2[:, 1]) / logits.std(dim=1))   # gap in units of the logit spread
        cur = logits.argmax(dim=1)
    return out, torch.stack(margins, dim=1)


@torch.no_grad()
def rescore(stepper, sequences, lengths, length_penalty=1.0):
    """Oracle score of given hypotheses (one row per image): sum of log p(token) over the generated tokens
    divided by generated_length ** length_penalty -- what HF's finished-beam score is for that sequence."""
    R, T = sequences.shape
    total = torch.zeros(R)
    for t in range(T - 1):
        logp = torch.log_softmax(stepper(sequences[:, t]).float(), dim=-1)
        step_lp = logp.gather(1, sequences[:, t + 1:t + 2].clamp(min=0)).squeeze(1)
        total += torch.where(t + 1 < lengths, step_lp, torch.zeros(R))
    return total / ((lengths - 1).float() ** length_penalty)


@torch.no_grad()
def sample_rollout(stepper, num_rows, max_length, uniforms, bos_token_id=1, eos_token_id=2,
                   forced_tokens=None):
    """uniforms [R, max_length-1] in [0,1).  Returns (ids [R,<=T], log_probs [R,steps], edge_dist [R,steps]).
    If forced_tokens [R, steps] is given the draw is replaced by those tokens (used to compare
    per-step log-probs along an identical path)."""
    ids = torch.full((num_rows, 1), bos_token_id, dtype=torch.long)
    lps, edges = [], []
    for t in range(max_length - 1):
        logits = stepper(ids[:, -1]).to(torch.float32)
        logp = torch.log_softmax(logits, dim=-1)
        cdf = torch.softmax(logits.double(), dim=-1).cumsum(dim=-1)
        u = uniforms[:, t].double().unsqueeze(1)
        tok = (cdf <= u).sum(dim=1).clamp(max=logits.size(1) - 1)
        edges.append((cdf - u).abs().min(dim=1).values.float())
        if forced_tokens is not None:
            tok = forced_tokens[:, t]
        lps.append(logp.gather(1, tok[:, None]).squeeze(1))
        ids = torch.cat([ids, tok[:, None]], dim=1)
        if bool((tok == eos_token_id).all()):
            break
    return ids, torch.stack(lps, dim=1), torch.stack(edges, dim=1)


@torch.no_grad()
def sample_captions_loop(decoder_forward, num_rows, max_length, uniforms, bos_token_id=1, eos_token_id=2, device="cpu"):
    """trainer.py:383-438 with the Categorical draw replaced by the inverse-CDF draw on shared uniforms.
    decoder_forward(input_ids int64 [R,t]) must return {"logits": [R,t,V]}.  -> (input_ids [R,<=T], log_probs [R,steps])"""
    input_ids = torch.full((num_rows, 1), bos_token_id, dtype=torch.long, device=device)          # :402-407
    log_probs = []
    for t in range(max_length - 1):                                                             # :413
        logits = decoder_forward(input_ids)["logits"][:, -1, :].float()                         # :415-420
        probs = torch.softmax(logits, dim=-1)                                                   # :423
        cdf = probs.double().cumsum(dim=-1)
        u = uniforms[:, t].to(device).double().unsqueeze(1)
        next_token = (cdf <= u).sum(dim=1).clamp(max=logits.size(1) - 1)                        # :424-425 (draw)
        log_probs.append(torch.log(probs.gather(1, next_token[:, None]).squeeze(1)))            # :428 Categorical.log_prob
        input_ids = torch.cat([input_ids, next_token.unsqueeze(1)], dim=1)                      # :432
        if bool((next_token == eos_token_id).all()):                                            # :435
            break
    return input_ids, torch.stack(log_probs, dim=1)
