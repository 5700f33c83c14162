"""Beam search oracle (TEST INFRASTRUCTURE ONLY).

The reference contains no hand-written beam search; the only one it ever invokes is HuggingFace's
`GPT2LMHeadModel.generate(num_beams=...)` (/root/reference/src/models/decoders.py:645).  That
algorithm lives in the un-vendored third-party package `transformers` (requirements.txt:3 pins
only `>=4.20.0`; 5.5.0 is what is installed here).  This file restates its published static-shape
algorithm -- transformers/generation/utils.py::GenerationMixin._beam_search and its helpers
_get_top_k_continuations / _get_running_beams_for_next_iteration / _update_finished_beams /
_check_early_stop_heuristic / _beam_search_has_unfinished_sequences -- over a generic
`stepper(tokens[R]) -> logits[R,V]` with `stepper.reorder(row_index[R])`, so the same driver wraps
the legacy LSTM step, the src LSTMDecoder step, and (for pinning) a real HF GPT-2.

HF defaults that apply at the reference call site: length_penalty=1.0, early_stopping=False,
one EOS id => beams_to_keep = 2*num_beams, stopping criteria = {MaxLength, EosToken}.
Pinned against `transformers` itself in tests/test_oracle_pin.py::test_beam_driver_matches_hf.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

NEG = -1.0e9


def _gather(t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    while idx.dim() < t.dim():
        idx = idx.unsqueeze(-1)
    return torch.take_along_dim(t, idx, dim=1)


@torch.no_grad()
def beam_search(stepper, batch_size: int, num_beams: int, max_length: int, bos_token_id: int = 1,
                eos_token_id: Optional[int] = 2, pad_token_id: int = 0, length_penalty: float = 1.0,
                early_stopping=False, record_steps: bool = False) -> Dict[str, torch.Tensor]:
    """Returns dict(sequences [B,max_length] int64 padded with pad, lengths [B], scores [B],
    all_sequences [B,k,max_length], all_scores [B,k], steps=[per-step dict] if record_steps)."""
    B, k = batch_size, num_beams
    V = stepper.vocab_size
    keep = 2 * k                                       # max(2, 1 + n_eos) * num_beams
    prompt_len = 1
    cur_len = 1
    top_mask = torch.cat([torch.ones(k, dtype=torch.bool), torch.zeros(keep - k, dtype=torch.bool)])

    fill = pad_token_id if pad_token_id else (eos_token_id if eos_token_id is not None else -1)
    running_seq = torch.full((B, k, max_length), fill, dtype=torch.int64)
    running_seq[:, :, 0] = bos_token_id
    sequences = running_seq.clone()
    running_scores = torch.zeros(B, k)
    running_scores[:, 1:] = NEG
    beam_scores = torch.full((B, k), NEG)
    finished = torch.zeros(B, k, dtype=torch.bool)
    unsatisfied = torch.ones(B, 1, dtype=torch.bool)
    running_beam_idx = torch.full((B, k, max_length - 1), -1, dtype=torch.int32)
    beam_idx_fin = running_beam_idx.clone()
    steps = []

    while True:
        tokens = running_seq[:, :, cur_len - 1].reshape(B * k)
        logits = stepper(tokens).to(torch.float32)
        log_probs = torch.log_softmax(logits, dim=-1).view(B, k, V)
        acc = (log_probs + running_scores[:, :, None]).reshape(B, k * V)

        # _get_top_k_continuations
        top_lp, top_idx = torch.topk(acc, k=keep)
        top_beam = top_idx // V
        top_tok = top_idx % V
        top_seq = _gather(running_seq, top_beam)
        top_seq[:, :, cur_len] = top_tok
        top_bidx = _gather(running_beam_idx, top_beam)
        top_bidx[:, :, cur_len - prompt_len] = (top_beam + torch.arange(B).view(-1, 1) * k).to(torch.int32)

        # stopping criteria: MaxLengthCriteria | EosTokenCriteria on the just-extended sequences
        hits = torch.full((B, keep), cur_len + 1 >= max_length, dtype=torch.bool)
        if eos_token_id is not None:
            hits = hits | (top_tok == eos_token_id)

        # _get_running_beams_for_next_iteration
        run_lp = top_lp + hits.to(torch.float32) * NEG
        nxt = torch.topk(run_lp, k=k)[1]
        running_seq = _gather(top_seq, nxt)
        running_scores = _gather(run_lp, nxt)
        running_beam_idx = _gather(top_bidx, nxt)
        src_beam = _gather(top_beam, nxt)               # back-pointer: which old beam each new beam extends

        # _update_finished_beams
        just_fin = hits & top_mask[None, :]
        fin_lp = top_lp / ((cur_len + 1 - prompt_len) ** length_penalty)
        full = torch.all(finished, dim=-1, keepdim=True) & (early_stopping is True)
        fin_lp = fin_lp + full.to(torch.float32) * NEG
        fin_lp = fin_lp + (~unsatisfied).to(torch.float32) * NEG
        fin_lp = fin_lp + (~just_fin) * NEG
        m_seq = torch.cat([sequences, top_seq], dim=1)
        m_sc = torch.cat([beam_scores, fin_lp], dim=1)
        m_bidx = torch.cat([beam_idx_fin, top_bidx], dim=1)
        m_fin = torch.cat([finished, just_fin], dim=1)
        sel = torch.topk(m_sc, k=k)[1]
        sequences = _gather(m_seq, sel)
        beam_scores = _gather(m_sc, sel)
        beam_idx_fin = _gather(m_bidx, sel)
        finished = _gather(m_fin, sel)

        if record_steps:
            steps.append(dict(top_lp=top_lp.clone(), top_tok=top_tok.clone(), top_beam=top_beam.clone(),
                              running_scores=running_scores.clone(), src_beam=src_beam.clone(),
                              running_tok=running_seq[:, :, cur_len].clone()))

        # reorder model state by back-pointer (HF: cache.reorder_cache(beam_idx))
        stepper.reorder((src_beam + torch.arange(B).view(-1, 1) * k).reshape(B * k))

        cur_len += 1
        # _check_early_stop_heuristic
        if early_stopping == "never" and length_penalty > 0.0:
            best_len = max_length - prompt_len
        else:
            best_len = cur_len - prompt_len
        best_possible = running_scores[:, :1] / (best_len ** length_penalty)
        worst_fin = torch.where(finished, torch.min(beam_scores, dim=1, keepdim=True)[0],
                                torch.tensor(NEG))
        unsatisfied = unsatisfied & torch.any(best_possible > worst_fin, dim=-1, keepdim=True)
        # _beam_search_has_unfinished_sequences
        improvement_possible = bool(torch.any(unsatisfied))
        exists_open_beam = not (bool(torch.all(finished)) and (early_stopping is True))
        valid_continuations = not bool(torch.all(hits))
        if not (improvement_possible and exists_open_beam and valid_continuations):
            break

    best = sequences[:, 0]
    gen_len = (beam_idx_fin[:, 0] + 1).bool().sum(dim=1)
    out = dict(sequences=best, lengths=gen_len + prompt_len, scores=beam_scores[:, 0],
               all_sequences=sequences, all_scores=beam_scores, n_steps=cur_len - 1)
    if record_steps:
        out["steps"] = steps
    return out


def crop_like_hf(sequences: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
    """HF crops the static [B,max_length] buffer to the longest returned hypothesis (utils.py:3383-3385)."""
    return sequences[:, : int(lengths.max())]
