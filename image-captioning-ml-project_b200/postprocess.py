"""Token -> text boundary behind the decode path.

The reference turns generated ids into text one caption at a time, `tokenizer.decode(caption, skip_special_tokens=True)`
inside a Python loop over a device tensor (src/train/trainer.py:546-547, src/evaluate/metrics.py:322-323,
src/main.py demo), i.e. one device->host synchronisation per caption, and writes COCO results as
`[{"image_id": int, "caption": str}, ...]` (src/evaluate/metrics.py:326-336).  At 10^5 captions/s that loop is the
bottleneck, so this module keeps the same outputs but does the per-caption work vectorised on the device and crosses
to the host once per batch:

  trim_at_eos      lengths + padding after the first EOS, on whatever device the tokens live on (no sync)
  to_token_lists   ONE device->host copy of the int32 [B,T] block, then plain Python lists
  decode_captions  `tokenizer.batch_decode` when the tokenizer has it, else the reference's per-caption `decode`
  coco_results / write_results_json   the reference's results format

Tokens that live on a CUDA device are trimmed by libcapdec's `capdec_trim_at_eos` kernel (one launch, no
synchronisation); the torch formulation below it serves host tensors only.  Beam-search output needs no trimming at
all: `capdec_decode_beam` already returns `lengths` and fills everything behind the EOS (pass `lengths=` through).
"""
import json
from typing import Iterable, List, Optional, Sequence, Tuple

import torch


def trim_at_eos(tokens: torch.Tensor, eos_token_id: int, pad_token_id: int = 0,
                keep_eos: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """tokens [B,T] (any integer dtype, any device) -> (trimmed [B,T], lengths [B] int64).

    Everything after a row's first EOS is replaced by `pad_token_id`; `lengths` counts the tokens kept (including
    the EOS itself when `keep_eos`).  Rows without an EOS keep all T tokens.  (The reference's greedy loops keep
    writing argmax tokens after EOS -- decoders.py:300-306, 481-491 -- and rely on `skip_special_tokens` plus the
    tokenizer to hide them; trimming first makes the text independent of what follows the EOS.)"""
    if tokens.dim() != 2:
        raise ValueError(f"tokens must be [B,T], got {tuple(tokens.shape)}")
    if tokens.is_cuda:
        from .engine import trim_at_eos_device
        out, lengths = trim_at_eos_device(tokens.to(torch.int32), eos_token_id, pad_token_id, keep_eos)
        return out.to(tokens.dtype), lengths.to(torch.int64)
    B, T = tokens.shape
    is_eos = tokens == eos_token_id
    pos = torch.arange(T, device=tokens.device).expand(B, T)
    first = torch.where(is_eos, pos, torch.full_like(pos, T)).min(dim=1).values      # T when the row has no EOS
    lengths = torch.clamp(first + (1 if keep_eos else 0), max=T)
    keep = pos < lengths.unsqueeze(1)
    return torch.where(keep, tokens, torch.full_like(tokens, pad_token_id)), lengths.to(torch.int64)


def to_token_lists(tokens: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                   skip_ids: Iterable[int] = ()) -> List[List[int]]:
    """One device->host transfer for the whole batch; rows cut to `lengths` and stripped of `skip_ids`."""
    host = tokens.detach().to("cpu", non_blocking=False)
    lens = None if lengths is None else lengths.detach().to("cpu").tolist()
    skip = set(int(s) for s in skip_ids)
    out = []
    for i, row in enumerate(host.tolist()):
        if lens is not None:
            row = row[: int(lens[i])]
        out.append([t for t in row if t not in skip] if skip else row)
    return out


def decode_captions(tokens: torch.Tensor, tokenizer, eos_token_id: Optional[int] = None, pad_token_id: int = 0,
                    skip_special_tokens: bool = True, lengths: Optional[torch.Tensor] = None) -> List[str]:
    """The reference's `[tokenizer.decode(c, skip_special_tokens=True) for c in captions]` for a [B,T] block:
    trimmed at EOS on the device, copied once, decoded with `batch_decode` when available.  `lengths` (beam search
    returns them) skips the trimming pass."""
    if eos_token_id is None:
        eos_token_id = getattr(tokenizer, "eos_token_id", None)
    if lengths is None and eos_token_id is not None:
        tokens, lengths = trim_at_eos(tokens, int(eos_token_id), pad_token_id)
    rows = to_token_lists(tokens, lengths)
    if hasattr(tokenizer, "batch_decode"):
        return list(tokenizer.batch_decode(rows, skip_special_tokens=skip_special_tokens))
    return [tokenizer.decode(r, skip_special_tokens=skip_special_tokens) for r in rows]


def coco_results(image_ids: Sequence, captions: Sequence[str]) -> List[dict]:
    """`[{"image_id": int, "caption": str}]`, the list src/evaluate/metrics.py:326-336 builds and dumps."""
    if len(image_ids) != len(captions):
        raise ValueError(f"{len(image_ids)} image ids for {len(captions)} captions")
    ids = image_ids.tolist() if isinstance(image_ids, torch.Tensor) else list(image_ids)
    return [{"image_id": int(i), "caption": str(c)} for i, c in zip(ids, captions)]


def write_results_json(path: str, image_ids: Sequence, captions: Sequence[str]) -> str:
    with open(path, "w") as f:
        json.dump(coco_results(image_ids, captions), f)
    return path
