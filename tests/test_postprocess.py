"""Token -> text boundary: the vectorised helpers give what the reference's per-caption loop
gives (src/train/trainer.py:546-547, src/evaluate/metrics.py:322-336)."""
import json
import random

import pytest
import torch

import capdec_b200 as cd


class FakeTokenizer:
    """`decode(ids, skip_special_tokens)` like a HF tokenizer over a toy vocabulary; specials: 0 pad, 1 bos, 2 eos."""
    eos_token_id = 2
    special = {0, 1, 2}

    def decode(self, ids, skip_special_tokens=True):
        ids = ids.tolist() if hasattr(ids, "tolist") else list(ids)
        return " ".join(f"w{t}" for t in ids if not (skip_special_tokens and t in self.special))


class BatchTokenizer(FakeTokenizer):
    calls = 0

    def batch_decode(self, rows, skip_special_tokens=True):
        BatchTokenizer.calls += 1
        return [self.decode(r, skip_special_tokens) for r in rows]


def _naive_trim(row, eos, pad, keep_eos=True):
    if eos in row:
        n = row.index(eos) + (1 if keep_eos else 0)
        return row[:n] + [pad] * (len(row) - n), n
    return list(row), len(row)


@pytest.mark.parametrize("dtype", [torch.int32, torch.int64])
@pytest.mark.parametrize("keep_eos", [True, False])
def test_trim_at_eos_matches_a_python_loop(dtype, keep_eos):
    rng = random.Random(0)
    rows = [[rng.choice([2, 3, 4, 5, 6, 7, 8, 9]) for _ in range(12)] for _ in range(64)]
    rows[0] = [3] * 12                    # no EOS
    rows[1] = [2] + [5] * 11              # EOS first
    rows[2] = [5] * 11 + [2]              # EOS last
    tok = torch.tensor(rows, dtype=dtype)
    out, lens = cd.trim_at_eos(tok, eos_token_id=2, pad_token_id=0, keep_eos=keep_eos)
    assert out.dtype == dtype and lens.dtype == torch.int64
    for r, o, n in zip(rows, out.tolist(), lens.tolist()):
        want, wn = _naive_trim(r, 2, 0, keep_eos)
        assert o == want and n == wn


def test_trim_rejects_wrong_rank():
    with pytest.raises(ValueError):
        cd.trim_at_eos(torch.zeros(5, dtype=torch.int32), 2)


def test_decode_captions_equals_reference_loop_after_eos_garbage_is_removed():
    rng = random.Random(1)
    rows = [[1] + [rng.choice([2, 3, 4, 5, 6, 7]) for _ in range(9)] for _ in range(32)]
    tok = torch.tensor(rows, dtype=torch.int32)
    tk = FakeTokenizer()
    got = cd.decode_captions(tok, tk)
    # the reference loop on the same ids, cut at the first EOS (what a sane caller wants from it)
    want = [tk.decode(_naive_trim(r, 2, 0)[0], skip_special_tokens=True) for r in rows]
    assert got == want
    # a tokenizer with batch_decode is called once for the whole batch
    BatchTokenizer.calls = 0
    assert cd.decode_captions(tok, BatchTokenizer()) == want and BatchTokenizer.calls == 1
    # rows that never emit EOS decode exactly like the reference's per-caption call
    noeos = torch.tensor([[1, 3, 4, 5], [1, 6, 7, 3]], dtype=torch.int64)
    assert cd.decode_captions(noeos, tk) == [tk.decode(r) for r in noeos]


def test_to_token_lists_and_results_json(tmp_path):
    tok = torch.tensor([[1, 5, 6, 2, 0], [1, 7, 2, 0, 0]], dtype=torch.int32)
    _, lens = cd.trim_at_eos(tok, 2)
    assert cd.to_token_lists(tok, lens, skip_ids=(1, 2)) == [[5, 6], [7]]
    res = cd.coco_results(torch.tensor([42, 7]), ["a cat", "a dog"])
    assert res == [{"image_id": 42, "caption": "a cat"}, {"image_id": 7, "caption": "a dog"}]
    p = cd.write_results_json(str(tmp_path / "results.json"), [42, 7], ["a cat", "a dog"])
    assert json.load(open(p)) == res
    with pytest.raises(ValueError):
        cd.coco_results([1], ["a", "b"])
