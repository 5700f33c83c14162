"""GPU suite: token -> text boundary on device tensors produced by the decode entry points.

The reference turns captions into text one row at a time -- `[tokenizer.decode(caption, skip_special_tokens=True) for
caption in gen_captions]` (src/train/trainer.py:546-547, src/evaluate/metrics.py:322-323) -- and builds
`{"image_id", "caption"}` records (metrics.py:326-336).  Here the same outputs come from capdec_decode_beam's device
tensors through capdec_trim_at_eos (CUDA) + one device->host copy, and must equal that per-caption loop."""
import pytest
import torch

import capdec_b200 as cd
from capdec_b200 import engine as eng_mod
from tests.helpers import legacy_features, legacy_weights, transformer_decoder, lstm_inputs

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


class ToyTokenizer:
    """id -> word; special ids 0 (pad), 1 (bos), 2 (eos) are dropped by skip_special_tokens like a HF tokenizer"""
    eos_token_id, pad_token_id, bos_token_id = 2, 0, 1

    def decode(self, ids, skip_special_tokens=True):
        ids = [int(i) for i in (ids.tolist() if hasattr(ids, "tolist") else ids)]
        return " ".join(f"w{i}" for i in ids if not (skip_special_tokens and i in (0, 1, 2)))

    def batch_decode(self, rows, skip_special_tokens=True):
        return [self.decode(r, skip_special_tokens) for r in rows]


def _reference_loop(captions, tok):
    """src/train/trainer.py:546-547 / metrics.py:322-323, one caption (and one device sync) at a time; the reference
    relies on everything after EOS being special tokens, which holds for beam output (HF fills with pad/eos)."""
    return [tok.decode(caption, skip_special_tokens=True) for caption in captions]


def test_beam_output_to_text_and_coco_records(cuda):
    B, k, T, V = 40, 3, 12, 600
    m, sd = legacy_weights(V, 5)
    sd["fc.weight"] *= 6.0
    sd["fc.weight"][2] *= 7.0           # EOS wins often: captions of mixed lengths
    m.load_state_dict(sd)
    out = m.to(cuda).beam_search(legacy_features(B, seed=7).to(cuda), beam_size=k, max_length=T)
    tok = ToyTokenizer()
    assert out["tokens"].is_cuda and out["tokens"].dtype == torch.int32
    assert int((out["lengths"] < T).sum()) >= 3
    texts = cd.decode_captions(out["tokens"], tok, lengths=out["lengths"])          # no trimming pass needed
    assert texts == _reference_loop(out["tokens"], tok)
    assert cd.decode_captions(out["tokens"], tok) == texts                          # CUDA trim path, same text
    ids = torch.arange(1000, 1000 + B)
    recs = cd.coco_results(ids, texts)
    assert recs == [{"image_id": int(i), "caption": t} for i, t in zip(ids.tolist(), _reference_loop(out["tokens"], tok))]


def test_trim_kernel_on_greedy_output(cuda):
    """greedy decodes keep writing argmax tokens after EOS (decoders.py:481-491): the CUDA trim equals the torch
    formulation and cuts the text at the first EOS."""
    B, T, H, V = 33, 14, 128, 40
    m, sd = transformer_decoder(H=H, layers=2, heads=4, V=V, seed=9)
    feats, _, _ = lstm_inputs(B, 49, H, seed=8)
    ids, _ = m.to(cuda).generate({"features": feats.to(cuda)}, T)                  # int64 on the device
    assert bool((ids[:, 1:] == 2).any()), "vocabulary of 40: some row must emit EOS"
    trimmed, lengths = cd.trim_at_eos(ids, 2, 0)
    ref_t, ref_l = cd.trim_at_eos(ids.cpu(), 2, 0)                                  # host formulation (torch ops)
    assert trimmed.is_cuda and torch.equal(trimmed.cpu(), ref_t) and torch.equal(lengths.cpu(), ref_l)
    t2, l2 = eng_mod.trim_at_eos_device(ids.int(), 2, 0, keep_eos=False)
    r2, rl2 = cd.trim_at_eos(ids.cpu(), 2, 0, keep_eos=False)
    assert torch.equal(t2.cpu().long(), r2) and torch.equal(l2.cpu().long(), rl2)
    tok = ToyTokenizer()
    texts = cd.decode_captions(ids, tok)
    for row, text in zip(ids.cpu().tolist(), texts):
        cut = row[: row.index(2) + 1] if 2 in row[1:] else row
        assert text == tok.decode(cut)
