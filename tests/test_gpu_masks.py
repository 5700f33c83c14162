"""GPU suite: padded region sets -- the feature shapes of the reference's alternative producers,
ObjectRegionEncoder (36 regions x 2048 -> hidden, padding mask, src/models/encoders.py:233-296) and QFormer (32 queries x 768,
src/models/captioning_model.py:153-245), through beam search against the CPU oracle: key_padding_mask = ~attention_mask
semantics of src/models/attention.py:97-100 (additive) / :183-186 (multi-head), and the transformer decoder's
memory_key_padding_mask (decoders.py:393-398).  Padded regions must get exactly zero attention weight."""
import pytest
import torch

from oracle import beam as obeam, lstm as olstm, sample as osample, transformer as otr
from tests.helpers import lstm_decoder, lstm_inputs, transformer_decoder

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def _object_region_inputs(B, L, H, seed):
    """what ObjectRegionEncoder emits: features [B,36,H], a ragged 0/1 region mask, pooled = masked mean (encoders.py:283-289)"""
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, L, H, generator=g)
    mask = torch.zeros(B, L)
    for b in range(B):
        mask[b, : 1 + int(torch.randint(0, L, (1,), generator=g))] = 1
    mask[0] = 1                      # one image with all 36 regions valid
    mask[1, 1:] = 0                  # ... and one with a single region
    m3 = mask.unsqueeze(-1)
    pooled = (feats * m3).sum(1) / (m3.expand_as(feats).sum(1) + 1e-10)
    return feats, pooled, mask


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("kind,heads", [("multi_head", 8), ("aoa", 8), ("aoa", 1), ("soft", 1)])
def test_object_region_features_beam_vs_oracle(cuda, kind, heads, precision):
    B, L, H, V, T, k = 12, 36, 256, 1500, 10, 3
    m, sd = lstm_decoder(kind, H=H, layers=1, heads=heads, V=V, seed=11)
    m.precision = precision
    feats, pooled, mask = _object_region_inputs(B, L, H, 31)
    kpm = ~mask.bool()
    ref = obeam.beam_search(olstm.LSTMStepper(sd, feats, pooled, kind, 1, heads, k, kpm), B, k, T, record_steps=True)
    ef = {"features": feats.to(cuda), "pooled_features": pooled.to(cuda), "attention_mask": mask.to(cuda)}   # float mask, as emitted
    mg = m.to(cuda)
    seq, info = mg.generate(ef, T, num_beams=k, trace=True)
    same = (torch.nn.functional.pad(seq, (0, T - seq.shape[1]), value=2).cpu() == ref["sequences"]).all(dim=1)
    ref_lp = torch.stack([s["top_lp"] for s in ref["steps"]])
    err0 = (info["top_logprob"].cpu()[0] - ref_lp[0]).abs().max().item()        # step 0: every image comparable
    print(f"[object regions {kind}/{heads} {precision}] identical beams {int(same.sum())}/{B}, step-0 max |dlogp| {err0:.2e}")
    assert int((~same).sum()) <= 1 and err0 < 1e-3
    if not bool(same.all()):            # a differing beam must be a near-tie for the oracle
        st1 = olstm.LSTMStepper(sd, feats, pooled, kind, 1, heads, 1, kpm)
        full = torch.nn.functional.pad(seq, (0, T - seq.shape[1]), value=2).cpu()
        gap = (ref["scores"] - osample.rescore(st1, full, info["lengths"].cpu()))[~same]
        assert float(gap.abs().max()) < 2e-3
    # greedy: padded regions receive exactly zero weight
    ids, ginfo = mg.generate(ef, T)
    w = ginfo["attention_weights"].cpu()
    assert float((w * (1 - mask)[:, None, :]).abs().max()) == 0.0
    assert torch.allclose(w.sum(-1), torch.ones(B, T), atol=1e-5)


def test_qformer_query_features_beam_vs_oracle(cuda):
    """Q-Former output: 32 queries x 768, no padding (captioning_model.py:81-88 resets the mask to ones)."""
    B, L, H, V, T, k = 6, 32, 768, 2000, 10, 3
    m, sd = lstm_decoder("multi_head", H=H, layers=1, heads=8, V=V, seed=12)
    m.precision = "bf16x3"
    feats, pooled, _ = lstm_inputs(B, L, H, seed=32)
    ref = obeam.beam_search(olstm.LSTMStepper(sd, feats, pooled, "multi_head", 1, 8, k), B, k, T)
    ef = {"features": feats.to(cuda), "pooled_features": pooled.to(cuda), "attention_mask": torch.ones(B, L, device=cuda)}
    seq, info = m.to(cuda).generate(ef, T, num_beams=k)
    same = (torch.nn.functional.pad(seq, (0, T - seq.shape[1]), value=2).cpu() == ref["sequences"]).all(dim=1)
    assert int((~same).sum()) <= 1
    assert torch.allclose(info["scores"].cpu()[same], ref["scores"][same], atol=1e-3)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_transformer_memory_key_padding_mask_beam_vs_oracle(cuda, precision):
    """cross-attention over a padded region set: masked keys get an additive -1e9 (decoders.py:393-398 intent)."""
    H, layers, heads, V, L, B, T, k = 128, 2, 4, 500, 36, 8, 9, 3
    m, sd = transformer_decoder(H=H, layers=layers, heads=heads, V=V, seed=13)
    m.precision = precision
    feats, _, mask = _object_region_inputs(B, L, H, 33)
    kpm = ~mask.bool()

    class MaskedStepper(otr.TransformerStepper):          # the oracle stepper with memory_key_padding_mask
        def __init__(self, rows):
            super().__init__(sd, feats, layers, heads, rows)
            self.mkpm = (kpm.float() * -1e9).repeat_interleave(rows, 0)

        def reorder(self, idx):
            super().reorder(idx)

        def __call__(self, tokens):
            import torch.nn as nn
            import torch.nn.functional as F
            self.prefix = tokens[:, None] if self.prefix is None else torch.cat([self.prefix, tokens[:, None]], 1)
            n = self.prefix.size(1)
            x = F.embedding(self.prefix, sd["embedding.weight"]) + F.embedding(torch.arange(n), sd["position_encoding.weight"])[None]
            out = self.dec(tgt=x, memory=self.mem, tgt_mask=nn.Transformer.generate_square_subsequent_mask(n).to(x.dtype),
                           memory_key_padding_mask=self.mkpm)
            return F.linear(out[:, -1], sd["output_layer.weight"], sd["output_layer.bias"])

    ref = obeam.beam_search(MaskedStepper(k), B, k, T, record_steps=True)
    seq, info = m.to(cuda).generate({"features": feats.to(cuda)}, T, num_beams=k, trace=True, memory_key_padding_mask=kpm.to(cuda))
    same = (torch.nn.functional.pad(seq, (0, T - seq.shape[1]), value=2).cpu() == ref["sequences"]).all(dim=1)
    ref_lp = torch.stack([s["top_lp"] for s in ref["steps"]])
    err0 = (info["top_logprob"].cpu()[0] - ref_lp[0]).abs().max().item()
    print(f"[transformer masked memory {precision}] identical beams {int(same.sum())}/{B}, step-0 max |dlogp| {err0:.2e}")
    assert int((~same).sum()) <= 1 and err0 < 1e-3
    unmasked, _ = m.generate({"features": feats.to(cuda)}, T, num_beams=k)
    assert not torch.equal(torch.nn.functional.pad(unmasked, (0, T - unmasked.shape[1]), value=2), torch.nn.functional.pad(seq, (0, T - seq.shape[1]), value=2))
