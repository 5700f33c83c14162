"""GPU suite: encoder -> decoder feature hand-off (include/capdec.h "feature hand-off").

capdec_ingest_features must be EXACTLY the reference's eager hand-off ops followed by what the decode prologue does:
  models/encoder.py:15     out.permute(0, 2, 3, 1)                 (NCHW trunk output -> [B,14,14,2048])
  encoders.py:122,213      last_hidden_state[:, 1:, :]             (CLS dropped)
  .float()                 widening of an autocast encoder's bf16 / fp16 output
so decoding a tile set equals decoding the fp32 [B,L,D] tensor those ops produce -- bit for bit, because the tiles hold
the same operand bytes the prologue would have packed.  The p24 source format (16 significant bits) is checked against
its host restatement (tests/test_p24_format.py) and through the decode at the mode's tolerance."""
import numpy as np
import pytest
import torch

import capdec_b200 as cd
from capdec_b200 import engine as eng_mod
from capdec_b200._capi import CapdecError
from oracle import beam as obeam, legacy as olegacy, lstm as olstm
from tests.helpers import gpt2_decoder, legacy_weights, lstm_decoder, lstm_inputs
from tests.test_p24_format import decode as p24_decode, encode as p24_encode

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def _trunk(B, seed, dtype=torch.float32):
    """a resnet-trunk-shaped output [B,2048,14,14] (post-ReLU, like models/encoder.py:13)"""
    g = torch.Generator().manual_seed(seed)
    return torch.relu(torch.randn(B, 2048, 14, 14, generator=g)).to(dtype)


def _same(a, b):
    return torch.equal(a["tokens"], b["tokens"]) and torch.equal(a["scores"], b["scores"]) and torch.equal(a["lengths"], b["lengths"])


@pytest.mark.parametrize("precision", ["bf16x3", "tf32x3", "fp32"])
@pytest.mark.parametrize("src_dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_ingest_nchw_equals_permute_then_decode(cuda, precision, src_dtype):
    """[B,2048,14,14] in the encoder's dtype -> tiles -> beam search == beam search on permute(0,2,3,1).float()."""
    B, k, T = 37, 3, 8
    m, _ = legacy_weights(1000, 0)
    m.precision = precision
    m = m.to(cuda)
    trunk = _trunk(B, 5, src_dtype).to(cuda)
    ref_feats = trunk.permute(0, 2, 3, 1).float().contiguous()          # models/encoder.py:15 + widening
    ref = m.beam_search(ref_feats, beam_size=k, max_length=T)
    tiles = m.ingest(trunk)                                             # layout "nchw"
    out = m.beam_search(tiles, beam_size=k, max_length=T)
    assert _same(out, ref)
    # the row-major form of the same hand-off
    tiles2 = m.ingest(trunk.permute(0, 2, 3, 1).contiguous(), layout="bld")
    assert _same(m.beam_search(tiles2, beam_size=k, max_length=T), ref)


def test_ingest_matches_oracle_on_handed_off_features(cuda):
    """the tile path against the CPU oracle run on the permuted fp32 tensor (the reference's own hand-off)."""
    B, k, T = 12, 3, 10
    m, sd = legacy_weights(2000, 1)
    m.precision = "bf16x3"
    trunk = _trunk(B, 6)
    enc = trunk.permute(0, 2, 3, 1).contiguous()
    ref = obeam.beam_search(olegacy.LegacyStepper(sd, enc, k), B, k, T)
    m = m.to(cuda)
    out = m.beam_search(m.ingest(trunk.to(cuda)), beam_size=k, max_length=T)
    same = (out["tokens"].cpu().long() == ref["sequences"]).all(dim=1)
    assert int((~same).sum()) <= 1, same
    assert torch.allclose(out["scores"].cpu()[same], ref["scores"][same], atol=1e-3)


def test_tile_planes_are_the_p24_encoding_and_mean(cuda):
    """white box: the bf16x3 tile set = {top-16-bit plane, byte plane, bf16 lo operand, region mean} of the transposed
    source, equal to the host restatement of the format."""
    B, L, D = 3, 196, 2048
    m, _ = legacy_weights(100, 0)
    m.precision = "bf16x3"
    m = m.to(cuda)
    trunk = _trunk(B, 7)
    tiles = m.ingest(trunk.to(cuda))
    x = trunk.permute(0, 2, 3, 1).reshape(B, L, D).contiguous()
    hi_ref, q_ref = p24_encode(x.numpy().reshape(-1))
    n = B * L * D
    buf = tiles.buf.cpu().numpy()
    hi = buf[: 2 * n].view(np.uint16)
    off = (2 * n + 255) // 256 * 256
    q = buf[off: off + n]
    assert np.array_equal(hi, hi_ref) and np.array_equal(q, q_ref)
    off = (off + n + 255) // 256 * 256
    lo = torch.from_numpy(buf[off: off + 2 * n].copy()).view(torch.bfloat16).float().reshape(B, L, D)
    rem = x - torch.from_numpy((hi_ref.astype(np.uint32) << 16).view(np.float32)).reshape(B, L, D)
    assert torch.equal(lo, rem.bfloat16().float())
    off = (off + 2 * n + 255) // 256 * 256
    mean = torch.from_numpy(buf[off: off + 4 * B * D].copy()).view(torch.float32).reshape(B, D)
    assert torch.allclose(mean, x.mean(dim=1), atol=1e-6)
    dec = torch.from_numpy(p24_decode(hi, q)).reshape(B, L, D)
    assert ((dec - x).abs() <= x.abs() * 2.0 ** -16 + 1e-38).all()


@pytest.mark.parametrize("kind,heads", [("soft", 1), ("multi_head", 8)])
def test_ingest_cls_drop_for_the_src_decoders(cuda, kind, heads):
    """ViT / CLIP last_hidden_state [B,1+L,H] (bf16): tiles == features[:, 1:, :].float() (encoders.py:122,213)."""
    B, L, H, T, k = 9, 49, 256, 9, 3
    m, sd = lstm_decoder(kind, H=H, layers=1, heads=heads, V=1500, seed=4)
    m.precision = "bf16x3"
    m = m.to(cuda)
    g = torch.Generator().manual_seed(8)
    hidden = torch.randn(B, 1 + L, H, generator=g).bfloat16()
    pooled = torch.randn(B, H, generator=g)
    feats = hidden[:, 1:, :].float().contiguous()
    eng = m._engine(cuda)
    tiles = eng.ingest_features(hidden.to(cuda), layout="cls_bld")
    ref = eng.decode_beam(feats.to(cuda), pooled.to(cuda), None, k, T)
    out = eng.decode_beam(tiles, pooled.to(cuda), None, k, T)
    assert _same(out, ref)
    # and against the oracle on the cropped tensor
    oref = obeam.beam_search(olstm.LSTMStepper(sd, feats, pooled, kind, 1, heads, k), B, k, T)
    same = (out["tokens"].cpu().long() == oref["sequences"]).all(dim=1)
    assert int((~same).sum()) <= 1


def test_host_entry_formats_match_device_path(cuda):
    """capdec_decode_beam_host_ex: fp32 / bf16 host features in [B,L,D] and NCHW layouts, and the 3-byte p24 host
    block, against the device entry point on the equivalent fp32 tensor."""
    B, k, T = 300, 5, 12
    m, _ = legacy_weights(3000, 0)
    m.precision = "bf16x3"
    m = m.to(cuda)
    eng = m._engine(cuda)
    trunk = _trunk(B, 9)
    bld = trunk.permute(0, 2, 3, 1).reshape(B, 196, 2048).contiguous()
    ref = eng.decode_beam(bld.to(cuda), None, None, k, T)

    def host(x, **kw):
        o = eng.decode_beam_host(x.contiguous().pin_memory(), None, k, T, chunk_images=128, **kw)
        return {n: o[n].to(cuda) for n in ("tokens", "scores", "lengths")}
    assert _same(host(bld), ref)                                            # fp32 [B,L,D]: the round-1 entry point
    assert _same(host(trunk, layout="nchw"), ref)                           # fp32 NCHW
    ref16 = eng.decode_beam(bld.bfloat16().float().to(cuda), None, None, k, T)
    assert _same(host(bld.bfloat16()), ref16)                               # bf16 [B,L,D]: half the PCIe bytes
    assert _same(host(trunk.bfloat16(), layout="nchw"), ref16)
    # p24 block: 3 bytes per element; equals decoding the p24-decoded fp32 tensor
    packed = eng_mod.pack_p24_host(bld)
    hi, q = p24_encode(bld.numpy().reshape(-1))
    dec = torch.from_numpy(p24_decode(hi, q)).reshape(B, 196, 2048)
    ref24 = eng.decode_beam(dec.to(cuda), None, None, k, T)
    out24 = host(packed, dtype="p24", num_regions=196)
    assert _same(out24, ref24)
    # ... and stays within the mode's tolerance of the full-fp32 decode
    same = (out24["tokens"] == ref["tokens"]).all(dim=1)
    assert same.float().mean().item() >= 0.99
    assert (out24["scores"][same] - ref["scores"][same]).abs().max().item() < 2e-4


def test_host_entry_with_padding_mask(cuda):
    """mask argument of the host entry point (object-region features with padding, encoders.py:233-296)."""
    B, L, H, T, k = 70, 36, 256, 8, 3
    m, _ = lstm_decoder("multi_head", H=H, layers=1, heads=8, V=1200, seed=5)
    m.precision = "bf16x3"
    m = m.to(cuda)
    feats, pooled, mask = lstm_inputs(B, L, H, seed=11, ragged=True)
    eng = m._engine(cuda)
    ref = eng.decode_beam(feats.to(cuda), pooled.to(cuda), (~mask).to(cuda), k, T)
    out = eng.decode_beam_host(feats.pin_memory(), pooled.pin_memory(), k, T, chunk_images=32, mask_host=~mask)
    assert torch.equal(out["tokens"], ref["tokens"].cpu()) and torch.equal(out["scores"], ref["scores"].cpu())
    unmasked = eng.decode_beam(feats.to(cuda), pooled.to(cuda), None, k, T)
    assert not torch.equal(unmasked["tokens"], ref["tokens"])               # the mask does change the captions


def test_ingest_errors_are_loud(cuda):
    m, _ = legacy_weights(100, 0)
    m = m.to(cuda)
    eng = m._engine(cuda)
    with pytest.raises(ValueError, match="needs"):
        eng.ingest_features(torch.zeros(2, 2048, 13, 14, device=cuda), layout="nchw", num_regions=196)
    with pytest.raises(CapdecError, match="p24 sources"):
        eng.ingest_features(torch.zeros(2, 3 * 196 * 2048, dtype=torch.uint8, device=cuda), layout="nchw", dtype="p24", num_regions=196)
    g, _ = gpt2_decoder()
    with pytest.raises(CapdecError, match="pooled_features only"):
        g.to(cuda)._engine(cuda).ingest_features(torch.zeros(2, 5, 64, device=cuda))
