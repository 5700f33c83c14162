"""GPU suite: the teacher-forced pass over given tokens (SURVEY.md section 8 a9 boundary + (f) rank 2).

`decoder(encoder_features, captions=ids)["logits"]` of the drop-in modules against the oracle restatements of the
reference's `forward` (oracle/teacher.py, pinned to the reference modules in tests/test_oracle_pin.py), `score_tokens`
against log_softmax + gather of those logits, and the reference trainer's sampling loop
(src/train/trainer.py:383-438, restated in oracle/sample.py::sample_captions_loop and pinned to the unmodified
CaptioningTrainer._sample_captions) running UNCHANGED on the drop-in's forward -- same tokens and log-probs as the
single-call CUDA rollout capdec_decode_sample on the same uniforms."""
import copy

import pytest
import torch

from oracle import gpt2 as ogpt, lstm as olstm, sample as osample, teacher as oteach, transformer as otr
from tests.helpers import gpt2_decoder, legacy_features, legacy_weights, lstm_decoder, lstm_inputs, transformer_decoder

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
LOGIT_TOL = {"fp32": 2e-4, "tf32x3": 1e-3, "bf16x3": 1e-3}


def _caps(B, T, V, seed, pad_rows=()):
    c = torch.randint(3, V, (B, T), generator=torch.Generator().manual_seed(seed))
    c[:, 0] = 1
    for r, t in pad_rows:
        c[r, t:] = 0
    return c


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("kind,heads,layers,ragged", [("soft", 1, 1, False), ("multi_head", 8, 2, True), ("aoa", 8, 1, False),
                                                      ("adaptive", 8, 1, False)])
def test_lstm_forward_captions_vs_oracle(cuda, kind, heads, layers, ragged, precision):
    B, T, H, L, V = 6, 9, 256, 49, 1500
    m, sd = lstm_decoder(kind, H=H, layers=layers, heads=heads, V=V, seed=3)
    m.precision = precision
    feats, pooled, mask = lstm_inputs(B, L, H, seed=5, ragged=ragged)
    caps = _caps(B, T, V, 6)
    ref_logits, ref_alpha = oteach.lstm_logits(sd, feats, pooled, kind, layers, heads, caps, None if mask is None else ~mask)
    ef = {"features": feats.to(cuda), "pooled_features": pooled.to(cuda)}
    if mask is not None:
        ef["attention_mask"] = mask.to(cuda)
    out = m.to(cuda)(ef, captions=caps.to(cuda))
    assert out["logits"].shape == (B, T, V) and out["attention_weights"].shape == (B, T, L)
    err = (out["logits"].cpu() - ref_logits).abs().max().item()
    print(f"[lstm forward {kind} {precision}] max |dlogit| {err:.2e}")
    assert err < LOGIT_TOL[precision]
    assert torch.allclose(out["attention_weights"].cpu(), ref_alpha, atol=1e-4 if precision == "fp32" else 1e-3)
    lp = m.score_tokens(ef, caps.to(cuda))
    assert torch.allclose(lp.cpu(), oteach.token_logprobs(ref_logits, caps), atol=1e-3)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_transformer_forward_captions_vs_oracle(cuda, precision):
    """pad tokens inside the captions are masked as keys (tgt_key_padding_mask); an attention_mask masks region keys."""
    B, T, H, L, V, layers, heads = 5, 10, 128, 49, 500, 2, 4
    m, sd = transformer_decoder(H=H, layers=layers, heads=heads, V=V, seed=4)
    m.precision = precision
    feats, _, mask = lstm_inputs(B, L, H, seed=7, ragged=True)
    caps = _caps(B, T, V, 8, pad_rows=((2, 6),))
    caps[1, 3] = 0
    mg = m.to(cuda)
    for region_mask in (None, mask):
        ref = oteach.transformer_logits(sd, feats, layers, heads, caps, 0, None if region_mask is None else ~region_mask)
        ef = {"features": feats.to(cuda)}
        if region_mask is not None:
            ef["attention_mask"] = region_mask.to(cuda)
        out = mg(ef, captions=caps.to(cuda))
        assert set(out) == {"logits"} and out["logits"].shape == (B, T, V)
        err = (out["logits"].cpu() - ref).abs().max().item()
        print(f"[transformer forward {precision} mask={region_mask is not None}] max |dlogit| {err:.2e}")
        assert err < LOGIT_TOL[precision]
    lp = mg.score_tokens({"features": feats.to(cuda)}, caps.to(cuda))
    ref = oteach.transformer_logits(sd, feats, layers, heads, caps, 0)
    assert torch.allclose(lp.cpu(), oteach.token_logprobs(ref, caps), atol=1e-3)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_gpt2_forward_captions_vs_hf(cuda, precision):
    B, T, H, layers, heads, V = 5, 9, 64, 2, 4, 300
    m, sd = gpt2_decoder(H=H, layers=layers, heads=heads, V=V)
    m.precision = precision
    hf = copy.deepcopy(m.model)
    pooled = torch.randn(B, H, generator=torch.Generator().manual_seed(9))
    caps = _caps(B, T, V, 10, pad_rows=((3, 5),))
    ref_logits, ref_loss = oteach.gpt2_logits(hf, sd, pooled, caps, 0)
    out = m.to(cuda)({"pooled_features": pooled.to(cuda)}, captions=caps.to(cuda))
    err = (out["logits"].cpu() - ref_logits).abs().max().item()
    print(f"[gpt2 forward {precision}] max |dlogit| {err:.2e}, loss {float(out['loss']):.5f} vs {float(ref_loss):.5f}")
    assert err < LOGIT_TOL[precision]
    assert abs(float(out["loss"]) - float(ref_loss)) < 1e-3


def test_legacy_score_tokens_vs_oracle(cuda):
    """re-scoring of sampled captions in one pass: rows of an image share its tiles (rows_per_image = 3)."""
    from oracle import legacy as olegacy
    B, k, T, V = 7, 3, 8, 900
    m, sd = legacy_weights(V, 2)
    enc = legacy_features(B, seed=3)
    caps = _caps(B * k, T, V, 11)
    st = olegacy.LegacyStepper(sd, enc, k)
    ref = torch.stack([st(caps[:, t]) for t in range(T)], dim=1)
    for precision in ("fp32", "bf16x3"):
        m.precision = precision
        lp = m.to(cuda).score_tokens(enc.to(cuda), caps.to(cuda), rows_per_image=k)
        assert lp.shape == (B * k, T - 1)
        assert torch.allclose(lp.cpu(), oteach.token_logprobs(ref, caps), atol=1e-3)


@pytest.mark.parametrize("arch", ["transformer", "gpt2", "lstm"])
def test_reference_sampling_loop_runs_unmodified_on_the_dropin(cuda, arch):
    """The trainer's rollout (trainer.py:413-436: decoder(encoder_features, captions=input_ids)["logits"][:, -1] ->
    softmax -> draw -> log_prob -> cat -> all-EOS break) driven on the drop-in's `forward`, with shared uniforms:
    tokens and log-probs equal the oracle loop over the reference modules' restatement AND the single-call CUDA rollout
    (capdec_decode_sample), up to draws that sit within 1e-5 of a CDF edge."""
    B, T = 6, 9
    u = torch.rand(B, T - 1, generator=torch.Generator().manual_seed(12))
    if arch == "transformer":
        H, layers, heads, V, L = 128, 2, 4, 500, 49
        m, sd = transformer_decoder(H=H, layers=layers, heads=heads, V=V, seed=5)
        feats, _, _ = lstm_inputs(B, L, H, seed=13)
        ef = {"features": feats.to(cuda)}
        ofwd = lambda ids: {"logits": oteach.transformer_logits(sd, feats, layers, heads, ids, 0)}
    elif arch == "gpt2":
        H, layers, heads, V = 64, 2, 4, 300
        m, sd = gpt2_decoder(H=H, layers=layers, heads=heads, V=V)
        hf = copy.deepcopy(m.model)
        pooled = torch.randn(B, H, generator=torch.Generator().manual_seed(14))
        ef = {"pooled_features": pooled.to(cuda)}
        ofwd = lambda ids: {"logits": oteach.gpt2_logits(hf, sd, pooled, ids, 0)[0]}
    else:
        H, V, L = 128, 800, 49
        m, sd = lstm_decoder("aoa", H=H, layers=1, heads=8, V=V, seed=6)
        feats, pooled, _ = lstm_inputs(B, L, H, seed=15)
        ef = {"features": feats.to(cuda), "pooled_features": pooled.to(cuda)}
        ofwd = lambda ids: {"logits": oteach.lstm_logits(sd, feats, pooled, "aoa", 1, 8, ids)[0]}
    mg = m.to(cuda)
    ids_o, lp_o = osample.sample_captions_loop(ofwd, B, T, u, bos_token_id=1, eos_token_id=2)
    ids_g, lp_g = osample.sample_captions_loop(lambda ids: mg(ef, captions=ids), B, T, u, bos_token_id=mg.bos_token_id,
                                               eos_token_id=mg.eos_token_id, device=cuda)
    ids_g, lp_g = ids_g.cpu(), lp_g.cpu()
    rows_ok = (ids_g == ids_o).all(dim=1)
    assert int((~rows_ok).sum()) <= 1, (ids_g, ids_o)              # a draw on a CDF edge may differ
    assert torch.allclose(lp_g[rows_ok], lp_o[rows_ok], atol=1e-3)
    # the fast path: one CUDA call for the whole rollout
    tok, info = mg.generate(ef, T, do_sample=True, num_samples=1, uniforms=u.to(cuda))
    tok, lp = tok.cpu(), info["log_probs"].cpu()
    clean = rows_ok & ~(ids_g[:, 1:-1] == 0).any(dim=1)            # a sampled pad token changes the prefix pass only (key mask)
    n = min(tok.shape[1], ids_g.shape[1])
    assert torch.equal(tok[clean][:, :n], ids_g[clean][:, :n])
    assert torch.allclose(lp[clean][:, : n - 1], lp_g[clean][:, : n - 1], atol=1e-3)
