"""GPU suite, BASELINE.json's FULL sizes (configs[0..4]) through size-independent properties: the CPU oracle cannot
run these batch sizes in seconds, so each config is checked for
  * determinism (two runs are bit-identical),
  * image independence / batch-size invariance: a 64-image slice decoded alone equals the same rows of the full batch
    (this is what catches index arithmetic that only breaks at size, and any dependence of a row on its neighbours),
  * structural invariants (token ranges, BOS at position 0, lengths, scores < 0, sorted candidates).
Parity against the oracle for the same shapes at oracle-sized batches lives in test_gpu_parity.py."""
import pytest
import torch

from tests.helpers import gpt2_decoder, legacy_weights, lstm_decoder, transformer_decoder

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def _rand(shape, seed, device, relu=False):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(*shape, generator=g, device=device)
    return x.relu_() if relu else x


def _check_tokens(tok, V, bos=1):
    assert tok[:, 0].eq(bos).all() and (tok >= 0).all() and (tok < V).all()


def test_config1_legacy_beam3_64_images(cuda):
    """configs[0]: ResNet-101 features + LSTM + soft attention, beam 3, max_len 20, 64 images, vocab 10k."""
    B, k, T, V = 64, 3, 20, 10000
    m, _ = legacy_weights(V, 0)
    m.precision = "bf16x3"
    m = m.to(cuda)
    enc = _rand((B, 14, 14, 2048), 1, cuda, relu=True)
    a = m.beam_search(enc, beam_size=k, max_length=T, trace=True)
    b = m.beam_search(enc, beam_size=k, max_length=T)
    assert torch.equal(a["tokens"], b["tokens"]) and torch.equal(a["scores"], b["scores"])
    sub = m.beam_search(enc[16:32].contiguous(), beam_size=k, max_length=T)
    assert torch.equal(sub["tokens"], a["tokens"][16:32]) and torch.equal(sub["scores"], a["scores"][16:32])
    _check_tokens(a["tokens"], V)
    lp = a["top_logprob"]
    assert (lp[:, :, :-1] >= lp[:, :, 1:]).all()


@pytest.mark.parametrize("precision", ["bf16x3", "tf32x3"])
def test_config2_legacy_beam5_4096_images(cuda, precision):
    """configs[1] at the benchmark's per-GPU size: 4096 images x beam 5 = 20480 rows, 6.6 GB of features."""
    B, k, T, V = 4096, 5, 20, 10000
    m, _ = legacy_weights(V, 0)
    m.precision = precision
    m = m.to(cuda)
    enc = _rand((B, 196, 2048), 2, cuda, relu=True)
    a = m.beam_search(enc, beam_size=k, max_length=T)
    b = m.beam_search(enc, beam_size=k, max_length=T)
    assert torch.equal(a["tokens"], b["tokens"]) and torch.equal(a["scores"], b["scores"])
    for lo in (0, 2048, 4032):       # first / middle / last 64 images decoded alone
        sub = m.beam_search(enc[lo:lo + 64].contiguous(), beam_size=k, max_length=T)
        assert torch.equal(sub["tokens"], a["tokens"][lo:lo + 64]), lo
        assert torch.equal(sub["scores"], a["scores"][lo:lo + 64]), lo
        assert torch.equal(sub["lengths"], a["lengths"][lo:lo + 64]), lo
    _check_tokens(a["tokens"], V)
    assert (a["lengths"] >= 2).all() and (a["lengths"] <= T).all() and (a["scores"] < 0).all()
    # images are different, so (with random-init weights) their captions must not all collapse to one sequence
    assert len({tuple(r.tolist()) for r in a["tokens"][:256].cpu()}) > 8


def test_config2_src_lstm_soft_beam5_4096_images(cuda):
    """The src/ form of the same config (LSTMDecoder + SoftAttention over 196 x 512 projected features)."""
    B, k, T, V, H, L = 4096, 5, 20, 10000, 512, 196
    m, _ = lstm_decoder("soft", H=H, layers=1, heads=8, V=V)
    m.precision = "bf16x3"
    m = m.to(cuda)
    ef = {"features": _rand((B, L, H), 3, cuda), "pooled_features": _rand((B, H), 4, cuda)}
    ids, info = m.generate(ef, T, num_beams=k)
    ids2, _ = m.generate(ef, T, num_beams=k)
    assert torch.equal(ids, ids2)
    sub_ef = {k_: v[1000:1064].contiguous() for k_, v in ef.items()}
    sub, sinfo = m.generate(sub_ef, T, num_beams=k)
    n = min(sub.shape[1], ids.shape[1])
    assert torch.equal(sub[:, :n], ids[1000:1064, :n]) and torch.equal(sinfo["scores"], info["scores"][1000:1064])
    _check_tokens(ids, V)


def test_config3_transformer_beam3_2048_images(cuda):
    """configs[2]: ViT-B/16 features (196 x 768) + 6-layer transformer decoder (8 heads), beam 3, KV-cached, batch 2048."""
    B, k, T, V, H, L = 2048, 3, 20, 10000, 768, 196
    m, _ = transformer_decoder(H=H, layers=6, heads=8, V=V, max_length=50)
    m.precision = "bf16x3"
    m = m.to(cuda)
    ef = {"features": _rand((B, L, H), 5, cuda)}
    ids, info = m.generate(ef, T, num_beams=k)
    ids2, info2 = m.generate(ef, T, num_beams=k)
    assert torch.equal(ids, ids2) and torch.equal(info["scores"], info2["scores"])
    sub, sinfo = m.generate({"features": ef["features"][512:576].contiguous()}, T, num_beams=k)
    n = min(sub.shape[1], ids.shape[1])
    assert torch.equal(sub[:, :n], ids[512:576, :n]) and torch.equal(sinfo["scores"], info["scores"][512:576])
    _check_tokens(ids, V)
    assert (info["scores"] < 0).all()


def test_config4_gpt2_124m_bf16_beam5_1024_images(cuda):
    """configs[3]: CLIP pooled features + GPT-2 124M (12 layers, 12 heads, vocab 50257), beam 5, bf16, batch 1024."""
    B, k, T, V, H = 1024, 5, 20, 50257, 768
    m, _ = gpt2_decoder(H=H, layers=12, heads=12, V=V, max_length=64)
    m.precision = "bf16"
    m = m.to(cuda)
    ef = {"pooled_features": _rand((B, H), 6, cuda)}
    ids, info = m.generate(ef, T, num_beams=k, return_scores=True)
    ids2, info2 = m.generate(ef, T, num_beams=k, return_scores=True)
    assert torch.equal(ids, ids2) and torch.equal(info["scores"], info2["scores"])
    sub, sinfo = m.generate({"pooled_features": ef["pooled_features"][960:1024].contiguous()}, T, num_beams=k,
                            return_scores=True)
    n = min(sub.shape[1], ids.shape[1])
    assert torch.equal(sub[:, :n], ids[960:1024, :n]) and torch.equal(sinfo["scores"], info["scores"][960:1024])
    _check_tokens(ids, V)


def test_config5_scst_rollout_512_images(cuda):
    """configs[4]: SCST rollout on CLIP + GPT-2: 5 multinomial samples + 1 greedy row per image, batch 512 (3072 rows).
    Same uniforms => same draws; the greedy row equals greedy decoding; per-token log-probs are valid."""
    B, S, T, V, H = 512, 5, 20, 50257, 768
    m, _ = gpt2_decoder(H=H, layers=12, heads=12, V=V, max_length=64)
    m.precision = "bf16x3"
    m = m.to(cuda)
    ef = {"pooled_features": _rand((B, H), 7, cuda)}
    u = torch.rand(B * (S + 1), T - 1, generator=torch.Generator(device=cuda).manual_seed(8), device=cuda)
    tok, info = m.generate(ef, T, do_sample=True, num_samples=S, with_greedy=True, uniforms=u)
    tok2, info2 = m.generate(ef, T, do_sample=True, num_samples=S, with_greedy=True, uniforms=u)
    assert torch.equal(tok, tok2) and torch.equal(info["log_probs"], info2["log_probs"])
    assert tok.shape[0] == B * (S + 1)
    _check_tokens(tok, V)
    lp = info["log_probs"][:, : tok.shape[1] - 1]
    assert (lp <= 0).all() and torch.isfinite(lp).all()
    # sub-batch invariance: rows of images [128, 192)
    sl = slice(128 * (S + 1), 192 * (S + 1))
    sub, sinfo = m.generate({"pooled_features": ef["pooled_features"][128:192].contiguous()}, T, do_sample=True,
                            num_samples=S, with_greedy=True, uniforms=u[sl].contiguous())
    n = min(sub.shape[1], tok.shape[1])
    assert torch.equal(sub[:, :n], tok[sl, :n])
    # the 5 samples of an image differ from each other for (almost) every image with near-flat random-init logits
    rows = tok.view(B, S + 1, -1)
    distinct = (rows[:, 0] != rows[:, 1]).any(dim=1).float().mean().item()
    assert distinct > 0.9, distinct


@pytest.mark.parametrize("precision", ["bf16x3", "tf32x3", "bf16"])
def test_stream_k_gemm_keeps_batch_invariance(cuda, precision):
    """At 512 images (2560 rows) the gate GEMM has 80 tiles for 74 resident CTA pairs and runs as stream-K: tiles are cut
    at K-block boundaries between CTA groups, the later part CONTINUES the earlier part's accumulator (copied back into
    TMEM), so every output is summed in the same order as by one group.  The captions must therefore equal, bit for bit,
    those of the same images decoded in batches small enough for whole-tile scheduling -- the property the N-GPU sharded
    decode relies on (bench.py asserts gathered == single-GPU)."""
    B, k, T, V = 512, 5, 20, 10000
    m, _ = legacy_weights(V, 0)
    m.precision = precision
    m = m.to(cuda)
    enc = _rand((B, 196, 2048), 9, cuda, relu=True)
    full = m.beam_search(enc, beam_size=k, max_length=T, trace=True)
    for lo, hi in ((0, 256), (256, 512), (100, 164)):
        part = m.beam_search(enc[lo:hi].contiguous(), beam_size=k, max_length=T, trace=True)
        assert torch.equal(part["tokens"], full["tokens"][lo:hi]), (precision, lo)
        assert torch.equal(part["scores"], full["scores"][lo:hi]), (precision, lo)
        assert torch.equal(part["top_logprob"], full["top_logprob"][:, lo:hi]), (precision, lo)
    # and at 1024 / 2048 images (160 / 320 tiles: 3 / 5 waves on 74 pairs, also stream-K)
    big = _rand((2048, 196, 2048), 10, cuda, relu=True)
    a = m.beam_search(big, beam_size=k, max_length=T)
    b = m.beam_search(big[:1024].contiguous(), beam_size=k, max_length=T)
    c = m.beam_search(big[1024:1280].contiguous(), beam_size=k, max_length=T)
    assert torch.equal(b["tokens"], a["tokens"][:1024]) and torch.equal(b["scores"], a["scores"][:1024])
    assert torch.equal(c["tokens"], a["tokens"][1024:1280]) and torch.equal(c["scores"], a["scores"][1024:1280])
