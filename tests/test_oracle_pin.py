"""CPU suite: pin the oracle (a) against the reference's own modules when /root/reference is present
(the build container), (b) against the committed golden vectors those modules produced, (c) the beam
driver against transformers' own generate(num_beams=k)."""
import glob
import os

import pytest
import torch

from oracle import attention as oatt, beam as obeam, gpt2 as ogpt, legacy as olegacy, lstm as olstm, refshim, sample as osample, transformer as otr
from tests.helpers import GOLDEN, gpt2_decoder, legacy_features, legacy_weights, lstm_decoder, lstm_inputs, transformer_decoder

needs_ref = pytest.mark.skipif(not refshim.reference_available(), reason="/root/reference not present on this box")
torch.set_grad_enabled(False)


# ---------------------------------------------------------------- (a) against the reference itself
@needs_ref
def test_dropin_init_equals_reference_init():
    ns = refshim.load_reference()
    torch.manual_seed(0)
    ref = ns.legacy.Decoder(300, False, "cpu")
    _, sd = legacy_weights(300, 0)
    assert set(sd) == set(ref.state_dict())
    assert all(torch.equal(sd[k], v) for k, v in ref.state_dict().items())
    C = ns.config
    for kind in ("soft", "multi_head", "adaptive", "aoa"):
        for heads in (1, 4):
            torch.manual_seed(0)
            r = ns.decoders.LSTMDecoder(
                C.DecoderConfig(decoder_type=C.DecoderType.LSTM, hidden_dim=32, num_layers=2),
                C.AttentionConfig(attention_type=C.AttentionType(kind), num_heads=heads, hidden_dim=32),
                vocab_size=50, pad_token_id=0)
            _, sd2 = lstm_decoder(kind, H=32, layers=2, heads=heads, V=50)
            assert set(sd2) == set(r.state_dict())
            assert all(torch.equal(sd2[k], v) for k, v in r.state_dict().items())


@needs_ref
def test_legacy_restatement_matches_reference_forward():
    ns = refshim.load_reference()
    torch.manual_seed(3)
    ref = ns.legacy.Decoder(400, False, "cpu").eval()
    sd = {k: v.detach() for k, v in ref.state_dict().items()}
    enc = legacy_features(4, seed=5)
    g = torch.Generator().manual_seed(9)
    lens = [7, 7, 4, 2]
    caps = torch.randint(0, 400, (4, 7), generator=g)
    preds, _, dec_len, alphas = ref(enc, caps, lens)
    p2, a2, d2 = olegacy.forward_teacher_forced(sd, enc, caps, lens)
    assert dec_len == d2
    assert torch.equal(preds, p2) and torch.equal(alphas, a2)


@needs_ref
@pytest.mark.parametrize("kind,heads", [("soft", 8), ("multi_head", 4), ("aoa", 4), ("aoa", 1), ("adaptive", 4), ("adaptive", 1)])
@pytest.mark.parametrize("ragged", [False, True])
def test_lstm_restatement_matches_reference_generate(kind, heads, ragged):
    ns = refshim.load_reference()
    C = ns.config
    torch.manual_seed(1)
    H, layers, V, L, B, T = 64, 2, 300, 21, 5, 10
    ref = ns.decoders.LSTMDecoder(
        C.DecoderConfig(decoder_type=C.DecoderType.LSTM, hidden_dim=H, num_layers=layers),
        C.AttentionConfig(attention_type=C.AttentionType(kind), num_heads=heads, hidden_dim=H),
        vocab_size=V, pad_token_id=0).eval()
    sd = {k: v.detach() for k, v in ref.state_dict().items()}
    feats, pooled, mask = lstm_inputs(B, L, H, seed=2, ragged=ragged)
    ef = {"features": feats, "pooled_features": pooled}
    if mask is not None:
        ef["attention_mask"] = mask
    ids, info = ref.generate(ef, T)
    ids2, al2 = olstm.generate_greedy(sd, feats, pooled, kind, layers, T, num_heads=heads,
                                      mask=None if mask is None else ~mask)
    assert torch.equal(ids, ids2)
    assert torch.allclose(info["attention_weights"], al2, atol=1e-6)


@needs_ref
@pytest.mark.parametrize("kind,heads", [("soft", 1), ("multi_head", 8), ("aoa", 8), ("adaptive", 8)])
def test_attention_restatement_matches_reference_module(kind, heads):
    ns = refshim.load_reference()
    C = ns.config
    torch.manual_seed(4)
    H, L, B = 64, 17, 6
    mod = ns.attention.build_attention(C.AttentionConfig(attention_type=C.AttentionType(kind), num_heads=heads,
                                                         hidden_dim=H, temperature=1.7)).eval()
    sd = {"attention." + k: v.detach() for k, v in mod.state_dict().items()}
    g = torch.Generator().manual_seed(6)
    q, feats = torch.randn(B, H, generator=g), torch.randn(B, L, H, generator=g)
    mem, cell = torch.randn(B, H, generator=g), torch.randn(B, H, generator=g)
    pad = torch.zeros(B, L, dtype=torch.bool)
    pad[2, 9:] = True
    ctx, w = mod(q, feats, feats, pad, memory_state=mem, cell_state=cell)
    ctx2, w2 = oatt.attend(kind, sd, "attention.", q, feats, heads, None, pad, 1.7, mem, cell)
    assert torch.allclose(ctx, ctx2, atol=1e-6) and torch.allclose(w, w2, atol=1e-6)


@needs_ref
def test_transformer_restatement_matches_reference_generate():
    ns = refshim.load_reference()
    C = ns.config
    torch.manual_seed(5)
    H, layers, heads, V = 64, 2, 4, 120
    ref = ns.decoders.TransformerDecoder(
        C.DecoderConfig(decoder_type=C.DecoderType.TRANSFORMER, hidden_dim=H, num_layers=layers, num_heads=heads,
                        max_length=50), vocab_size=V, pad_token_id=0, bos_token_id=1, eos_token_id=2).eval()
    sd = {k: v.detach() for k, v in ref.state_dict().items()}
    feats, _, _ = lstm_inputs(4, 19, H, seed=8)
    ids, info = ref.generate({"features": feats}, 11)
    assert info == {}
    assert torch.equal(ids, otr.generate_greedy(sd, feats, layers, heads, 11))
    # the all-rows-EOS break (decoders.py:490): bias EOS so every row emits it at once
    sd2 = dict(sd)
    sd2["output_layer.bias"] = sd["output_layer.bias"].clone()
    sd2["output_layer.bias"][2] = 50.0
    ref.load_state_dict(sd2)
    ids, _ = ref.generate({"features": feats}, 11)
    assert ids.shape == (4, 2) and torch.equal(ids, otr.generate_greedy(sd2, feats, layers, heads, 11))


# ---------------------------------------------------------------- (b) against the committed goldens
def test_golden_legacy_teacher_forced():
    gd = torch.load(os.path.join(GOLDEN, "legacy_teacher.pt"))
    _, sd = legacy_weights(gd["vocab"], gd["seed"])
    enc = legacy_features(gd["B"], gd["feat_seed"])
    preds, alphas, _ = olegacy.forward_teacher_forced(sd, enc, gd["caps"], gd["lens"])
    assert torch.allclose(preds[:, :, ::97], gd["preds_sub"], atol=1e-5)
    assert torch.equal(preds.argmax(-1), gd["argmax"])
    assert torch.allclose(alphas, gd["alphas"], atol=1e-6)


def test_golden_legacy_beam3():
    gd = torch.load(os.path.join(GOLDEN, "legacy_beam3.pt"))
    _, sd = legacy_weights(gd["vocab"], gd["seed"])
    B = 3  # first images only: keeps the CPU suite short; the GPU suite checks all of them
    enc = legacy_features(gd["B"], gd["feat_seed"])[:B]
    out = obeam.beam_search(olegacy.LegacyStepper(sd, enc, gd["k"]), B, gd["k"], gd["T"], record_steps=True)
    assert torch.equal(out["sequences"], gd["sequences"][:B])
    assert torch.allclose(out["scores"], gd["scores"][:B], atol=1e-4)
    lp = torch.stack([s["top_lp"] for s in out["steps"]])
    live = gd["top_lp"][:, :B] > -1e8
    assert torch.allclose(lp[live], gd["top_lp"][:, :B][live], atol=1e-4)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "lstm_greedy_*_H256_*.pt"))))
def test_golden_lstm_greedy(path):
    gd = torch.load(path)
    _, sd = lstm_decoder(gd["kind"], H=gd["H"], layers=gd["layers"], heads=gd["heads"], V=gd["vocab"], seed=gd["seed"])
    feats, pooled, mask = lstm_inputs(gd["B"], gd["L"], gd["H"], gd["feat_seed"], gd["ragged"])
    ids, al = olstm.generate_greedy(sd, feats, pooled, gd["kind"], gd["layers"], gd["T"], num_heads=gd["heads"],
                                    mask=None if mask is None else ~mask)
    assert torch.equal(ids, gd["ids"])
    assert torch.allclose(al, gd["attention_weights"], atol=1e-6)


def test_golden_transformer_greedy():
    gd = torch.load(os.path.join(GOLDEN, "transformer_greedy_H128_l2_h4_L49.pt"))
    _, sd = transformer_decoder(H=gd["H"], layers=gd["layers"], heads=gd["heads"], V=gd["vocab"], seed=gd["seed"])
    feats, _, _ = lstm_inputs(gd["B"], gd["L"], gd["H"], gd["feat_seed"])
    assert torch.equal(otr.generate_greedy(sd, feats, gd["layers"], gd["heads"], gd["T"]), gd["ids"])


# ---------------------------------------------------------------- (c) beam driver against transformers
class _HFStepper:
    """prefix-recompute stepper over a real HF GPT-2 (logits sharpened / EOS-biased to exercise finishing)."""

    def __init__(self, model, first_tokens, scale, eos_bias):
        self.m, self.first, self.scale, self.eos_bias = model, first_tokens, scale, eos_bias
        self.state = None
        self.vocab_size = model.config.vocab_size

    def reorder(self, idx):
        self.state = self.state[idx]

    def __call__(self, tokens):
        if self.state is None:
            tokens = self.first
        seqs = tokens[:, None] if self.state is None else torch.cat([self.state, tokens[:, None]], 1)
        self.state = seqs
        logits = self.m(input_ids=seqs).logits[:, -1] * self.scale
        logits[..., 2] += self.eos_bias
        return logits


@pytest.mark.parametrize("k", [3, 5])
@pytest.mark.parametrize("scale,eos_bias", [(1.0, 0.0), (1.0, 4.0), (30.0, 12.0)])
@pytest.mark.parametrize("lp", [1.0, 0.8, 0.0])
def test_beam_driver_matches_hf(k, scale, eos_bias, lp):
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(0)
    V, B, T = 97, 6, 12
    m = GPT2LMHeadModel(GPT2Config(vocab_size=V, n_positions=32, n_embd=32, n_layer=2, n_head=2, bos_token_id=1,
                                   eos_token_id=2, pad_token_id=0)).eval()
    ids = torch.arange(3, 3 + B)[:, None]
    orig = m.forward

    def fwd(*a, **kw):
        out = orig(*a, **kw)
        out.logits = out.logits * scale
        out.logits[..., 2] += eos_bias
        return out

    m.forward = fwd
    hf = m.generate(input_ids=ids, max_length=T, num_beams=k, do_sample=False, pad_token_id=0, bos_token_id=1,
                    eos_token_id=2, length_penalty=lp, use_cache=False, return_dict_in_generate=True,
                    output_scores=True, early_stopping=False)
    m.forward = orig
    out = obeam.beam_search(_HFStepper(m, ids.repeat_interleave(k, 0)[:, 0], scale, eos_bias), B, k, T,
                            bos_token_id=1, eos_token_id=2, pad_token_id=0, length_penalty=lp)
    seq = out["sequences"].clone()
    seq[:, 0] = ids[:, 0]
    seq = obeam.crop_like_hf(seq, out["lengths"])
    assert hf.sequences.shape == seq.shape and torch.equal(hf.sequences, seq)
    assert torch.allclose(hf.sequences_scores, out["scores"], atol=1e-5)


@needs_ref
def test_gpt2_dropin_init_equals_reference_init():
    ns = refshim.load_reference()
    C = ns.config
    torch.manual_seed(0)
    ref = ns.decoders.GPT2Decoder(
        C.DecoderConfig(decoder_type=C.DecoderType.GPT2, pretrained_model_name="", hidden_dim=64, num_layers=2,
                        num_heads=4, dropout=0.0, max_length=64), vocab_size=300, pad_token_id=0, bos_token_id=1,
        eos_token_id=2)
    _, sd = gpt2_decoder()
    assert set(sd) == set(ref.state_dict())
    assert all(torch.equal(sd[k], v) for k, v in ref.state_dict().items())


@pytest.mark.parametrize("k,lp", [(4, 1.0), (5, 0.8)])
def test_gpt2_oracle_stepper_matches_hf_generate(k, lp):
    """a8: the pinned beam driver over an HF GPT-2 stepper with the image prefix as past K == V reproduces
    transformers' own generate() for the computation GPT2Decoder.generate intends (decoders.py:619-656)."""
    m, sd = gpt2_decoder()
    pooled = torch.randn(3, 64, generator=torch.Generator().manual_seed(5))
    seq, sc = ogpt.hf_generate(m.model, sd, pooled, k, 12, length_penalty=lp)
    out = obeam.beam_search(ogpt.HFStepper(m.model, sd, pooled, k), 3, k, 12, length_penalty=lp)
    assert torch.equal(seq, obeam.crop_like_hf(out["sequences"], out["lengths"]))
    assert torch.allclose(sc, out["scores"], atol=1e-5)


def test_sample_rollout_inverse_cdf():
    _, sd = lstm_decoder("soft", H=32, layers=1, V=50)
    feats, pooled, _ = lstm_inputs(3, 7, 32)
    u = torch.rand(6, 9, generator=torch.Generator().manual_seed(3))
    st = olstm.LSTMStepper(sd, feats, pooled, "soft", 1, rows_per_image=2)
    ids, lps, edges = osample.sample_rollout(st, 6, 10, u)
    assert ids.shape[0] == 6 and ids[:, 0].eq(1).all() and lps.shape == (6, ids.shape[1] - 1)
    assert (lps <= 0).all() and (ids >= 0).all() and (ids < 50).all()
    # same uniforms => same draw; different uniforms => (almost surely) different draw
    st2 = olstm.LSTMStepper(sd, feats, pooled, "soft", 1, rows_per_image=2)
    ids2, _, _ = osample.sample_rollout(st2, 6, 10, u)
    assert torch.equal(ids, ids2)


# ---------------------------------------------------------------- teacher-forced forward + the trainer's sampling loop
@needs_ref
@pytest.mark.parametrize("kind,heads,ragged", [("soft", 1, False), ("multi_head", 4, True), ("aoa", 4, False)])
def test_teacher_lstm_logits_match_reference_forward(kind, heads, ragged):
    """oracle/teacher.py::lstm_logits == the reference LSTMDecoder.forward(captions=...) (decoders.py:137-234)."""
    from oracle import teacher as oteach
    ns = refshim.load_reference()
    C = ns.config
    torch.manual_seed(7)
    H, layers, V, L, B, T = 64, 2, 300, 21, 5, 9
    ref = ns.decoders.LSTMDecoder(
        C.DecoderConfig(decoder_type=C.DecoderType.LSTM, hidden_dim=H, num_layers=layers),
        C.AttentionConfig(attention_type=C.AttentionType(kind), num_heads=heads, hidden_dim=H),
        vocab_size=V, pad_token_id=0).eval()
    sd = {k: v.detach() for k, v in ref.state_dict().items()}
    feats, pooled, mask = lstm_inputs(B, L, H, seed=3, ragged=ragged)
    caps = torch.randint(0, V, (B, T), generator=torch.Generator().manual_seed(4))
    ef = {"features": feats, "pooled_features": pooled}
    if mask is not None:
        ef["attention_mask"] = mask
    out = ref(ef, captions=caps)
    logits, alphas = oteach.lstm_logits(sd, feats, pooled, kind, layers, heads, caps, None if mask is None else ~mask)
    assert torch.allclose(out["logits"], logits, atol=1e-5) and torch.allclose(out["attention_weights"], alphas, atol=1e-6)
    # (with caption_lengths the reference sorts captions / features but takes h0, c0 from the UNSORTED pooled features,
    #  decoders.py:157-171, so rows get another image's initial state; _sample_captions passes None and the drop-in
    #  ignores the argument rather than reproduce the mis-pairing)


@needs_ref
def test_teacher_transformer_logits_match_reference_forward():
    """oracle/teacher.py::transformer_logits == the reference TransformerDecoder.forward(captions=...) (decoders.py:377-438),
    pad tokens inside the captions included (tgt_key_padding_mask)."""
    from oracle import teacher as oteach
    ns = refshim.load_reference()
    C = ns.config
    torch.manual_seed(8)
    H, layers, heads, V, B, T = 64, 2, 4, 120, 4, 8
    ref = ns.decoders.TransformerDecoder(
        C.DecoderConfig(decoder_type=C.DecoderType.TRANSFORMER, hidden_dim=H, num_layers=layers, num_heads=heads,
                        max_length=50), vocab_size=V, pad_token_id=0, bos_token_id=1, eos_token_id=2).eval()
    sd = {k: v.detach() for k, v in ref.state_dict().items()}
    feats, _, _ = lstm_inputs(B, 19, H, seed=9)
    caps = torch.randint(1, V, (B, T), generator=torch.Generator().manual_seed(5))
    caps[:, 0] = 1
    caps[1, 3] = 0          # a pad token in the middle of a prefix: masked as a key for every later position
    caps[2, 6:] = 0
    out = ref({"features": feats}, captions=caps)
    logits = oteach.transformer_logits(sd, feats, layers, heads, caps, pad_token_id=0)
    assert torch.allclose(out["logits"], logits, atol=1e-5), (out["logits"] - logits).abs().max()
    # the KV-cached stepper (what the CUDA path is built like) agrees wherever no pad key is involved
    st = otr.TransformerStepper(sd, feats, layers, heads, 1)
    step_logits = torch.stack([st(caps[:, t]) for t in range(T)], dim=1)
    assert torch.allclose(step_logits[0], logits[0], atol=1e-4) and torch.allclose(step_logits[3], logits[3], atol=1e-4)
    assert torch.allclose(step_logits[1, :3], logits[1, :3], atol=1e-4)
    assert not torch.allclose(step_logits[1, 4:], logits[1, 4:], atol=1e-4)      # the pad key does matter


@needs_ref
def test_sample_loop_restatement_matches_reference_trainer():
    """oracle/sample.py::sample_captions_loop (and the step-wise sample_rollout) == the UNMODIFIED
    CaptioningTrainer._sample_captions (trainer.py:383-438) when torch.distributions.Categorical is patched to draw by
    inverse CDF from the same uniforms."""
    import types
    rt = refshim.load_reference_trainer()
    ns = refshim.load_reference()
    C = ns.config
    torch.manual_seed(9)
    H, layers, heads, V, B, T = 64, 2, 4, 90, 6, 9
    dec = ns.decoders.TransformerDecoder(
        C.DecoderConfig(decoder_type=C.DecoderType.TRANSFORMER, hidden_dim=H, num_layers=layers, num_heads=heads,
                        max_length=50), vocab_size=V, pad_token_id=0, bos_token_id=1, eos_token_id=2).eval()
    sd = {k: v.detach() for k, v in dec.state_dict().items()}
    feats, _, _ = lstm_inputs(B, 19, H, seed=10)
    u = torch.rand(B, T - 1, generator=torch.Generator().manual_seed(11))

    class FakeCategorical:                       # consumes column `step` of the shared uniforms, in call order
        step = 0

        def __init__(self, probs):
            self.probs = probs

        def sample(self):
            cdf = self.probs.double().cumsum(-1)
            tok = (cdf <= u[:, FakeCategorical.step:FakeCategorical.step + 1].double()).sum(1).clamp(max=self.probs.size(1) - 1)
            FakeCategorical.step += 1
            return tok

        def log_prob(self, tok):
            return torch.log(self.probs.gather(1, tok[:, None]).squeeze(1))

    fake_self = types.SimpleNamespace(
        config=types.SimpleNamespace(inference=types.SimpleNamespace(max_length=T)),
        model=types.SimpleNamespace(encoder=lambda images: {"features": feats}, decoder=dec))
    orig = torch.distributions.Categorical
    torch.distributions.Categorical = FakeCategorical
    try:
        ids_ref, lp_ref = rt.CaptioningTrainer._sample_captions(fake_self, torch.zeros(B, 3, 8, 8))
    finally:
        torch.distributions.Categorical = orig
    ids2, lp2 = osample.sample_captions_loop(lambda ids: dec({"features": feats}, captions=ids), B, T, u)
    assert torch.equal(ids_ref, ids2) and torch.allclose(lp_ref, lp2, atol=1e-6)
    # step-wise rollout over the cached stepper: same tokens unless a sampled pad token (id 0) sits in a prefix
    ids3, lp3, _ = osample.sample_rollout(otr.TransformerStepper(sd, feats, layers, heads, 1), B, T, u)
    clean = ~(ids_ref[:, 1:-1] == 0).any(dim=1)
    assert clean.any()
    assert torch.equal(ids_ref[clean], ids3[clean]) and torch.allclose(lp_ref[clean], lp3[clean], atol=1e-4)


@needs_ref
def test_feature_producers_with_padding_masks_run_through_the_oracle():
    """the reference's alternative feature producers -- ObjectRegionEncoder (36 x 2048 regions with a
    padding mask, encoders.py:233-296) and QFormer (32 x 768 queries, captioning_model.py:153-245) -- emit exactly the
    {features, pooled_features, attention_mask} contract the decode path consumes; the reference LSTMDecoder on their
    output equals the oracle restatement with key_padding_mask = ~attention_mask."""
    rt = refshim.load_reference_trainer()
    ns = refshim.load_reference()
    C = ns.config
    torch.manual_seed(12)
    H, B, L = 64, 4, 36
    enc = rt.ObjectRegionEncoder(C.EncoderConfig(feature_dim=H, use_object_features=True)).eval()
    g = torch.Generator().manual_seed(13)
    region_mask = torch.ones(B, L)
    for b, n in enumerate((36, 20, 7, 1)):
        region_mask[b, n:] = 0
    ef = enc({"region_features": torch.randn(B, L, 2048, generator=g), "region_boxes": torch.rand(B, L, 4, generator=g),
              "region_mask": region_mask})
    assert ef["features"].shape == (B, L, H) and ef["pooled_features"].shape == (B, H)
    dec = ns.decoders.LSTMDecoder(
        C.DecoderConfig(decoder_type=C.DecoderType.LSTM, hidden_dim=H, num_layers=1),
        C.AttentionConfig(attention_type=C.AttentionType.MULTI_HEAD, num_heads=4, hidden_dim=H), vocab_size=200,
        pad_token_id=0).eval()
    sd = {k: v.detach() for k, v in dec.state_dict().items()}
    ef_bool = dict(ef, attention_mask=ef["attention_mask"].bool())        # the float mask trips `~`
    ids, info = dec.generate(ef_bool, 8)
    ids2, al2 = olstm.generate_greedy(sd, ef["features"], ef["pooled_features"], "multi_head", 1, 8, num_heads=4,
                                      mask=~ef_bool["attention_mask"])
    assert torch.equal(ids, ids2) and torch.allclose(info["attention_weights"], al2, atol=1e-6)
    assert float(info["attention_weights"][2, :, 7:].abs().max()) < 1e-12       # padded regions get no weight
    q = rt.QFormer(query_dim=H, vision_dim=H, num_queries=32, num_layers=1, num_heads=4, dropout=0.0).eval()
    out = q(ef["features"])
    assert out["queries"].shape == (B, 32, H)


def test_plain_torch_weights_equal_dropin_init():
    """oracle/weights.py (what bench.py's reference arm uses, so that arm never imports the product) == the seeded init of
    the drop-in Decoder == the reference Decoder's own init (test_dropin_init_equals_reference_init)."""
    from oracle.weights import legacy_state_dict
    a = legacy_state_dict(300, 3)
    _, b = legacy_weights(300, 3)
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
