"""CPU suite: host-side behaviour of the drop-in modules that needs no GPU -- module copies / pickles never carry the
ctypes engine handle, unsupported generation arguments fail loudly instead of being dropped (the reference forwards
**kwargs to transformers' generate, src/models/decoders.py:640-650), and the C-ABI host packer for the p24 feature format."""
import copy
import ctypes
import pickle

import numpy as np
import pytest
import torch

import capdec_b200 as cd
from capdec_b200 import engine as eng_mod
from tests.helpers import gpt2_decoder, legacy_weights, lstm_decoder, transformer_decoder
from tests.test_p24_format import decode as p24_decode, encode as p24_encode


@pytest.mark.parametrize("make", [lambda: legacy_weights(50, 0)[0], lambda: lstm_decoder("aoa", H=32, layers=1, heads=4, V=50)[0],
                                  lambda: transformer_decoder(H=32, layers=1, heads=4, V=50)[0],
                                  lambda: cd.build_attention(cd.AttentionConfig(attention_type=cd.AttentionType.SOFT, hidden_dim=32))])
def test_module_copies_drop_the_engine_handle(make):
    m = make()
    object.__setattr__(m, "_eng", ctypes.c_void_p(1234))        # stands in for a live Engine (unpicklable ctypes pointer)
    object.__setattr__(m, "_eng_sig", ("sig",))
    m2 = copy.deepcopy(m)
    assert "_eng" not in m2.__dict__ and "_eng_sig" not in m2.__dict__ and "_eng" in m.__dict__
    m3 = pickle.loads(pickle.dumps(m))
    assert "_eng" not in m3.__dict__
    for (k, a), (_, b) in zip(m.state_dict().items(), m3.state_dict().items()):
        assert torch.equal(a, b), k


def test_unsupported_generation_arguments_raise():
    g, _ = gpt2_decoder()
    ef = {"pooled_features": torch.zeros(2, 64)}
    with pytest.raises(NotImplementedError, match="repetition_penalty"):
        g.generate(ef, 8, repetition_penalty=1.3)
    with pytest.raises(NotImplementedError, match="early_stopping"):
        g.generate(ef, 8, early_stopping=True)
    with pytest.raises(TypeError, match="no_such_option"):
        g.generate(ef, 8, no_such_option=1)
    m, _ = lstm_decoder("soft", H=32, layers=1, heads=1, V=50)
    ef = {"features": torch.zeros(2, 5, 32), "pooled_features": torch.zeros(2, 32)}
    with pytest.raises(TypeError, match="min_length"):
        m.generate(ef, 8, num_beams=3, min_length=5)
    with pytest.raises(ValueError, match="start_token_id"):
        m.generate(ef, 8, num_beams=3, start_token_id=7)       # beam search starts from bos; only greedy honours start_token_id
    t, _ = transformer_decoder(H=32, layers=1, heads=4, V=50)
    with pytest.raises(TypeError, match="top_k"):
        t.generate({"features": torch.zeros(2, 5, 32)}, 8, top_k=5)
    # HF defaults passed explicitly are what the path implements: accepted up to the point where a GPU is needed
    with pytest.raises(RuntimeError, match="CUDA"):
        g.generate({"pooled_features": torch.zeros(2, 64)}, 8, early_stopping=False, repetition_penalty=1.0)


def test_pack_p24_host_matches_the_format_restatement():
    x = torch.randn(3, 7, 16, generator=torch.Generator().manual_seed(0)) * 5
    packed = eng_mod.pack_p24_host(x).numpy()
    n = 7 * 16
    for b in range(3):
        hi, q = p24_encode(x[b].numpy().reshape(-1))
        assert np.array_equal(packed[b, : 2 * n].view(np.uint16), hi) and np.array_equal(packed[b, 2 * n:], q)
        y = p24_decode(hi, q)
        assert np.abs(y - x[b].numpy().reshape(-1)).max() <= np.abs(x[b].numpy()).max() * 2.0 ** -16


def test_source_bytes():
    lib = cd._capi.lib
    assert lib.capdec_source_bytes(0, 0, 4, 196, 2048) == 4 * 196 * 2048 * 4          # [B,L,D] fp32
    assert lib.capdec_source_bytes(1, 1, 4, 196, 2048) == 4 * 196 * 2048 * 2          # NCHW bf16
    assert lib.capdec_source_bytes(2, 2, 4, 196, 768) == 4 * 197 * 768 * 2            # CLS + patches, fp16
    assert lib.capdec_source_bytes(0, 3, 4, 196, 2048) == 4 * 196 * 2048 * 3          # p24
