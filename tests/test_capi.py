"""CPU suite: the C-ABI library loads, exports every symbol include/capdec.h declares, and refuses to
run without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import subprocess

import pytest
import torch

import capdec_b200 as cd
from capdec_b200 import _capi


def test_library_is_in_tree_and_exports_header():
    assert os.path.isfile(_capi.LIB_PATH) and "image-captioning-ml-project_b200/csrc" in _capi.LIB_PATH
    syms = _capi.declared_symbols()
    assert len(syms) >= 28 and {"capdec_decode_beam", "capdec_decode_beam_host", "capdec_decode_beam_host_ex", "capdec_forward_tokens",
                                 "capdec_ingest_features", "capdec_decode_beam_tiles", "capdec_trim_at_eos"} <= set(syms)
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(syms) <= exported, set(syms) - exported
    assert _capi.lib.capdec_version() == 200


def test_library_contains_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_config_struct_matches_header():
    text = open(_capi.HEADER_PATH).read()
    body = text[text.index("typedef struct {"):text.index("} capdec_config;")]
    names = []
    for line in body.splitlines():
        line = line.split("/*")[0].strip()
        if line.startswith(("int32_t", "float")):
            names += [n.strip() for n in line.split(None, 1)[1].rstrip(";").split(",")]
    assert names == [f[0] for f in _capi.Config._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    m = cd.Decoder(50)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.beam_search(torch.zeros(1, 196, 2048))
    cfg = _capi.Config(arch=0, attention=0, precision=0, vocab_size=50, hidden_dim=512, embed_dim=512, feature_dim=2048,
                       attention_dim=512, num_layers=1, num_heads=1, temperature=1.0, pad_token_id=0, bos_token_id=1,
                       eos_token_id=2)
    h = C.c_void_p()
    st = _capi.lib.capdec_create(C.byref(cfg), C.byref(h))
    assert st == -3 and b"no CPU fallback" in _capi.lib.capdec_last_error()


def test_factories_mirror_reference_errors():
    with pytest.raises(ValueError, match="Unsupported attention type"):
        cd.build_attention(cd.AttentionConfig(attention_type=cd.AttentionType.OBJECT, hidden_dim=32))

    class Bogus:
        decoder_type = "nope"
    with pytest.raises(ValueError, match="Unsupported decoder type"):
        cd.build_decoder(Bogus(), cd.AttentionConfig(hidden_dim=32), 10, 0, 1, 2)
    d = cd.build_decoder(cd.DecoderConfig(decoder_type=cd.DecoderType.LSTM, hidden_dim=32, num_layers=1),
                         cd.AttentionConfig(attention_type=cd.AttentionType.SOFT, hidden_dim=32), 10, 0, 1, 2)
    assert isinstance(d, cd.LSTMDecoder) and isinstance(d, cd.CaptionDecoder)
    assert {"embedding.weight", "lstm.weight_ih_l0", "attention.energy.weight", "output_layer.bias",
            "init_h.weight"} <= set(d.state_dict())
