"""ctypes binding of libcapdec.so (include/capdec.h).

The library is the product: importing this module fails loudly when the shared object is missing
or does not export every symbol the header declares.  There is no Python / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcapdec.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "capdec.h")

ARCH_LEGACY_SAT, ARCH_LSTM, ARCH_TRANSFORMER, ARCH_GPT2 = 0, 1, 2, 3
ATT = {"soft": 0, "multi_head": 1, "adaptive": 2, "aoa": 3}
PREC = {"fp32": 0, "tf32x3": 1, "bf16": 2, "tf32": 3, "bf16x3": 4}
LAYOUT = {"bld": 0, "bdl": 1, "nchw": 1, "cls_bld": 2}          # capdec_layout
DTYPE = {"f32": 0, "bf16": 1, "f16": 2, "p24": 3}               # capdec_dtype


class CapdecError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"capdec error {status}: {message}")
        self.status = status


class Config(C.Structure):
    _fields_ = [
        ("arch", C.c_int32), ("attention", C.c_int32), ("precision", C.c_int32), ("vocab_size", C.c_int32),
        ("hidden_dim", C.c_int32), ("embed_dim", C.c_int32), ("feature_dim", C.c_int32),
        ("attention_dim", C.c_int32), ("num_layers", C.c_int32), ("num_heads", C.c_int32),
        ("temperature", C.c_float), ("pad_token_id", C.c_int32), ("bos_token_id", C.c_int32),
        ("eos_token_id", C.c_int32),
    ]


def declared_symbols(header_path: str = HEADER_PATH):
    """Names of every function include/capdec.h declares (used by the CPU-side export test)."""
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(capdec_[a-z0-9_]+)\s*\(", text)))


def _load() -> C.CDLL:
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no fallback implementation)")
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    if missing:
        raise ImportError(f"libcapdec.so does not export {missing}")
    p, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
    lib.capdec_last_error.restype = C.c_char_p
    lib.capdec_version.restype = C.c_int
    lib.capdec_launch_count.restype = C.c_int64
    lib.capdec_create.argtypes = [C.POINTER(Config), C.POINTER(p)]
    lib.capdec_destroy.argtypes = [p]
    lib.capdec_destroy.restype = None
    lib.capdec_set_weight.argtypes = [p, C.c_char_p, p, C.POINTER(i64), i32, p]
    lib.capdec_finalize.argtypes = [p, p]
    lib.capdec_workspace_bytes.argtypes = [p, i32, i32, i32, i32]
    lib.capdec_workspace_bytes.restype = sz
    lib.capdec_decode_beam.argtypes = [p, p, p, p, i32, i32, i32, i32, f32, p, p, p, p, p, p, p, sz, p]
    lib.capdec_decode_greedy.argtypes = [p, p, p, p, i32, i32, i32, i32, p, p, p, sz, p]
    lib.capdec_decode_sample.argtypes = [p, p, p, p, i32, i32, i32, i32, i32, p, p, p, p, sz, p]
    lib.capdec_forward_teacher.argtypes = [p, p, i32, i32, p, i32, C.POINTER(i32), p, p, p, sz, p]
    lib.capdec_attention_forward.argtypes = [p, p, p, p, p, p, i32, i32, i32, p, p, p, sz, p]
    lib.capdec_decode_beam_host.argtypes = [p, p, p, i32, i32, i32, i32, f32, i32, p, p, p]
    lib.capdec_forward_tokens.argtypes = [p, p, p, p, i32, i32, i32, p, i32, i32, p, p, p, i32, p, sz, p]
    lib.capdec_trim_at_eos.argtypes = [p, i64, i32, i32, i32, i32, i32, p, i64, p, p]
    lib.capdec_source_bytes.argtypes = [i32, i32, i32, i32, i32]
    lib.capdec_source_bytes.restype = sz
    lib.capdec_tiles_bytes.argtypes = [p, i32, i32]
    lib.capdec_tiles_bytes.restype = sz
    lib.capdec_ingest_features.argtypes = [p, p, i32, i32, i32, i32, p, sz, p]
    lib.capdec_decode_beam_tiles.argtypes = [p, p, p, p, i32, i32, i32, i32, f32, p, p, p, p, p, p, p, sz, p]
    lib.capdec_pack_p24_host.argtypes = [p, i64, i64, p]
    lib.capdec_decode_beam_host_ex.argtypes = [p, p, i32, i32, p, p, i32, i32, i32, i32, f32, i32, p, p, p]
    lib.capdec_stage_timing.argtypes = [p, i32]
    lib.capdec_stage_times.argtypes = [p, C.POINTER(f32), C.POINTER(i32)]
    lib.capdec_linear.argtypes = [i32, p, i64, p, i64, p, p, i64, i32, i32, i32, p]
    lib.capdec_lse_topk.argtypes = [p, i64, i32, i32, i32, p, p, p, p]
    lib.capdec_linear_topk_workspace.argtypes = [i32, i32, i32]
    lib.capdec_linear_topk_workspace.restype = sz
    lib.capdec_linear_topk.argtypes = [i32, p, i64, p, i64, p, i32, i32, i32, i32, p, p, p, p, sz, p]
    return lib


lib = _load()


def check(status: int) -> None:
    if status != 0:
        raise CapdecError(status, lib.capdec_last_error().decode("utf-8", "replace"))
