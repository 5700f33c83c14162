"""Host-side session over a libcapdec handle.

PyTorch is plumbing here: it owns device memory (features, outputs, workspace) and the CUDA
stream; every computation happens in libcapdec's kernels behind the C ABI (include/capdec.h).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

from . import _capi
from ._capi import Config, check, lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()



_TORCH_DTYPE = {torch.float32: "f32", torch.bfloat16: "bf16", torch.float16: "f16"}


class Tiles:
    """A tile set written by capdec_ingest_features: the region features in the layout the decode kernels stream
    (include/capdec.h, "encoder -> decoder feature hand-off").  Owns its device buffer."""

    def __init__(self, buf: torch.Tensor, num_images: int, num_regions: int):
        self.buf, self.num_images, self.num_regions = buf, num_images, num_regions

    @property
    def device(self):
        return self.buf.device


class EngineOwner:
    """Mixin for the drop-in nn.Modules: the lazily built Engine (a ctypes handle) is per-process device state, not part
    of the module.  copy.deepcopy / pickle / torch.save(module) drop it (it is rebuilt on the next call), so EMA copies
    and snapshots work and two modules never share -- and double-free -- one handle."""

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_eng", None)
        state.pop("_eng_sig", None)
        return state


class Engine:
    """One capdec_handle bound to a module's parameters on one CUDA device."""

    def __init__(self, cfg: Config, state_dict: Dict[str, torch.Tensor], device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("capdec runs on CUDA devices only (no CPU fallback); move the module to cuda")
        self.device = device
        self.cfg = cfg
        self._h = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.capdec_create(C.byref(cfg), C.byref(self._h)))
            s = _stream(device)
            for name, t in state_dict.items():
                t = _f32(t, device)
                shape = (C.c_int64 * t.dim())(*t.shape)
                check(lib.capdec_set_weight(self._h, name.encode(), _ptr(t), shape, t.dim(), s))
            check(lib.capdec_finalize(self._h, s))
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and lib is not None:      # `lib` is already None during interpreter shutdown
            lib.capdec_destroy(h)
            self._h = None

    # -- workspace ------------------------------------------------------------------------------
    def _workspace(self, B: int, L: int, k: int, T: int) -> torch.Tensor:
        need = int(lib.capdec_workspace_bytes(self._h, B, L, k, T))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        if os.environ.get("CAPDEC_POISON_WORKSPACE"):     # tests: every call starts from NaN-filled scratch
            self._ws.fill_(0xFF)
        return self._ws

    def _mask(self, key_padding_mask, B, L):
        if key_padding_mask is None:
            return None
        m = key_padding_mask.to(device=self.device)
        if m.shape != (B, L):
            raise ValueError(f"key_padding_mask must be [B,L]=({B},{L}), got {tuple(m.shape)}")
        return m.to(torch.uint8).contiguous()

    # -- decode ---------------------------------------------------------------------------------
    def decode_beam(self, features, pooled, key_padding_mask, num_beams, max_length, length_penalty=1.0,
                    trace=False):
        tiles = features if isinstance(features, Tiles) else None
        if tiles is not None:
            B, L = tiles.num_images, tiles.num_regions
        else:
            feats = _f32(features, self.device)
            B, L = feats.shape[0], feats.shape[1]
        pooled = None if pooled is None else _f32(pooled, self.device)
        mask = self._mask(key_padding_mask, B, L)
        dev = self.device
        tok = torch.empty(B, max_length, dtype=torch.int32, device=dev)
        length = torch.empty(B, dtype=torch.int32, device=dev)
        score = torch.empty(B, dtype=torch.float32, device=dev)
        steps, k2 = max_length - 1, 2 * num_beams
        dlp = torch.empty(steps, B, k2, dtype=torch.float32, device=dev) if trace else None
        dtok = torch.empty(steps, B, k2, dtype=torch.int32, device=dev) if trace else None
        dbeam = torch.empty(steps, B, k2, dtype=torch.int32, device=dev) if trace else None
        with torch.cuda.device(dev):
            ws = self._workspace(B, L, num_beams, max_length)
            fn, src = (lib.capdec_decode_beam_tiles, tiles.buf) if tiles is not None else (lib.capdec_decode_beam, feats)
            check(fn(self._h, _ptr(src), _ptr(pooled), _ptr(mask), B, L, num_beams, max_length,
                     float(length_penalty), _ptr(tok), _ptr(length), _ptr(score), _ptr(dlp),
                     _ptr(dtok), _ptr(dbeam), _ptr(ws), ws.numel(), _stream(dev)))
        out = {"tokens": tok, "lengths": length, "scores": score}
        if trace:
            out.update(top_logprob=dlp, top_token=dtok, top_beam=dbeam)
        return out

    def decode_greedy(self, features, pooled, key_padding_mask, max_length, start_token_id=1, want_alpha=True):
        feats = _f32(features, self.device)
        B, L = feats.shape[0], feats.shape[1]
        pooled = None if pooled is None else _f32(pooled, self.device)
        mask = self._mask(key_padding_mask, B, L)
        dev = self.device
        tok = torch.empty(B, max_length, dtype=torch.int32, device=dev)
        alpha = torch.empty(B, max_length, L, dtype=torch.float32, device=dev) if want_alpha else None
        with torch.cuda.device(dev):
            ws = self._workspace(B, L, 1, max_length)
            check(lib.capdec_decode_greedy(self._h, _ptr(feats), _ptr(pooled), _ptr(mask), B, L, max_length,
                                           int(start_token_id), _ptr(tok), _ptr(alpha), _ptr(ws), ws.numel(),
                                           _stream(dev)))
        return tok, alpha

    def decode_sample(self, features, pooled, key_padding_mask, num_samples, with_greedy, max_length, uniforms):
        feats = _f32(features, self.device)
        B, L = feats.shape[0], feats.shape[1]
        pooled = None if pooled is None else _f32(pooled, self.device)
        mask = self._mask(key_padding_mask, B, L)
        dev = self.device
        k = num_samples + (1 if with_greedy else 0)
        R = B * k
        u = _f32(uniforms, dev)
        if u.shape != (R, max_length - 1):
            raise ValueError(f"uniforms must be [{R},{max_length - 1}], got {tuple(u.shape)}")
        tok = torch.empty(R, max_length, dtype=torch.int32, device=dev)
        lp = torch.empty(R, max_length - 1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            ws = self._workspace(B, L, k, max_length)
            check(lib.capdec_decode_sample(self._h, _ptr(feats), _ptr(pooled), _ptr(mask), B, L, num_samples,
                                           1 if with_greedy else 0, max_length, _ptr(u), _ptr(tok), _ptr(lp),
                                           _ptr(ws), ws.numel(), _stream(dev)))
        return tok, lp

    def forward_teacher(self, features, captions, dec_len):
        feats = _f32(features, self.device)
        B, L = feats.shape[0], feats.shape[1]
        dev = self.device
        T = max(dec_len)
        V = self.cfg.vocab_size
        caps = captions.to(device=dev, dtype=torch.int32).contiguous()
        preds = torch.zeros(B, T, V, dtype=torch.float32, device=dev)
        alphas = torch.zeros(B, T, L, dtype=torch.float32, device=dev)
        dl = (C.c_int32 * B)(*[int(x) for x in dec_len])
        with torch.cuda.device(dev):
            ws = self._workspace(B, L, 1, T + 1)
            check(lib.capdec_forward_teacher(self._h, _ptr(feats), B, L, _ptr(caps), caps.shape[1], dl, _ptr(preds),
                                             _ptr(alphas), _ptr(ws), ws.numel(), _stream(dev)))
        return preds, alphas

    def attention_forward(self, query, features, key_padding_mask, memory_state, cell_state, rows_per_image=1):
        feats = _f32(features, self.device)
        B, L = feats.shape[0], feats.shape[1]
        q = _f32(query, self.device)
        R, H = q.shape
        if R != B * rows_per_image:
            raise ValueError(f"query rows {R} != images {B} * rows_per_image {rows_per_image}")
        mask = self._mask(key_padding_mask, B, L)
        mem = None if memory_state is None else _f32(memory_state, self.device)
        cell = None if cell_state is None else _f32(cell_state, self.device)
        ctx = torch.empty(R, H, dtype=torch.float32, device=self.device)
        w = torch.empty(R, L, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ws = self._workspace(B, L, rows_per_image, 2)
            check(lib.capdec_attention_forward(self._h, _ptr(q), _ptr(feats), _ptr(mask), _ptr(mem), _ptr(cell), B, L,
                                               rows_per_image, _ptr(ctx), _ptr(w), _ptr(ws), ws.numel(),
                                               _stream(self.device)))
        return ctx, w

    def decode_beam_host(self, features_host, pooled_host, num_beams, max_length, length_penalty=1.0,
                         chunk_images=0, out=None, layout="bld", dtype=None, num_regions=None, mask_host=None):
        """End-to-end path: host (pinned) buffers in, host buffers out, copies inside the call.  `features_host` is
        fp32 / bf16 / fp16 in `layout` ("bld" [B,L,D], "nchw" [B,D,h,w], "cls_bld" [B,1+L,D]) or, with dtype="p24", the
        uint8 block written by pack_p24_host."""
        assert features_host.device.type == "cpu" and features_host.is_contiguous()
        if dtype is None:
            dtype = _TORCH_DTYPE[features_host.dtype]
        B = features_host.shape[0]
        if num_regions is None:
            if layout in ("bdl", "nchw"):
                num_regions = features_host[0, 0].numel()
            elif layout == "cls_bld":
                num_regions = features_host.shape[1] - 1
            else:
                num_regions = features_host.shape[1]
        L = int(num_regions)
        need = int(lib.capdec_source_bytes(_capi.LAYOUT[layout], _capi.DTYPE[dtype], B, L, self.cfg.feature_dim))
        if features_host.numel() * features_host.element_size() != need:
            raise ValueError(f"features_host holds {features_host.numel() * features_host.element_size()} bytes, "
                             f"{layout}/{dtype} for {B} images x {L} regions x {self.cfg.feature_dim} needs {need}")
        if mask_host is not None:
            mask_host = mask_host.to(torch.uint8).contiguous()
            if tuple(mask_host.shape) != (B, L):
                raise ValueError(f"mask_host must be [B,L]=({B},{L})")
        if out is None:
            out = {"tokens": torch.empty(B, max_length, dtype=torch.int32).pin_memory(),
                   "lengths": torch.empty(B, dtype=torch.int32).pin_memory(),
                   "scores": torch.empty(B, dtype=torch.float32).pin_memory()}
        with torch.cuda.device(self.device):
            check(lib.capdec_decode_beam_host_ex(self._h, _ptr(features_host), _capi.LAYOUT[layout], _capi.DTYPE[dtype],
                                                 _ptr(pooled_host), _ptr(mask_host), B, L, num_beams, max_length,
                                                 float(length_penalty), int(chunk_images), _ptr(out["tokens"]),
                                                 _ptr(out["lengths"]), _ptr(out["scores"])))
        return out

    # -- encoder -> decoder hand-off ---------------------------------------------------------------
    def ingest_features(self, src: torch.Tensor, layout: str = "bld", dtype: Optional[str] = None,
                        num_regions: Optional[int] = None) -> Tiles:
        """What an encoder emits -> the tile set the decode streams, in one pass (capdec_ingest_features).
        src: device tensor, fp32 / bf16 / fp16, in `layout`: "bld" [B,L,D]; "nchw" [B,D,h,w] (resnet trunk output before
        models/encoder.py:15's permute); "cls_bld" [B,1+L,D] (ViT / CLIP last_hidden_state, CLS dropped as in
        src/models/encoders.py:122,213).  dtype="p24" takes the uint8 block of pack_p24_host (needs num_regions)."""
        src = src.detach().to(self.device).contiguous()
        if dtype is None:
            dtype = _TORCH_DTYPE[src.dtype]
        B = src.shape[0]
        if num_regions is None:
            num_regions = src[0, 0].numel() if layout in ("bdl", "nchw") else src.shape[1] - (1 if layout == "cls_bld" else 0)
        L = int(num_regions)
        need = int(lib.capdec_source_bytes(_capi.LAYOUT[layout], _capi.DTYPE[dtype], B, L, self.cfg.feature_dim))
        if src.numel() * src.element_size() != need:
            raise ValueError(f"source holds {src.numel() * src.element_size()} bytes, {layout}/{dtype} for {B} x {L} x "
                             f"{self.cfg.feature_dim} needs {need}")
        nbytes = int(lib.capdec_tiles_bytes(self._h, B, L))
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.capdec_ingest_features(self._h, _ptr(src), _capi.LAYOUT[layout], _capi.DTYPE[dtype], B, L, _ptr(buf),
                                             buf.numel(), _stream(self.device)))
        return Tiles(buf, B, L)

    # -- teacher-forced pass over given tokens (SCST boundary) -------------------------------------
    def forward_tokens(self, features, pooled, key_padding_mask, tokens, rows_per_image=1, want_logits=True,
                       want_logprob=False, want_alpha=False, mask_pad_keys=True):
        """-> (logits [R,t,V] | None, logprob [R,t-1] | None, alpha [R,t,L] | None) for tokens [R,t]."""
        dev = self.device
        if features is None:        # GPT-2 reads pooled_features only
            B, L, feats = pooled.shape[0], 1, None
        else:
            feats = _f32(features, dev)
            B, L = feats.shape[0], feats.shape[1]
        pooled = None if pooled is None else _f32(pooled, dev)
        mask = self._mask(key_padding_mask, B, L)
        tok = tokens.to(device=dev, dtype=torch.int32).contiguous()
        R, t = tok.shape
        if R != B * rows_per_image:
            raise ValueError(f"tokens rows {R} != images {B} * rows_per_image {rows_per_image}")
        V = self.cfg.vocab_size
        logits = torch.empty(R, t, V, dtype=torch.float32, device=dev) if want_logits else None
        lp = torch.empty(R, max(t - 1, 0), dtype=torch.float32, device=dev) if want_logprob else None
        alpha = torch.empty(R, t, L, dtype=torch.float32, device=dev) if want_alpha else None
        dummy = torch.empty(B, 1, 4, device=dev) if feats is None else feats
        with torch.cuda.device(dev):
            ws = self._workspace(B, L, rows_per_image, t + 1)
            check(lib.capdec_forward_tokens(self._h, _ptr(dummy), _ptr(pooled), _ptr(mask), B, L, rows_per_image, _ptr(tok),
                                            tok.stride(0), t, _ptr(logits), _ptr(lp), _ptr(alpha), 1 if mask_pad_keys else 0,
                                            _ptr(ws), ws.numel(), _stream(dev)))
        return logits, lp, alpha


STAGES = ("prologue", "small_gemm", "attention", "gate_gemm", "vocab_gemm", "select", "beam", "gather")


def _stage_timing(self, enable: bool) -> None:
    check(lib.capdec_stage_timing(self._h, 1 if enable else 0))


def _stage_times(self):
    """-> {stage: (milliseconds, launches)} since the last call (waits for the recorded events)."""
    ms = (C.c_float * 8)()
    n = (C.c_int32 * 8)()
    check(lib.capdec_stage_times(self._h, ms, n))
    return {name: (float(ms[i]), int(n[i])) for i, name in enumerate(STAGES)}


Engine.stage_timing = _stage_timing
Engine.stage_times = _stage_times


def trim_at_eos_device(tokens: torch.Tensor, eos_token_id: int, pad_token_id: int = 0, keep_eos: bool = True):
    """capdec_trim_at_eos on a CUDA int32 [B,T] block -> (trimmed int32 [B,T], lengths int32 [B]); one kernel, no sync."""
    assert tokens.is_cuda and tokens.dtype == torch.int32 and tokens.dim() == 2
    tokens = tokens.contiguous()
    B, T = tokens.shape
    out = torch.empty_like(tokens)
    lengths = torch.empty(B, dtype=torch.int32, device=tokens.device)
    with torch.cuda.device(tokens.device):
        check(lib.capdec_trim_at_eos(_ptr(tokens), tokens.stride(0), B, T, int(eos_token_id), int(pad_token_id),
                                     1 if keep_eos else 0, _ptr(out), out.stride(0), _ptr(lengths), _stream(tokens.device)))
    return out, lengths


def pack_p24_host(features: torch.Tensor) -> torch.Tensor:
    """fp32 host features [B, ...] -> the CAPDEC_DT_P24 source block (uint8 [B, 3 * elems]): per image a uint16 plane
    (top 16 bits) followed by a uint8 plane (round(low16 / 257)).  3 bytes per element, 16 significant bits."""
    f = features.detach().to("cpu", torch.float32).contiguous()
    B = f.shape[0]
    elems = f[0].numel() if B else 0
    out = torch.empty(B, 3 * elems, dtype=torch.uint8)
    check(lib.capdec_pack_p24_host(_ptr(f), B, elems, _ptr(out)))
    return out


def launch_count() -> int:
    return int(lib.capdec_launch_count())


def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], precision: str = "fp32") -> torch.Tensor:
    """Stage-level entry (unit tests): nn.Linear through libcapdec's GEMM."""
    M, K = a.shape
    N = w.shape[0]
    c = torch.empty(M, N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(lib.capdec_linear(_capi.PREC[precision], _ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(c),
                                c.stride(0), M, N, K, _stream(a.device)))
    return c


def lse_topk(logits: torch.Tensor, topk: int):
    R, V = logits.shape
    lp = torch.empty(R, topk, dtype=torch.float32, device=logits.device)
    idx = torch.empty(R, topk, dtype=torch.int32, device=logits.device)
    lse = torch.empty(R, dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        check(lib.capdec_lse_topk(_ptr(logits), logits.stride(0), R, V, topk, _ptr(lp), _ptr(idx), _ptr(lse),
                                  _stream(logits.device)))
    return lp, idx, lse


def linear_topk(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], topk: int, precision: str = "tf32x3"):
    """Stage-level entry (unit tests): vocabulary projection fused with log-softmax + top-k (no logits in HBM)."""
    M, K = a.shape
    N = w.shape[0]
    dev = a.device
    lp = torch.empty(M, topk, dtype=torch.float32, device=dev)
    idx = torch.empty(M, topk, dtype=torch.int32, device=dev)
    lse = torch.empty(M, dtype=torch.float32, device=dev)
    nbytes = int(lib.capdec_linear_topk_workspace(M, N, topk))
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev)
    if os.environ.get("CAPDEC_POISON_WORKSPACE"):
        ws.fill_(0xFF)
    with torch.cuda.device(dev):
        check(lib.capdec_linear_topk(_capi.PREC[precision], _ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(bias), M, N, K,
                                     topk, _ptr(lp), _ptr(idx), _ptr(lse), _ptr(ws), ws.numel(), _stream(dev)))
    return lp, idx, lse
