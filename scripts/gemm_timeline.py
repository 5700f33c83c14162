"""Phase timeline of one tcgen05 GEMM launch (debug build only: CAPDEC_NVCC_FLAGS=-DCAPDEC_TIMELINE bash csrc/build.sh).
Every CTA stamps %globaltimer / clock64 at: entry (0), set-up done (1), pdl wait done (2), first TMA issued (3), last TMA
issued (4), first operands landed (5), MMAs of work item i issued (8+i), epilogue of item i started / finished (16+2i /
17+2i), epilogue warps done (62), exit (63).  Prints medians over the CTAs in microseconds relative to the launch's
first entry stamp.
  python scripts/gemm_timeline.py [precision]            stand-alone GPT-2 / legacy shapes through capdec_linear
  python scripts/gemm_timeline.py decode [precision]     in situ: the LAST stamped GEMM launch of a GPT-2 124M beam-5 decode
  python scripts/gemm_timeline.py legacy [images]        in situ: ... of the headline legacy decode (bf16x3)
  python scripts/gemm_timeline.py c3                     in situ: ... of the configs[2] transformer decode (bf16x3)
For the in-situ modes build with -DCAPDEC_TL_EPI=<epilogue id> (0 store, 2 LSTM, 6 GELU-tanh, 7 fused top-k) so that only
launches with that epilogue stamp; scripts/build_variant.sh builds such variants next to the product library."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import capdec_b200 as cd

lib = cd._capi.lib
assert hasattr(lib, "capdec_debug_timeline"), "build with CAPDEC_NVCC_FLAGS=-DCAPDEC_TIMELINE"
lib.capdec_debug_timeline.argtypes = [C.c_void_p, C.c_int]
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
SHAPES = [("c_attn", 5120, 2304, 768), ("attn.c_proj", 5120, 768, 768), ("mlp.c_fc", 5120, 3072, 768),
          ("mlp.c_proj", 5120, 768, 3072), ("c3.in_proj", 6144, 2304, 768), ("rows2560", 2560, 2048, 2560)]


def fetch():
    buf = np.zeros((2, 320, 64), dtype=np.uint64)
    assert lib.capdec_debug_timeline(buf.ctypes.data_as(C.c_void_p), 0) == 0
    return buf


def report(g, c, title):
    live = g[:, 0] > 0
    g, c = g[live].astype(np.int64), c[live].astype(np.int64)
    t0 = g[:, 0].min()
    print(f"== {title}: {live.sum()} CTAs; first entry -> last exit {(g[:, 63].max() - t0) / 1e3:.2f} us")

    def med(i):
        x = g[:, i]
        x = x[x > 0]
        return (np.median(x) - t0) / 1e3 if len(x) else float("nan")

    def mx(i):
        x = g[:, i]
        x = x[x > 0]
        return (x.max() - t0) / 1e3 if len(x) else float("nan")

    def cyc(i, j):
        ok = (c[:, i] > 0) & (c[:, j] > 0)
        return float(np.median(c[ok, i] - c[ok, j])) if ok.any() else float("nan")
    print(f"   entry med {med(0):.2f} max {mx(0):.2f} | setup done {med(1):.2f} | pdl wait done {med(2):.2f} | first TMA issued {med(3):.2f} "
          f"| first operands landed {med(5):.2f} | last TMA issued {med(4):.2f}")
    items = [i for i in range(8) if (g[:, 8 + i] > 0).any()]
    print("   MMAs of item i issued:   " + "  ".join(f"{i}:{med(8 + i):.2f}" for i in items))
    items = [i for i in range(16) if (g[:, 16 + 2 * i] > 0).any()]
    print("   epilogue start / end:    " + "  ".join(f"{i}:{med(16 + 2 * i):.2f}/{med(17 + 2 * i):.2f}" for i in items))
    print(f"   cycles: setup {cyc(1, 0):.0f}, pdl wait {cyc(2, 1):.0f}, first operands after wait {cyc(5, 2):.0f}, "
          f"epilogue of item 0 {cyc(17, 16):.0f}, last epilogue end -> exit {cyc(63, 62):.0f}")
    print(f"   epilogue warps done med {med(62):.2f} max {mx(62):.2f} | exit med {med(63):.2f} max {mx(63):.2f}")


if prec == "legacy":
    # in-situ probe on the headline config (legacy decoder, beam 5, 4096 images, bf16x3): -DCAPDEC_TL_EPI=7 stamps the
    # vocabulary GEMM (fused top-k), =2 the gate GEMM (LSTM epilogue)
    from tests.helpers import legacy_weights
    torch.set_grad_enabled(False)
    m, _ = legacy_weights(10000, 0); m.precision = "bf16x3"; m = m.to(dev)
    enc = torch.randn(int(sys.argv[2]) if len(sys.argv) > 2 else 4096, 196, 2048, device=dev).relu_()
    for _ in range(2):
        m.beam_search(enc, beam_size=5, max_length=20)
    torch.cuda.synchronize()
    lib.capdec_debug_timeline(None, 1)
    m.beam_search(enc, beam_size=5, max_length=20)
    torch.cuda.synchronize()
    g, c = fetch()
    report(g, c, f"last stamped GEMM launch of the legacy beam-5 decode, {enc.shape[0]} images, bf16x3")
    sys.exit(0)

if prec == "c3":
    # in-situ probe on configs[2] (transformer decoder, 2048 images, beam 3, bf16x3): -DCAPDEC_TL_EPI=5 stamps linear1 + GELU,
    # =0 the plain projections (the last one of a decode is the last layer's linear2)
    from tests.helpers import transformer_decoder
    torch.set_grad_enabled(False)
    m, _ = transformer_decoder(H=768, layers=6, heads=8, V=10000, max_length=50); m.precision = "bf16x3"; m = m.to(dev)
    ef = {"features": torch.randn(2048, 196, 768, device=dev)}
    for _ in range(2):
        m.generate(ef, 20, num_beams=3)
    torch.cuda.synchronize()
    lib.capdec_debug_timeline(None, 1)
    m.generate(ef, 20, num_beams=3)
    torch.cuda.synchronize()
    g, c = fetch()
    report(g, c, "last stamped GEMM launch of the configs[2] transformer decode, 2048 images, beam 3, bf16x3")
    sys.exit(0)

if prec == "decode":
    # in-situ probe (build with -DCAPDEC_TL_EPI=<epilogue id>): the LAST launch with that epilogue inside a GPT-2 124M decode
    from tests.helpers import gpt2_decoder
    torch.set_grad_enabled(False)
    mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    m, _ = gpt2_decoder(H=768, layers=12, heads=12, V=50257, max_length=64); m.precision = mode; m = m.to(dev)
    ef = {"pooled_features": torch.randn(1024, 768, device=dev)}
    for _ in range(2):
        m.generate(ef, 20, num_beams=5)
    torch.cuda.synchronize()
    lib.capdec_debug_timeline(None, 1)
    m.generate(ef, 20, num_beams=5)
    torch.cuda.synchronize()
    g, c = fetch()
    report(g, c, f"last stamped GEMM launch of a GPT-2 124M beam-5 decode, 1024 images, {mode}")
    sys.exit(0)

for name, M, N, K in SHAPES:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.02
    b = torch.randn(N, device=dev)
    for _ in range(3):
        cd.engine.linear(a, w, b, prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        cd.engine.linear(a, w, b, prec)
    e1.record()
    torch.cuda.synchronize()
    lib.capdec_debug_timeline(None, 1)
    cd.engine.linear(a, w, b, prec)
    torch.cuda.synchronize()
    g, c = fetch()
    live = g[:, 0] > 0
    g, c = g[live].astype(np.int64), c[live].astype(np.int64)
    t0 = g[:, 0].min()
    print(f"== {name} M={M} N={N} K={K} {prec}: {live.sum()} CTAs, 10 calls (split kernels + GEMM) {e0.elapsed_time(e1) / 10 * 1e3:.1f} us each; "
          f"GEMM first entry -> last exit {(g[:, 63].max() - t0) / 1e3:.2f} us")

    def med(i, rows=None):
        x = g[:, i] if rows is None else g[rows, i]
        x = x[x > 0]
        return (np.median(x) - t0) / 1e3 if len(x) else float("nan")

    def mx(i):
        x = g[:, i]
        x = x[x > 0]
        return (x.max() - t0) / 1e3 if len(x) else float("nan")

    print(f"   entry med {med(0):.2f} max {mx(0):.2f} | setup done {med(1):.2f} | pdl wait done {med(2):.2f} | first TMA issued {med(3):.2f} "
          f"| first operands landed {med(5):.2f} | last TMA issued {med(4):.2f}")
    items = [i for i in range(8) if (g[:, 8 + i] > 0).any()]
    print("   MMAs of item i issued:   " + "  ".join(f"{i}:{med(8 + i):.2f}" for i in items))
    items = [i for i in range(16) if (g[:, 16 + 2 * i] > 0).any()]
    print("   epilogue start / end:    " + "  ".join(f"{i}:{med(16 + 2 * i):.2f}/{med(17 + 2 * i):.2f}" for i in items))
    # clock-cycle view of the same phases inside a CTA (medians of per-CTA differences)
    def cyc(i, j):
        ok = (c[:, i] > 0) & (c[:, j] > 0)
        return float(np.median(c[ok, i] - c[ok, j])) if ok.any() else float("nan")
    print(f"   cycles: setup {cyc(1, 0):.0f}, pdl wait {cyc(2, 1):.0f}, first operands after wait {cyc(5, 2):.0f}, "
          f"epilogue of item 0 {cyc(17, 16):.0f}, last epilogue end -> exit {cyc(63, 62):.0f}")
    print(f"   epilogue warps done med {med(62):.2f} max {mx(62):.2f} | exit med {med(63):.2f} max {mx(63):.2f}")
