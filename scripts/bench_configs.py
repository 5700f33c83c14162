"""Throughput of the BASELINE parity configs (configs[0], [2], [3], [4]) on one GPU -- not bench.py's headline
(configs[1]); reported in DESIGN.md so every decoder family has a measured number.
  python scripts/bench_configs.py [c1 c3 c4 c5]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import capdec_b200 as cd
from capdec_b200 import engine as eng_mod
from tests.helpers import gpt2_decoder, legacy_weights, transformer_decoder

dev = torch.device("cuda:0")
torch.set_grad_enabled(False)


def rnd(shape, seed, relu=False):
    x = torch.randn(*shape, generator=torch.Generator(device=dev).manual_seed(seed), device=dev)
    return x.relu_() if relu else x


def timeit(fn, n=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = cd.launch_count()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (cd.launch_count() - l0) // n


def stage(m, fn):
    eng = m._engine(dev)
    eng.stage_timing(True)
    fn()
    st = eng.stage_times()
    eng.stage_timing(False)
    return {k: round(v[0], 2) for k, v in st.items()}


which = sys.argv[1:] or ["c1", "c3", "c4", "c5"]
out = {}
if "c1" in which:
    for prec in ("fp32", "bf16x3"):
        m, _ = legacy_weights(10000, 0); m.precision = prec; m = m.to(dev)
        enc = rnd((64, 196, 2048), 1, relu=True)
        ms, nl = timeit(lambda: m.beam_search(enc, beam_size=3, max_length=20))
        out[f"c1_legacy_beam3_64img_{prec}"] = {"ms": ms, "images_per_s": 64 / ms * 1e3, "launches": nl,
                                                 "stage_ms": stage(m, lambda: m.beam_search(enc, beam_size=3, max_length=20))}
if "c2src" in which:
    from tests.helpers import lstm_decoder
    for kind, heads in (("soft", 8), ("aoa", 8)):
        m, _ = lstm_decoder(kind, H=512, layers=1, heads=heads, V=10000); m.precision = "bf16x3"; m = m.to(dev)
        ef = {"features": rnd((4096, 196, 512), 3), "pooled_features": rnd((4096, 512), 4)}
        ms, nl = timeit(lambda: m.generate(ef, 20, num_beams=5))
        out[f"c2src_lstm_{kind}_beam5_4096img_bf16x3"] = {"ms": ms, "images_per_s": 4096 / ms * 1e3, "launches": nl,
                                                          "stage_ms": stage(m, lambda: m.generate(ef, 20, num_beams=5))}
if "c3" in which:
    m, _ = transformer_decoder(H=768, layers=6, heads=8, V=10000, max_length=50); m.precision = "bf16x3"; m = m.to(dev)
    ef = {"features": rnd((2048, 196, 768), 5)}
    ms, nl = timeit(lambda: m.generate(ef, 20, num_beams=3))
    out["c3_transformer_beam3_2048img_bf16x3"] = {"ms": ms, "images_per_s": 2048 / ms * 1e3, "launches": nl,
                                                  "stage_ms": stage(m, lambda: m.generate(ef, 20, num_beams=3))}
if "c4" in which or "c4only" in which:
    for prec in (("bf16",) if "c4only" in which else ("bf16", "bf16x3")):
        m, _ = gpt2_decoder(H=768, layers=12, heads=12, V=50257, max_length=64); m.precision = prec; m = m.to(dev)
        ef = {"pooled_features": rnd((1024, 768), 6)}
        ms, nl = timeit(lambda: m.generate(ef, 20, num_beams=5))
        out[f"c4_gpt2_124m_beam5_1024img_{prec}"] = {"ms": ms, "images_per_s": 1024 / ms * 1e3, "launches": nl,
                                                      "stage_ms": stage(m, lambda: m.generate(ef, 20, num_beams=5))}
if "c5" in which:
    m, _ = gpt2_decoder(H=768, layers=12, heads=12, V=50257, max_length=64); m.precision = "bf16"; m = m.to(dev)
    ef = {"pooled_features": rnd((512, 768), 7)}
    u = torch.rand(512 * 6, 19, device=dev)
    fn = lambda: m.generate(ef, 20, do_sample=True, num_samples=5, with_greedy=True, uniforms=u)
    ms, nl = timeit(fn)
    out["c5_scst_rollout_512img_bf16"] = {"ms": ms, "images_per_s": 512 / ms * 1e3, "launches": nl, "stage_ms": stage(m, fn)}
for k, v in out.items():
    print(k, json.dumps(v))
