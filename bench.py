#!/usr/bin/env python
"""bench.py -- captioned images/s of the batched beam-search decode hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): legacy ResNet-101 + LSTM + soft attention (models/decoder.py),
beam 5, max_len 20, vocab 10k, 4096 images of synthetic 14x14x2048 features IN TOTAL, sharded over the N GPUs
(strong scaling, SURVEY.md section 8(e): 4096 -> 512 per GPU at N=8; each rank decodes its own contiguous shard, NCCL
all-gathers the finished captions + lengths + scores; no data-path collective).  A "step" is one full decode of the
batch (prologue + 19 beam steps) through the public drop-in API, `capdec_b200.Decoder.beam_search`.

  value         images/s, device-resident features, CUDA events on the launching stream, max over ranks
  e2e           the same through the host-buffer entry point capdec_decode_beam_host (pinned host fp32 features copied
                host->device inside the timed region, captions copied back); e2e_formats: bf16 / p24 host features
  roofline      attention kernel: algorithmic bytes / live CUDA-event time of that kernel vs measured HBM peak
  cpu_baseline  the oracle port of the reference (torch fp32, all host threads) on a bounded sample
  parity_in_run the GPU captions of that same sample against the oracle's (identical-beam fraction, max |d score|)
  fp32_class    the tf32x3 mode (fp32-class operands: 3-term TF32 split) over the same --steps, with its own roofline
  configs       BASELINE configs[2..4] (transformer / GPT-2 beam / SCST rollout) at their full sizes, one GPU
  weak_scaling  (N > 1) 4096 images PER GPU, the round-1 figure

  --impl reference   times the reference's own CPU implementation of the path.  The reference is pure
                     Python and /root/reference does not travel to the GPU box, so this is the oracle port
                     (oracle/legacy.py + oracle/beam.py, pinned bit-exact to the reference modules).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

VOCAB, BEAM, MAXLEN, L, D, A = 10000, 5, 20, 196, 2048, 512
STEPS_PER_DECODE = MAXLEN - 1
# per image-step the attention kernel must read att1 [L,A] + feats [L,D] once (fp32)
ATTN_BYTES_PER_IMAGE_STEP = L * (A + D) * 4


def env_int(name, default):
    return int(os.environ.get(name, default))


class ClockSampler:
    """SM clock + throttle reasons of this rank's GPU, sampled in-process through NVML every 20 ms while the timed
    region runs (nvidia-smi takes longer to start on an 8-GPU box than a short timed region lasts; it is only the
    fallback when pynvml is missing)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nv, self.h, self.stop = index, [], None, None, None, threading.Event()
        self.sm, self.mx, self.reasons, self.power, self.power_limit = [], None, set(), [], None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            try:
                self.power_limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0
            except Exception:
                self.power_limit = None
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap),
                 ("hw_power_brake_slowdown", nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown))
        while True:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names:
                    if mask & bit:
                        self.reasons.add(n)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            if self.stop.wait(0.01):
                break

    def __enter__(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.nv is not None:
            self.stop.set()
            self.t.join(timeout=2)
        elif self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        if self.nv is not None:
            sm = sorted(self.sm)
            pw = sorted(self.power)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml", "power_w": pw[len(pw) // 2] if pw else None,
                    "power_limit_w": self.power_limit}
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """sustained cuBLAS bf16 TFLOP/s (the GEMMs are timed inside a long step, so the sustained figure applies)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), "measured sustained bf16 (MEASURED_PEAKS.json)"
    return 1400.0, "fallback (B200_PROFILING.md)"


WORKLOAD = "configs[1]: ResNet-101 features + LSTM + soft attention, beam=5, max_len=20, vocab 10k, 4096 images sharded over the GPUs"
TOTAL_IMAGES = 4096
FEAT_BLOCK = 256     # images per seeded block: any rank can generate exactly its shard of the global batch


def global_features(lo, hi, pin=False):
    """images [lo, hi) of the global synthetic batch: relu(randn(14,14,2048)) per image, seeded per block
    of 256 images so shards are reproducible independently of the world size."""
    n = hi - lo
    out = torch.empty(n, 14, 14, D, dtype=torch.float32)
    if pin:
        out = out.pin_memory()
    b0 = lo // FEAT_BLOCK
    for blk in range(b0, (hi + FEAT_BLOCK - 1) // FEAT_BLOCK):
        g = torch.Generator().manual_seed(1234 + blk)
        x = torch.relu(torch.randn(FEAT_BLOCK, 14, 14, D, generator=g))
        s, e = max(lo, blk * FEAT_BLOCK), min(hi, (blk + 1) * FEAT_BLOCK)
        out[s - lo:e - lo] = x[s - blk * FEAT_BLOCK:e - blk * FEAT_BLOCK]
    return out


def cpu_decode(n_images, repeats=1):
    """The reference's CPU path (oracle port), fp32, every host thread, decode only, on the first n images of the
    global batch.  -> (images/s, seconds, oracle result).  Imports nothing of the product package."""
    from oracle import beam as obeam, legacy as olegacy
    from oracle.weights import legacy_state_dict
    torch.set_num_threads(os.cpu_count())
    sd = legacy_state_dict(VOCAB, 0)
    enc = global_features(0, n_images)
    best, res = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            res = obeam.beam_search(olegacy.LegacyStepper(sd, enc, BEAM), n_images, BEAM, MAXLEN)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_images / best, best, res


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n = env_int("CAPDEC_BENCH_CPU_IMAGES", 64)
    cpu_decode(min(n, 4))  # warm-up (thread pools, first-touch)
    times = []
    for _ in range(args.warmup):
        cpu_decode(n)
    for _ in range(args.steps):
        times.append(cpu_decode(n)[1])
    ms = 1000.0 * sum(times) / len(times)
    value = n / (ms / 1000.0)
    sample = f"{n} images x beam {BEAM} x {STEPS_PER_DECODE} steps per step (bounded sample of the {TOTAL_IMAGES}-image workload)"
    print(json.dumps({
        "impl": "reference", "metric": "captioned images/sec (beam=5, max_len=20)", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "total_images": TOTAL_IMAGES, "sample_images": n},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def time_device(fn, steps, barrier=None):
    """CUDA events on the current stream around `steps` calls of fn -> ms per call"""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    (barrier or torch.cuda.synchronize)()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    (barrier or torch.cuda.synchronize)()
    return e0.elapsed_time(e1) / steps


def attention_roofline(stage, images, tile_fmt):
    peak, peak_src = measured_peaks()
    att_ms, att_n = stage["attention"]
    per_launch_ms = att_ms / max(att_n, 1)
    attn_bytes = ATTN_BYTES_PER_IMAGE_STEP * {"bf16": 2, "p24": 3, "f32": 4}[tile_fmt] // 4
    achieved = attn_bytes * images / (per_launch_ms * 1e-3) / 1e9 if att_n else None
    traffic = None
    tf = os.path.join(ROOT, "profiles", "attention_traffic.json")
    if os.path.isfile(tf):
        tj = json.load(open(tf))
        if isinstance(tj.get(tile_fmt), dict) and images == TOTAL_IMAGES:      # the ncu capture is of the 4096-image launch
            traffic = tj[tile_fmt].get("dram_bytes_per_launch")
    return {"kernel": "additive_attention_stream_kernel<5,relu,2,%s>" % tile_fmt, "bound": "hbm", "achieved": achieved,
            "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
            "peak_source": peak_src, "tile_bytes_per_element": {"bf16": 2, "p24": 3, "f32": 4}[tile_fmt],
            "algorithmic_bytes_per_launch": attn_bytes * images, "avg_launch_ms": per_launch_ms, "launches": att_n}


def gemm_rooflines(stage, images, precision):
    """tensor-bound stages, reported two ways: per MMA FLOP actually issued (3 MMAs per product in the split modes; a TF32
    MMA counts double against the bf16 peak) and per ALGORITHMIC FLOP (2*M*N*K of the reference's fp32
    contraction, vocabulary not padded, no emulation terms)."""
    terms = {"fp32": 0, "tf32x3": 3, "bf16x3": 3, "bf16": 1, "tf32": 1}[precision]
    if not terms:
        return None
    tpeak, tsrc = measured_tensor_peak()
    scale = 2.0 if precision.startswith("tf32") else 1.0
    R, H4, Kg = images * BEAM, 4 * 512, 2048 + 512
    Nv = (VOCAB + 255) // 256 * 256 + A + D

    def rec(ms_n, flop_issued, flop_alg, shape, note=None):
        ms, n = ms_n
        if not n:
            return None
        t = ms / n * 1e-3
        issued, alg = flop_issued * terms * scale / t / 1e12, flop_alg / t / 1e12
        r = {"bound": "tensor", "achieved": issued, "peak": tpeak, "unit": "TFLOP/s (bf16-equivalent MMA work issued)",
             "frac": issued / tpeak, "achieved_algorithmic": alg, "frac_algorithmic": alg / tpeak, "shape": shape,
             "mma_terms": terms, "avg_launch_ms": ms / n}
        if note:
            r["note"] = note
        return r
    out = {"gate_gemm": rec(stage["gate_gemm"], 2.0 * R * H4 * Kg, 2.0 * R * H4 * (Kg + 512), [R, H4, Kg],
                            "algorithmic FLOPs include the embedding columns (K = 3072) that the per-token table removed from the GEMM"),
           "vocab_gemm": rec(stage["vocab_gemm"], 2.0 * R * Nv * 512, 2.0 * R * (VOCAB + A + D) * 512, [R, Nv, 512],
                             "vocabulary padded to 256 + the [dec_att|f_beta] tail; fused log-softmax/top-k epilogue"),
           "peak_source": tsrc}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="capdec", choices=["capdec", "reference"])
    ap.add_argument("--images", type=int, default=env_int("CAPDEC_BENCH_IMAGES", TOTAL_IMAGES), help="images in total (sharded over the GPUs)")
    ap.add_argument("--precision", default=os.environ.get("CAPDEC_BENCH_PRECISION", "bf16x3"))
    ap.add_argument("--chunk", type=int, default=env_int("CAPDEC_BENCH_CHUNK", 0),
                    help="e2e H2D pipeline chunk in images (0 = the library default, two images per SM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-modes", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # the JSON line must be the only thing on stdout, but NCCL prints its version banner there: route fd 1 to stderr for
    # the whole run and write the result line to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch.distributed as dist
    import capdec_b200 as cd
    from capdec_b200 import engine as eng_mod, sharding
    from tests.helpers import legacy_weights

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep this rank's threads (and, by first touch, its pinned feature buffer) on the CPUs NUMA-local to its GPU, so
        # the ranks' host->device streams do not all cross the same memory controller / socket link
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (os.cpu_count() + 63) // 64)
            cpus = [i * 64 + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=dev)

    N_total = args.images
    lo, hi = sharding.shard_range(N_total, rank, world)
    B = hi - lo
    model, _ = legacy_weights(VOCAB, 0)
    model.precision = args.precision
    model = model.to(dev)
    eng = model._engine(dev)

    # synthetic 14x14x2048 region features of this rank's shard, generated on the host then copied;
    # >= 0.8 GB per GPU >> the 126 MB L2, so no flush is needed between timed iterations
    feats_host = global_features(lo, hi, pin=not args.no_e2e)
    feats = feats_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(f=None):
        """the public API: Decoder.beam_search on the encoder's [B,14,14,2048] output, then the path's only exchange
        (all-gather of finished captions + lengths + scores, 88 B/image) through the product's sharding module"""
        out = model.beam_search(feats if f is None else f, beam_size=BEAM, max_length=MAXLEN, crop=False)
        return sharding.gather_captions(out, N_total if f is None else world * f.shape[0]) if world > 1 else out

    for _ in range(max(args.warmup, 3)):
        out = one_step()
    barrier()
    # host-side cost of enqueuing one decode (126 launches + tensor-map encodes) against its device time: when the
    # enqueue returns well before the GPU finishes, launch overhead is hidden and a CUDA graph has nothing to remove
    t0 = time.perf_counter()
    one_step()
    host_enqueue_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    # ---- timed region: device-resident features, CUDA events on the launching stream
    eng.stage_timing(True)
    launches0 = cd.launch_count()
    with ClockSampler(local_rank) as clocks:
        ms_total = time_device(one_step, args.steps, barrier) * args.steps
    launches = cd.launch_count() - launches0
    stage = eng.stage_times()
    eng.stage_timing(False)
    out = one_step()
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = N_total / (ms_step / 1000.0)

    # ---- N > 1: the gathered captions of the sharded run must equal a 1-GPU decode of the same images (rank 0 checks)
    sharded_ok = None
    if world > 1:
        if rank == 0:
            full = model.beam_search(global_features(0, N_total).to(dev), beam_size=BEAM, max_length=MAXLEN, crop=False)
            sharded_ok = all(bool(torch.equal(out[k], full[k])) for k in ("tokens", "lengths", "scores"))
            del full
            assert sharded_ok, "gathered captions of the sharded decode differ from the 1-GPU decode of the same images"
        barrier()

    # ---- N > 1: the weak-scaling figure of round 1 (4096 images PER GPU; the shard repeated to that size on the device)
    weak = None
    if world > 1 and not args.no_weak:
        reps = max(1, TOTAL_IMAGES // max(B, 1))
        big = feats.repeat(reps, 1, 1, 1)
        for _ in range(2):
            one_step(big)
        tw = torch.tensor([time_device(lambda: one_step(big), max(2, min(args.steps, 5)), barrier)], device=dev)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        weak = {"value": big.shape[0] * world / (float(tw.item()) / 1000.0), "unit": "images/s", "ms_per_step": float(tw.item()),
                "images_per_gpu": int(big.shape[0]), "scaling": "weak"}
        del big

    # ---- end to end: pinned host features -> C-ABI host entry point -> host captions
    def time_e2e(fh, **kw):
        host_out = {"tokens": torch.empty(B, MAXLEN, dtype=torch.int32).pin_memory(),
                    "lengths": torch.empty(B, dtype=torch.int32).pin_memory(),
                    "scores": torch.empty(B, dtype=torch.float32).pin_memory()}
        for _ in range(2):
            eng.decode_beam_host(fh, None, BEAM, MAXLEN, chunk_images=args.chunk, out=host_out, **kw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.decode_beam_host(fh, None, BEAM, MAXLEN, chunk_images=args.chunk, out=host_out, **kw)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ms = 1000.0 * float(dt.item()) / args.steps
        return ms, host_out

    e2e, e2e_formats = None, {}
    if not args.no_e2e:
        fh = feats_host.reshape(B, L, D)
        e2e_ms, host_out = time_e2e(fh)
        local_tok = out["tokens"][lo:hi] if world > 1 else out["tokens"]
        same = bool(torch.equal(host_out["tokens"], local_tok.cpu()))
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        chunk = args.chunk if args.chunk > 0 else 2 * sms
        e2e = {"value": N_total / (e2e_ms / 1000.0), "unit": "images/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(fh.numel() * 4) * world, "d2h_bytes_per_step": int(B * MAXLEN * 4 + B * 8) * world,
               "host_feature_format": "fp32 [B,196,2048] (the reference's decoder input)", "chunk_images": chunk,
               "matches_device_path": same}
        # the encoder hand-off formats: what an autocast encoder emits (bf16) and the 3-byte p24 block (16 significant
        # bits -- the precision the bf16x3 mode keeps of the features anyway); fewer bytes over PCIe, same decode
        fh16 = fh.bfloat16().pin_memory()
        ms16, ho16 = time_e2e(fh16)
        e2e_formats["bf16"] = {"value": N_total / (ms16 / 1000.0), "unit": "images/s", "ms_per_step": ms16,
                               "h2d_bytes_per_step": int(fh16.numel() * 2) * world, "chunk_images": args.chunk if args.chunk > 0 else 4 * sms,
                               "captions_identical_to_fp32_features": float((ho16["tokens"] == host_out["tokens"]).all(dim=1).float().mean())}
        del fh16
        if world == 1:
            p24 = eng_mod.pack_p24_host(fh).pin_memory()
            ms24, ho24 = time_e2e(p24, dtype="p24", num_regions=L)
            e2e_formats["p24"] = {"value": N_total / (ms24 / 1000.0), "unit": "images/s", "ms_per_step": ms24,
                                  "h2d_bytes_per_step": int(p24.numel()),
                                  "captions_identical_to_fp32_features": float((ho24["tokens"] == host_out["tokens"]).all(dim=1).float().mean())}
            del p24

    # ---- hand-off done by the encoder: features resident as tiles (NCHW trunk output -> capdec_ingest_features)
    ingested = None
    if world == 1:
        trunk = feats.permute(0, 3, 1, 2).contiguous()          # what models/encoder.py:13 holds before its permute
        ing_ms = time_device(lambda: model.ingest(trunk), 3)
        tiles = model.ingest(trunk)
        del trunk
        o2 = model.beam_search(tiles, beam_size=BEAM, max_length=MAXLEN, crop=False)
        t_ms = time_device(lambda: model.beam_search(tiles, beam_size=BEAM, max_length=MAXLEN, crop=False), max(2, min(args.steps, 5)))
        ingested = {"value": B / (t_ms / 1000.0), "unit": "images/s", "ms_per_step": t_ms, "ingest_ms": ing_ms,
                    "ingest_gbs": (feats.numel() * 4 + feats.numel() * 5) / (ing_ms * 1e-3) / 1e9,
                    "captions_identical": bool(torch.equal(o2["tokens"], out["tokens"])),
                    "note": "decode of a tile set written by capdec_ingest_features from the NCHW trunk output; ingest_ms is that pass (read fp32, write p24 planes + lo operand + mean)"}
        del tiles, o2

    # ---- the other precision modes on the same workload (N=1 only).  tf32x3 = the fp32-class record: full --steps,
    # its own roofline and clocks
    other_modes, fp32_class = {}, None
    if world == 1 and not args.no_other_modes:
        for prec in ("tf32x3", "fp32", "bf16"):   # fp32 = the exact CUDA-core mode: the parity anchor on the GPU
            if prec == args.precision:
                continue
            m2, _ = legacy_weights(VOCAB, 0)
            m2.precision = prec
            m2 = m2.to(dev)
            e2 = m2._engine(dev)
            fn = lambda: m2.beam_search(feats, beam_size=BEAM, max_length=MAXLEN, crop=False)
            for _ in range(2):
                fn()
            if prec == "tf32x3":
                e2.stage_timing(True)
                with ClockSampler(local_rank) as c2:
                    ms2 = time_device(fn, args.steps)
                st2 = e2.stage_times()
                e2.stage_timing(False)
            else:
                ms2 = time_device(fn, 2)
            o2 = fn()
            same = float((o2["tokens"] == out["tokens"]).all(dim=1).float().mean().item())
            other_modes[prec] = {"value": B / (ms2 / 1000.0), "unit": "images/s", "ms_per_step": ms2,
                                 "captions_identical_to_headline_mode": same}
            if prec == "tf32x3":
                fp32_class = {"dtype": "tf32x3", "note": "3-term TF32 split on tcgen05 kind::tf32, fp32 region tiles: every operand keeps "
                              ">= 21 significant bits; the same-precision-class companion of the headline", "value": B / (ms2 / 1000.0),
                              "unit": "images/s", "ms_per_step": ms2, "steps": args.steps, "clocks": c2.summary(),
                              "roofline": attention_roofline(st2, B, "f32"), "stage_roofline": gemm_rooflines(st2, B, "tf32x3"),
                              "stage_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in st2.items()},
                              "captions_identical_to_headline_mode": same}
            del e2, m2, o2

    # ---- BASELINE configs[2..4] at full size (N=1 only)
    configs = None
    if world == 1 and not args.no_configs:
        configs = bench_other_configs(dev)

    if rank == 0:
        peak, peak_src = measured_peaks()
        # tile format the attention kernel streams: bf16 mode -> bf16 tiles (2 B/element), bf16x3 mode -> p24 planes
        # (3 B/element, csrc/common.cuh), every other mode -> the fp32 tiles
        tile_fmt = "bf16" if args.precision == "bf16" else \
            "p24" if args.precision == "bf16x3" and not os.environ.get("CAPDEC_NO_P24_TILES") else "f32"
        total_stage_ms = sum(v[0] for v in stage.values()) or 1.0
        rec = {
            "metric": "captioned images/sec (beam=5, max_len=20)", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32x3": "tf32x3", "bf16": "bf16", "bf16x3": "bf16x3", "tf32": "tf32"}[args.precision], "data": "synthetic",
            "config": {"workload": WORKLOAD, "total_images": N_total, "images_per_gpu": B, "beam": BEAM, "max_len": MAXLEN,
                       "vocab": VOCAB, "regions": L, "feature_dim": D,
                       "parallelism": f"image-sharded x{world}, all-gather of captions (capdec_b200.sharding.gather_captions)",
                       "api": "capdec_b200.Decoder.beam_search (drop-in for models/decoder.py::Decoder)",
                       "l2_policy": f"inputs ({feats.numel() * 4 / 1e9:.1f} GB/GPU) larger than L2, no flush"},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms,
            "clocks": clocks.summary(),
            "roofline": attention_roofline(stage, B, tile_fmt),
            "stage_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in stage.items()},
            "stage_share": {k: round(v[0] / total_stage_ms, 4) for k, v in stage.items()},
        }
        sr = gemm_rooflines(stage, B, args.precision)
        if sr:
            rec["stage_roofline"] = sr
            # beam reorder (gather_rows_kernel): per row read new h and c (2 x 4H bytes), write h into the gate operand
            # as fp32 and as its hi/lo bf16 pair, write c  -> 5 x 4H bytes per row; it runs out of L2 as much as out of
            # HBM (its sources were written by the kernel before it), so "frac" can exceed 1 against the HBM peak
            ga_ms, ga_n = stage["gather"]
            if ga_n:
                ga_bytes = B * BEAM * 5 * 4 * 512
                ga_ach = ga_bytes / (ga_ms / ga_n * 1e-3) / 1e9
                rec["stage_roofline"]["reorder"] = {"kernel": "gather_rows_kernel", "bound": "hbm", "achieved": ga_ach, "peak": peak,
                                                    "unit": "GB/s", "frac": ga_ach / peak, "algorithmic_bytes_per_launch": ga_bytes,
                                                    "avg_launch_ms": ga_ms / ga_n}
        if sharded_ok is not None:
            rec["sharded_equals_single_gpu"] = sharded_ok
        if weak:
            rec["weak_scaling"] = weak
        if e2e:
            rec["e2e"] = e2e
            rec["e2e_formats"] = e2e_formats
        if ingested:
            rec["ingested"] = ingested
        if other_modes:
            rec["other_precision_modes"] = other_modes
        if fp32_class:
            rec["fp32_class"] = fp32_class
        if configs:
            rec["configs"] = configs
        if world == 1 and not args.no_cpu_baseline:
            n = min(env_int("CAPDEC_BENCH_CPU_IMAGES", 256), B)
            cpu_decode(2)
            ips, dt, ref = cpu_decode(n)
            rec["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"first {n} images x beam {BEAM} x {STEPS_PER_DECODE} steps, {dt:.1f} s, torch "
                                             f"{torch.__version__} fp32, {torch.get_num_threads()} threads"}
            # the GPU captions of those same images against the oracle's (north star: >= 99 % identical beams, 1e-3)
            g_tok, g_sc = out["tokens"][:n].cpu().long(), out["scores"][:n].cpu()
            same = (g_tok == ref["sequences"]).all(dim=1)
            rec["parity_in_run"] = {"n": n, "identical_frac": float(same.float().mean()),
                                    "max_dlogp": float((g_sc - ref["scores"])[same].abs().max()) if bool(same.any()) else None,
                                    "what": "best beam of the headline-mode GPU decode vs the fp32 CPU oracle on the same images; "
                                            "max_dlogp = max |score difference| over the identical beams (score = sum of log-probs / length)"}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(rec) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def bench_other_configs(dev):
    """BASELINE configs[2..4] at their full sizes on one GPU (device-resident inputs, CUDA events, 3 timed decodes after 2
    warm-ups).  Ceilings from SURVEY 8(d): C3 is HBM-bound on the cross-attention K/V (6 layers x 2 x 196 x 768 fp32 per
    image-step); C4 / C5 are tensor-bound (247 MFLOP per row-step, bf16)."""
    from tests.helpers import gpt2_decoder, transformer_decoder
    hbm, _ = measured_peaks()
    tpeak, _ = measured_tensor_peak()
    res = {}

    def rnd(shape, seed):
        return torch.randn(*shape, generator=torch.Generator(device=dev).manual_seed(seed), device=dev)

    def run(fn):
        for _ in range(2):
            fn()
        return time_device(fn, 3)
    m, _ = transformer_decoder(H=768, layers=6, heads=8, V=10000, max_length=50)
    m.precision = "bf16x3"
    m = m.to(dev)
    ef = {"features": rnd((2048, 196, 768), 5)}
    ms = run(lambda: m.generate(ef, 20, num_beams=3))
    c3_bytes = 2048 * (19 * 6 * 2 * 196 * 768 * 4)
    res["c3"] = {"workload": "configs[2]: ViT-B/16 features + 6-layer transformer decoder, beam 3, KV-cached, 2048 images, bf16x3",
                 "value": 2048 / ms * 1e3, "unit": "images/s", "ms": ms,
                 "roofline": {"bound": "hbm", "achieved": c3_bytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                              "frac": c3_bytes / (ms * 1e-3) / 1e9 / hbm,
                              "algorithmic": "19 steps x 6 layers x (K + V) x 196 x 768 fp32 per image (SURVEY 8(d): 7.23 MB per image-step)"}}
    del m, ef
    m, _ = gpt2_decoder(H=768, layers=12, heads=12, V=50257, max_length=64)
    m.precision = "bf16"
    m = m.to(dev)
    ef = {"pooled_features": rnd((1024, 768), 6)}
    ms = run(lambda: m.generate(ef, 20, num_beams=5))
    flop_row_step = 12 * 2 * (3 * 768 * 768 + 768 * 768 + 2 * 768 * 3072) + 2 * 768 * 50257
    c4_flop = 1024 * 5 * 19 * flop_row_step
    res["c4"] = {"workload": "configs[3]: GPT-2 124M, beam 5, vocab 50257, 1024 images, bf16", "value": 1024 / ms * 1e3,
                 "unit": "images/s", "ms": ms,
                 "roofline": {"bound": "tensor", "achieved": c4_flop / (ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                              "frac": c4_flop / (ms * 1e-3) / 1e12 / tpeak, "algorithmic": "247 MFLOP per row-step (SURVEY 8(d))"}}
    ef = {"pooled_features": rnd((512, 768), 7)}
    u = torch.rand(512 * 6, 19, device=dev)
    ms = run(lambda: m.generate(ef, 20, do_sample=True, num_samples=5, with_greedy=True, uniforms=u))
    c5_flop = 512 * 6 * 19 * flop_row_step
    res["c5"] = {"workload": "configs[4]: SCST rollout, GPT-2 124M, 5 samples + 1 greedy row per image, 512 images, bf16",
                 "value": 512 / ms * 1e3, "unit": "images/s", "ms": ms,
                 "roofline": {"bound": "tensor", "achieved": c5_flop / (ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s",
                              "frac": c5_flop / (ms * 1e-3) / 1e12 / tpeak, "algorithmic": "247 MFLOP per row-step (SURVEY 8(d))"}}
    return res


if __name__ == "__main__":
    main()
