#!/usr/bin/env python
"""bench.py -- captioned images/s of the batched beam-search decode hot path (BASELINE.json metric).

Workload (BASELINE.json configs[1]): legacy ResNet-101 + LSTM + soft attention (models/decoder.py),
beam 5, max_len 20, vocab 10k, 4096 images of synthetic 14x14x2048 features PER GPU (weak scaling: each rank
decodes its own shard, NCCL all-gathers the finished captions + scores; no data-path collective).
A "step" is one full decode of the batch (prologue + 19 beam steps).

  value        images/s, device-resident features, CUDA events on the launching stream, max over ranks
  e2e          the same through the host-buffer C-ABI entry point capdec_decode_beam_host (pinned host
               features copied host->device inside the timed region, captions copied back)
  roofline     attention kernel: algorithmic bytes / live CUDA-event time of that kernel vs measured HBM peak
  cpu_baseline the oracle port of the reference (torch fp32, all host threads) on a bounded sample

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


def cpu_decode_images_per_s(n_images, repeats=1):
    """The reference's CPU path (oracle port), fp32, every host thread, decode only."""
    from oracle import beam as obeam, legacy as olegacy
    from tests.helpers import legacy_features, legacy_weights
    torch.set_num_threads(os.cpu_count())
    _, sd = legacy_weights(VOCAB, 0)
    enc = legacy_features(n_images)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            obeam.beam_search(olegacy.LegacyStepper(sd, enc, BEAM), n_images, BEAM, MAXLEN)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_images / best, best


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    n = env_int("CAPDEC_BENCH_CPU_IMAGES", 64)
    cpu_decode_images_per_s(min(n, 4))  # warm-up (thread pools, first-touch)
    times = []
    for _ in range(args.warmup):
        cpu_decode_images_per_s(n)
    for _ in range(args.steps):
        ips, dt = cpu_decode_images_per_s(n)
        times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = n / (ms / 1000.0)
    sample = f"{n} images x beam {BEAM} x {STEPS_PER_DECODE} steps per step (bounded sample of the {args.images}-image workload)"
    print(json.dumps({
        "impl": "reference", "metric": "captioned images/sec (beam=5, max_len=20)", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: ResNet-101 features + LSTM + soft attention, beam=5, max_len=20, vocab 10k",
                   "images_per_gpu": args.images, "sample_images": n},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="capdec", choices=["capdec", "reference"])
    ap.add_argument("--images", type=int, default=env_int("CAPDEC_BENCH_IMAGES", 4096), help="images per GPU")
    ap.add_argument("--precision", default=os.environ.get("CAPDEC_BENCH_PRECISION", "bf16x3"))
    ap.add_argument("--chunk", type=int, default=env_int("CAPDEC_BENCH_CHUNK", 0),
                    help="e2e H2D pipeline chunk in images (0 = the library default, two images per SM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-modes", action="store_true")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.images
    model, _ = legacy_weights(VOCAB, 0)
    model.precision = args.precision
    model = model.to(dev)
    eng = model._engine(dev)

    # synthetic 14x14x2048 region features, seeded per rank, generated on the host then copied;
    # 6.6 GB per GPU >> the 126 MB L2, so no flush is needed between timed iterations
    g = torch.Generator().manual_seed(1234 + rank)
    feats_host = torch.empty(B, L, D, dtype=torch.float32).pin_memory() if not args.no_e2e else torch.empty(B, L, D)
    blk = 256
    for i in range(0, B, blk):
        n = min(blk, B - i)
        feats_host[i:i + n] = torch.relu(torch.randn(n, L, D, generator=g))
    feats = feats_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(out):
        """the path's only exchange: all-gather finished captions + scores (84 B/image)"""
        if world == 1:
            return
        toks = [torch.empty_like(out["tokens"]) for _ in range(world)]
        scs = [torch.empty_like(out["scores"]) for _ in range(world)]
        dist.all_gather(toks, out["tokens"])
        dist.all_gather(scs, out["scores"])

    def one_step():
        out = eng.decode_beam(feats, None, None, BEAM, MAXLEN)
        gather(out)
        return out

    for _ in range(max(args.warmup, 3)):
        out = one_step()
    barrier()

    # ---- timed region: device-resident features, CUDA events on the launching stream
    eng.stage_timing(True)
    launches0 = cd.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = one_step()
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = cd.launch_count() - launches0
    stage = eng.stage_times()
    eng.stage_timing(False)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = B * world / (ms_step / 1000.0)

    # ---- end to end: pinned host features -> C-ABI host entry point -> host captions
    e2e = None
    if not args.no_e2e:
        host_out = {"tokens": torch.empty(B, MAXLEN, dtype=torch.int32).pin_memory(),
                    "lengths": torch.empty(B, dtype=torch.int32).pin_memory(),
                    "scores": torch.empty(B, dtype=torch.float32).pin_memory()}
        for _ in range(2):
            eng.decode_beam_host(feats_host, None, BEAM, MAXLEN, chunk_images=args.chunk, out=host_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.decode_beam_host(feats_host, None, BEAM, MAXLEN, chunk_images=args.chunk, out=host_out)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_ms = 1000.0 * float(dt.item()) / args.steps
        same = bool(torch.equal(host_out["tokens"], out["tokens"].cpu()))
        e2e = {"value": B * world / (e2e_ms / 1000.0), "unit": "images/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(feats_host.numel() * 4),
               "d2h_bytes_per_step": int(B * MAXLEN * 4 + B * 8),
               "chunk_images": args.chunk if args.chunk > 0 else 2 * torch.cuda.get_device_properties(dev).multi_processor_count,
               "matches_device_path": same}

    # ---- the other tensor-core modes on the same workload (reported beside the headline, N=1 only)
    other_modes = {}
    if world == 1 and not args.no_other_modes:
        for prec in ("fp32", "tf32x3", "bf16x3", "bf16"):   # fp32 = the exact CUDA-core mode: the parity anchor on the GPU
            if prec == args.precision:
                continue
            m2, _ = legacy_weights(VOCAB, 0)
            m2.precision = prec
            e2 = m2.to(dev)._engine(dev)
            for _ in range(2):
                e2.decode_beam(feats, None, None, BEAM, MAXLEN)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a0.record()
            for _ in range(2):
                o2 = e2.decode_beam(feats, None, None, BEAM, MAXLEN)
            a1.record()
            torch.cuda.synchronize()
            ms2 = a0.elapsed_time(a1) / 2
            same = float((o2["tokens"] == out["tokens"]).all(dim=1).float().mean().item())
            other_modes[prec] = {"value": B / (ms2 / 1000.0), "unit": "images/s", "ms_per_step": ms2,
                                 "captions_identical_to_headline_mode": same}
            del e2, m2

    if rank == 0:
        peak, peak_src = measured_peaks()
        att_ms, att_n = stage["attention"]
        per_launch_ms = att_ms / max(att_n, 1)
        # the bf16 mode streams bf16 tiles (half the algorithmic bytes); every other mode streams the fp32 tiles
        # tile format the attention kernel streams: bf16 mode -> bf16 tiles (2 B/element), bf16x3 mode -> p24 planes
        # (3 B/element, csrc/common.cuh), every other mode -> the fp32 tiles
        tile_fmt = "bf16" if args.precision == "bf16" else \
            "p24" if args.precision == "bf16x3" and not os.environ.get("CAPDEC_NO_P24_TILES") else "f32"
        attn_bytes = ATTN_BYTES_PER_IMAGE_STEP * {"bf16": 2, "p24": 3, "f32": 4}[tile_fmt] // 4
        achieved = attn_bytes * B / (per_launch_ms * 1e-3) / 1e9 if att_n else None
        traffic = None
        tf = os.path.join(ROOT, "profiles", "attention_traffic.json")
        if os.path.isfile(tf):
            tj = json.load(open(tf))
            traffic = tj.get(tile_fmt, {}).get("dram_bytes_per_launch") if isinstance(tj.get(tile_fmt), dict) else None
        total_stage_ms = sum(v[0] for v in stage.values()) or 1.0
        rec = {
            "metric": "captioned images/sec (beam=5, max_len=20)", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32x3": "tf32x3", "bf16": "bf16", "bf16x3": "bf16x3", "tf32": "tf32"}[args.precision], "data": "synthetic",
            "config": {"workload": "configs[1]: ResNet-101 features + LSTM + soft attention, beam=5, max_len=20, vocab 10k",
                       "images_per_gpu": B, "beam": BEAM, "max_len": MAXLEN, "vocab": VOCAB, "regions": L,
                       "feature_dim": D, "parallelism": f"image-sharded x{world}, all-gather of captions",
                       "l2_policy": "inputs (6.6 GB/GPU) larger than L2, no flush"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": {"kernel": "additive_attention_stream_kernel<5,relu,2,%s>" % tile_fmt, "bound": "hbm", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "peak_source": peak_src, "tile_bytes_per_element": {"bf16": 2, "p24": 3, "f32": 4}[tile_fmt],
                         "algorithmic_bytes_per_launch": attn_bytes * B,
                         "avg_launch_ms": per_launch_ms, "launches": att_n},
            "stage_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in stage.items()},
            "stage_share": {k: round(v[0] / total_stage_ms, 4) for k, v in stage.items()},
        }
        # tensor-bound stages: MMA FLOPs actually issued (terms x 2MNK; a TF32 MMA counts double against the bf16 peak)
        terms = {"fp32": 0, "tf32x3": 3, "bf16x3": 3, "bf16": 1, "tf32": 1}[args.precision]
        if terms:
            tpeak, tsrc = measured_tensor_peak()
            scale = 2.0 if args.precision.startswith("tf32") else 1.0
            R, H4, Kg = B * BEAM, 4 * 512, 2048 + 512
            Nv = (VOCAB + 255) // 256 * 256 + A + D
            def tf(ms_n, flop):
                ms, n = ms_n
                return None if not n else flop * terms * scale / (ms / n * 1e-3) / 1e12
            g_ach, v_ach = tf(stage["gate_gemm"], 2.0 * R * H4 * Kg), tf(stage["vocab_gemm"], 2.0 * R * Nv * 512)
            rec["stage_roofline"] = {
                "gate_gemm": {"bound": "tensor", "achieved": g_ach, "peak": tpeak, "unit": "TFLOP/s (bf16-equivalent MMA work)",
                              "frac": g_ach / tpeak if g_ach else None, "shape": [R, H4, Kg], "mma_terms": terms},
                "vocab_gemm": {"bound": "tensor", "achieved": v_ach, "peak": tpeak, "unit": "TFLOP/s (bf16-equivalent MMA work)",
                               "frac": v_ach / tpeak if v_ach else None, "shape": [R, Nv, 512], "mma_terms": terms,
                               "note": "vocabulary padded to 256 + the [dec_att|f_beta] tail; fused log-softmax/top-k epilogue"},
                "peak_source": tsrc}
            # beam reorder (gather_rows_kernel): per row read new h and c (2 x 4H bytes), write h into the gate operand
            # as fp32 and as its hi/lo bf16 pair, write c  -> 5 x 4H bytes per row; it runs out of L2 as much as out of
            # HBM (its sources were written by the kernel before it), so "frac" can exceed 1 against the HBM peak
            ga_ms, ga_n = stage["gather"]
            if ga_n:
                ga_bytes = R * 5 * 4 * 512
                ga_ach = ga_bytes / (ga_ms / ga_n * 1e-3) / 1e9
                rec["stage_roofline"]["reorder"] = {"kernel": "gather_rows_kernel", "bound": "hbm", "achieved": ga_ach, "peak": peak,
                                                    "unit": "GB/s", "frac": ga_ach / peak, "algorithmic_bytes_per_launch": ga_bytes,
                                                    "avg_launch_ms": ga_ms / ga_n}
        if e2e:
            rec["e2e"] = e2e
        if other_modes:
            rec["other_precision_modes"] = other_modes
        if world == 1 and not args.no_cpu_baseline:
            n = env_int("CAPDEC_BENCH_CPU_IMAGES", 256)
            cpu_decode_images_per_s(2)
            ips, dt = cpu_decode_images_per_s(n)
            rec["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"{n} images x beam {BEAM} x {STEPS_PER_DECODE} steps, {dt:.1f} s, torch "
                                             f"{torch.__version__} fp32, {torch.get_num_threads()} threads"}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(rec) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
