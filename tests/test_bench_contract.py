"""bench.py contract checks that need no GPU: the reference arm (the oracle port timed on host cores) prints ONE JSON
line with the keys the driver reads, and under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ, CAPDEC_BENCH_CPU_IMAGES="8", **extra_env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        *args], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line():
    out = _run({})
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["higher_is_better"] is True and rec["unit"] == "images/s"
    assert rec["metric"] == "captioned images/sec (beam=5, max_len=20)" and rec["n_gpus"] == 1
    assert rec["value"] > 0 and rec["ms_per_step"] > 0 and rec["vs_baseline"] is None and rec["data"] == "synthetic"
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1
    assert rec["cpu_baseline"]["value"] == rec["value"]
    assert rec["e2e"] == {"value": rec["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in rec["config"]["workload"] and "model" not in rec["config"]
    assert rec["scaling"] == "strong" and rec["config"]["total_images"] == 4096


def test_reference_arm_never_loads_the_product_library():
    """the reference arm is the CPU oracle only: importing it must not pull in capdec_b200 / libcapdec.so"""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']; "
            "runpy.run_path('bench.py', run_name='__main__'); "
            "bad = [m for m in sys.modules if 'capdec' in m]; assert not bad, bad; "
            "assert 'libcapdec' not in open('/proc/self/maps').read()")
    env = dict(os.environ, CAPDEC_BENCH_CPU_IMAGES="4")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]


def test_global_features_are_shard_consistent():
    """any rank can generate exactly its shard of the global synthetic batch (strong scaling: 4096 images in total)"""
    import importlib
    import torch
    sys.path.insert(0, ROOT)
    try:
        bench = importlib.import_module("bench")
        full = bench.global_features(0, 600)
        assert torch.equal(bench.global_features(100, 300), full[100:300])
        assert torch.equal(bench.global_features(512, 600), full[512:600])
        assert full.shape == (600, 14, 14, 2048) and float(full.min()) == 0.0
    finally:
        sys.path.remove(ROOT)


def test_reference_arm_nonzero_rank_is_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert out.strip() == ""


def test_clock_sampler_nvml_path_with_a_stub(monkeypatch):
    """The in-process NVML sampler (no GPU here): a stub pynvml drives the same code path and the summary carries the
    keys the bench line reports."""
    import importlib
    import time
    import types
    stub = types.ModuleType("pynvml")
    stub.NVML_CLOCK_SM = 1
    stub.nvmlClocksThrottleReasonHwSlowdown = 0x8
    stub.nvmlClocksThrottleReasonHwThermalSlowdown = 0x40
    stub.nvmlClocksThrottleReasonSwThermalSlowdown = 0x20
    stub.nvmlClocksThrottleReasonSwPowerCap = 0x4
    stub.nvmlClocksThrottleReasonHwPowerBrakeSlowdown = 0x80
    stub.nvmlInit = lambda: None
    stub.nvmlDeviceGetHandleByUUID = lambda u: (_ for _ in ()).throw(RuntimeError("no uuid"))
    stub.nvmlDeviceGetHandleByIndex = lambda i: ("dev", i)
    stub.nvmlDeviceGetMaxClockInfo = lambda h, c: 1965
    stub.nvmlDeviceGetEnforcedPowerLimit = lambda h: 1000000
    stub.nvmlDeviceGetClockInfo = lambda h, c: 1700
    stub.nvmlDeviceGetCurrentClocksThrottleReasons = lambda h: 0x4
    stub.nvmlDeviceGetPowerUsage = lambda h: 990000
    monkeypatch.setitem(sys.modules, "pynvml", stub)
    sys.path.insert(0, ROOT)
    try:
        bench = importlib.import_module("bench")
        with bench.ClockSampler(0) as c:
            time.sleep(0.08)
        out = c.summary()
    finally:
        sys.path.remove(ROOT)
    assert out["source"] == "nvml" and out["samples"] >= 2 and out["sm_mhz"] == 1700.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["power_w"] == 990.0 and out["power_limit_w"] == 1000.0
