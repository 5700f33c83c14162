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


def test_reference_arm_nonzero_rank_is_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert out.strip() == ""
