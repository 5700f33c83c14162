"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches_summary.md [skip_first]
  python scripts/summarize_ncu.py full gpurun_out/attn_r1.ncu-rep profiles/r1_attention_full.md
"""
import collections
import csv
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"capdec::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)
    return name[:110]


def launches(path, out, skip=0):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    rows = [r for r in rows if r["Metric Name"] == "gpu__time_duration.sum"]
    all_rows = rows
    rows = rows[skip:]
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over the bench command; per-launch times are "
                "cold-cache and serialised, so compare SHARES, not absolutes.\n\n")
        f.write(f"launches profiled: {len(all_rows)} (summary skips the first {skip}: weight packing + warm-up), "
                f"summed kernel time {total:.2f} ms\n\n| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ms:.3f} | {1000 * ms / n:.1f} | {100 * ms / total:.1f}% |\n")
    print(open(out).read())


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum", "lts__t_sector_hit_rate.pct",
           "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "dram__cycles_active.avg",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
           "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    header, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\n`ncu --set full --clock-control none --import-source on`; values per launch.\n\n")
        for d in data:
            rec = dict(zip(header, d))
            f.write(f"## launch {rec.get('ID')}: `{short(rec.get('Kernel Name', ''))}` grid {rec.get('Grid Size')} block {rec.get('Block Size')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in METRICS:
                if m in rec:
                    f.write(f"| {m} | {rec[m]} | {units[header.index(m)]} |\n")
            f.write("\n")
    print(open(out).read()[:6000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        full(sys.argv[2], sys.argv[3])
