#!/usr/bin/env bash
# Run on the GPU box (under gpurun): plain run first, then the ncu launch list and --set full captures of the
# dominant kernels.  Outputs land in gpurun_out/; scripts/summarize_ncu.py turns them into profiles/*.md.
#   bash scripts/profile.sh <tag>
set -u
TAG=${1:-r1b}
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-modes"
mkdir -p gpurun_out
if [ -z "${SKIP_PLAIN:-}" ]; then
  $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
  tail -c 600 gpurun_out/plain_$TAG.log
fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:additive_attention_stream -s 60 -c 1 -f -o gpurun_out/attn_$TAG $CMD > gpurun_out/ncu_a_$TAG.log 2>&1
echo "attention capture rc=$?"
ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 131 -c 2 -f -o gpurun_out/gemm_$TAG $CMD > gpurun_out/ncu_g_$TAG.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none -k regex:gather_rows -s 30 -c 1 -f -o gpurun_out/gather_$TAG $CMD > gpurun_out/ncu_r_$TAG.log 2>&1
echo "reorder capture rc=$?"
# summaries are made here on the box: gpurun only returns gpurun_out/ when it is under 64 MiB, so the big reports stay behind
python scripts/summarize_ncu.py launches gpurun_out/launches_$TAG.csv gpurun_out/${TAG}_launches_summary.md > /dev/null
python scripts/summarize_ncu.py full gpurun_out/attn_$TAG.ncu-rep gpurun_out/${TAG}_attention_stream_full.md > /dev/null
python scripts/summarize_ncu.py full gpurun_out/gemm_$TAG.ncu-rep gpurun_out/${TAG}_gemm_pair_full.md > /dev/null
python scripts/summarize_ncu.py full gpurun_out/gather_$TAG.ncu-rep gpurun_out/${TAG}_reorder_full.md > /dev/null
rm -f gpurun_out/gemm_$TAG.ncu-rep
ls -la gpurun_out | tail -12
