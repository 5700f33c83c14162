#!/usr/bin/env bash
# Run on the GPU box (under gpurun): plain run first, then the ncu launch list and --set full captures of the
# dominant kernels.  Outputs land in gpurun_out/; scripts/summarize_ncu.py turns them into profiles/*.md.
#   bash scripts/profile.sh <tag>
set -u
TAG=${1:-r1b}
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-modes"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -c 600 gpurun_out/plain_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:additive_attention_stream -s 60 -c 1 -f -o gpurun_out/attn_$TAG $CMD > gpurun_out/ncu_a_$TAG.log 2>&1
echo "attention capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 131 -c 2 -f -o gpurun_out/gemm_$TAG $CMD > gpurun_out/ncu_g_$TAG.log 2>&1
echo "gemm capture rc=$?"
ls -la gpurun_out | tail -8
