#!/usr/bin/env bash
# Round-2 evidence run on the GPU box (under gpurun, ONE GPU): plain run first, then the ncu launch list of the same command
# and --set full captures of the dominant kernels; summaries are made on the box (the .ncu-rep files stay behind).
#   bash scripts/profile_r2.sh <tag>
set -u
TAG=${1:-r2}
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-modes --no-configs"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -c 300 gpurun_out/plain_$TAG.log; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:additive_attention_stream -s 60 -c 1 -f -o gpurun_out/attn_$TAG $CMD > gpurun_out/ncu_a_$TAG.log 2>&1
echo "attention capture rc=$?"
ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 131 -c 2 -f -o gpurun_out/gemm_$TAG $CMD > gpurun_out/ncu_g_$TAG.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none -k regex:ingest_ -s 2 -c 2 -f -o gpurun_out/ingest_$TAG $CMD > gpurun_out/ncu_i_$TAG.log 2>&1
echo "ingest capture rc=$?"
ncu --set full --clock-control none -k regex:beam_step -s 30 -c 1 -f -o gpurun_out/beam_$TAG $CMD > gpurun_out/ncu_b_$TAG.log 2>&1
echo "beam_step capture rc=$?"
# stream-K gate GEMM at the strong-scaling per-GPU size (512 images = 2560 rows)
CMD5="python bench.py --images 512 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-modes --no-configs"
ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 131 -c 2 -f -o gpurun_out/gemm512_$TAG $CMD5 > gpurun_out/ncu_g5_$TAG.log 2>&1
echo "gemm (512 images, stream-K) capture rc=$?"
# configs[3] (GPT-2 124M) launch list
ncu --metrics gpu__time_duration.sum --clock-control none -c 3600 --csv --log-file gpurun_out/launches_c4_$TAG.csv python scripts/bench_configs.py c4only > gpurun_out/ncu_c4_$TAG.log 2>&1
echo "c4 launch list rc=$?"
python scripts/summarize_ncu.py launches gpurun_out/launches_$TAG.csv gpurun_out/${TAG}_launches_summary.md > /dev/null
python scripts/summarize_ncu.py launches gpurun_out/launches_c4_$TAG.csv gpurun_out/${TAG}_c4_gpt2_launches_summary.md > /dev/null
for k in attn gemm ingest beam gemm512; do
  python scripts/summarize_ncu.py full gpurun_out/${k}_$TAG.ncu-rep gpurun_out/${TAG}_${k}_full.md > /dev/null
done
rm -f gpurun_out/*_$TAG.ncu-rep gpurun_out/launches_c4_$TAG.csv
ls -la gpurun_out | tail -14
