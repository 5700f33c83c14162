#!/usr/bin/env bash
# Round-2 final evidence run on the GPU box (under gpurun, ONE GPU): plain runs first, then ncu launch lists of the same
# commands and --set full captures of the dominant kernels; summaries are made on the box (the .ncu-rep files stay behind).
#   bash scripts/profile_r2f.sh <tag>
set -u
TAG=${1:-r2f}
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-other-modes --no-configs"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -c 300 gpurun_out/plain_$TAG.log; echo
python scripts/bench_configs.py c3 c4only > gpurun_out/plain_cfg_$TAG.log 2>&1 || { echo "config run failed"; tail -5 gpurun_out/plain_cfg_$TAG.log; exit 1; }
cat gpurun_out/plain_cfg_$TAG.log
# headline config: launch list + the three dominant kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:additive_attention_stream -s 60 -c 1 -f -o gpurun_out/attn_$TAG $CMD > gpurun_out/ncu_a_$TAG.log 2>&1
echo "attention capture rc=$?"
ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 131 -c 2 -f -o gpurun_out/gemm_$TAG $CMD > gpurun_out/ncu_g_$TAG.log 2>&1
echo "gemm capture rc=$?"
# configs[3] (GPT-2 124M) and configs[2] (transformer): launch lists; one self-attention launch and one layer's four GEMMs in full
ncu --metrics gpu__time_duration.sum --clock-control none -c 3600 --csv --log-file gpurun_out/launches_c4_$TAG.csv python scripts/bench_configs.py c4only > gpurun_out/ncu_c4_$TAG.log 2>&1
echo "c4 launch list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/launches_c3_$TAG.csv python scripts/bench_configs.py c3 > gpurun_out/ncu_c3_$TAG.log 2>&1
echo "c3 launch list rc=$?"
ncu --set full --clock-control none -k regex:self_attn_decode3 -s 300 -c 1 -f -o gpurun_out/selfattn_$TAG python scripts/bench_configs.py c4only > gpurun_out/ncu_sa_$TAG.log 2>&1
echo "self-attention capture rc=$?"
ncu --set full --clock-control none --cache-control none -k regex:gemm_tcgen05 -s 1200 -c 4 -f -o gpurun_out/gemmc4_$TAG python scripts/bench_configs.py c4only > gpurun_out/ncu_gc4_$TAG.log 2>&1
echo "GPT-2 GEMM capture rc=$?"
python scripts/summarize_ncu.py launches gpurun_out/launches_$TAG.csv gpurun_out/${TAG}_launches_summary.md > /dev/null
python scripts/summarize_ncu.py launches gpurun_out/launches_c4_$TAG.csv gpurun_out/${TAG}_c4_gpt2_launches_summary.md > /dev/null
python scripts/summarize_ncu.py launches gpurun_out/launches_c3_$TAG.csv gpurun_out/${TAG}_c3_transformer_launches_summary.md > /dev/null
for k in attn gemm selfattn gemmc4; do
  python scripts/summarize_ncu.py full gpurun_out/${k}_$TAG.ncu-rep gpurun_out/${TAG}_${k}_full.md > /dev/null
done
rm -f gpurun_out/*_$TAG.ncu-rep gpurun_out/launches_c4_$TAG.csv gpurun_out/launches_c3_$TAG.csv
ls -la gpurun_out | tail -12
