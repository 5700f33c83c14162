#!/usr/bin/env bash
# Build libcapdec.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall ${CAPDEC_NVCC_FLAGS:-}"
mkdir -p _build
pids=()
for f in capdec gemm_ffma gemm_tc attn_additive attn_stream attn_mha attn_mha_stream select transformer ingest; do
  if [ ! -f _build/$f.o ] || [ $f.cu -nt _build/$f.o ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer _build/$f.o)" ] || [ ../../include/capdec.h -nt _build/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o _build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libcapdec.so _build/*.o -lcudart
echo "built $(pwd)/libcapdec.so"
