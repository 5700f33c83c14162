#!/usr/bin/env bash
# Experiment helper: build a variant of ONE translation unit with extra flags and link it against the product's other
# objects into _exp/<name>.so (git-ignored; travels to the GPU box).  A/B runs copy it over csrc/libcapdec.so on the box.
#   bash scripts/build_variant.sh <name> <unit> "<extra nvcc flags>"
set -euo pipefail
NAME=$1; UNIT=$2; FLAGS=${3:-}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CSRC="$ROOT/image-captioning-ml-project_b200/csrc"
mkdir -p "$ROOT/_exp"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $FLAGS -I"$CSRC" -c "$CSRC/$UNIT.cu" -o "$ROOT/_exp/$NAME.$UNIT.o"
OBJS=$(ls "$CSRC"/_build/*.o | grep -v "/$UNIT.o")
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$ROOT/_exp/$NAME.so" $OBJS "$ROOT/_exp/$NAME.$UNIT.o" -lcudart
rm -f "$ROOT/_exp/$NAME.$UNIT.o"
echo "built _exp/$NAME.so"
