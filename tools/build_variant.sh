#!/bin/bash
# A/B builds of the library with extra nvcc flags:  tools/build_variant.sh <name> [-DMACRO=value ...]
# -> gpurun_variants/libreseq_b200_<name>.so (git-ignored, travels to the GPU box); use with RSQ_B200_LIB=<path>.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; shift
mkdir -p "$ROOT/gpurun_variants"
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 -shared \
  "$@" -o "$ROOT/gpurun_variants/libreseq_b200_$NAME.so" "$ROOT/reseq_b200/csrc/engine.cu" -lz
cuobjdump --dump-resource-usage "$ROOT/gpurun_variants/libreseq_b200_$NAME.so" 2>/dev/null | grep -A1 -E "Function.*k_spec_(scan|reads)" | grep -E "Function|REG" | paste - - | sed -E 's/.*Function ([^:]*):.*REG:([0-9]+) STACK:([0-9]+).*/REG \2 STACK \3  \1/' | cut -c1-80
