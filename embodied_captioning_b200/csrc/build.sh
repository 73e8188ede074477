#!/usr/bin/env bash
# Builds libec_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${EC_NVCC_EXTRA:-}"
mkdir -p build
pids=()
rm -f build/gemm_sm100.o
for f in common gemm_host layernorm attention attention_sm100 decode cross_absorbed preprocess disagreement rasterize textenc model api; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ -n "$(find . -maxdepth 1 -name '*.cuh' -newer build/$f.o)" ] \
     || [ ../../include/ec_b200.h -nt build/$f.o ]; then
    ( $NVCC $FLAGS -c $f.cu -o build/$f.o || { rm -f build/$f.o; exit 1; } ) &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]:-}"; do if [ -n "$p" ]; then wait "$p" || rc=1; fi; done
[ $rc -eq 0 ] || { echo "compile failed" >&2; exit 1; }
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o ${EC_LIB_OUT:-../libec_b200.so} build/*.o -lcudart_static -ldl -lrt -lpthread
echo "built $(realpath ${EC_LIB_OUT:-../libec_b200.so})"
