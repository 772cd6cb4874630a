#!/bin/bash
# Builds libk5.so (sm_100a only) next to this script's parent: kandinsky-5_b200/libk5.so
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${K5_EXTRA_FLAGS:-}"
B=../build${K5_SUFFIX:-}
OUT=../libk5${K5_SUFFIX:-}.so
mkdir -p $B
pids=()
for f in common gemm attention rowops nabla engine conv3d vae_ops vae api; do
  if [ ! -f $B/$f.o ] || [ $f.cu -nt $B/$f.o ] || [ -n "$(find . -name '*.h' -newer $B/$f.o -o -name '*.cuh' -newer $B/$f.o)" ] || [ ../../include/k5.h -nt $B/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o $B/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -shared -o $OUT $B/common.o $B/gemm.o $B/attention.o $B/rowops.o $B/nabla.o $B/engine.o $B/conv3d.o $B/vae_ops.o $B/vae.o $B/api.o -gencode arch=compute_100a,code=sm_100a
echo "built $(cd .. && pwd)/$(basename $OUT)"
