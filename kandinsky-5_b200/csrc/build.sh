#!/bin/bash
# Builds libk5.so (sm_100a only) next to this script's parent: kandinsky-5_b200/libk5.so
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p ../build
pids=()
for f in common gemm attention rowops nabla engine conv3d vae_ops vae api; do
  if [ ! -f ../build/$f.o ] || [ $f.cu -nt ../build/$f.o ] || [ -n "$(find . -name '*.h' -newer ../build/$f.o -o -name '*.cuh' -newer ../build/$f.o)" ] || [ ../../include/k5.h -nt ../build/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o ../build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -shared -o ../libk5.so ../build/common.o ../build/gemm.o ../build/attention.o ../build/rowops.o ../build/nabla.o ../build/engine.o ../build/conv3d.o ../build/vae_ops.o ../build/vae.o ../build/api.o -gencode arch=compute_100a,code=sm_100a
echo "built $(cd .. && pwd)/libk5.so"
