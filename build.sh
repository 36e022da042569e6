#!/usr/bin/env bash
# Build libfsgpu.so (sm_100a) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")/finetoolsflexstructures.jl_b200"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -Wno-deprecated-declarations"
mkdir -p build
pids=()
for f in fsgpu_core fsgpu_elements fsgpu_explicit fsgpu_tile; do
  $NVCC $FLAGS "$@" -c csrc/$f.cu -o build/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -Xcompiler -pthread -gencode arch=compute_100a,code=sm_100a -o libfsgpu.so build/fsgpu_core.o build/fsgpu_elements.o build/fsgpu_explicit.o build/fsgpu_tile.o
echo "built $(pwd)/libfsgpu.so"
