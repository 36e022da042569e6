#!/usr/bin/env bash
# Build variants/libfsgpu_NAME.so with extra nvcc flags applied to ONE translation unit (default fsgpu_elements),
# reusing the other objects of the last ./build.sh.  usage: scripts/build_variant.sh NAME [-DFLAG ...] [--unit fsgpu_tile]
set -euo pipefail
cd "$(dirname "$0")/../finetoolsflexstructures.jl_b200"
name=$1; shift
unit=fsgpu_elements
args=()
while [ $# -gt 0 ]; do
  if [ "$1" = "--unit" ]; then unit=$2; shift 2; else args+=("$1"); shift; fi
done
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -pthread -Wno-deprecated-declarations"
mkdir -p ../variants build
$NVCC $FLAGS "${args[@]}" -c csrc/$unit.cu -o build/${unit}_$name.o
objs=""
for f in fsgpu_core fsgpu_elements fsgpu_explicit fsgpu_tile; do
  if [ $f = $unit ]; then objs="$objs build/${unit}_$name.o"; else objs="$objs build/$f.o"; fi
done
$NVCC -shared -Xcompiler -pthread -gencode arch=compute_100a,code=sm_100a -o ../variants/libfsgpu_$name.so $objs
echo "built variants/libfsgpu_$name.so"
