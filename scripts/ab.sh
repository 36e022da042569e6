#!/usr/bin/env bash
# A/B timing of library variants on the GPU box: scripts/ab.sh "<run_op args>" name1 name2 ...   ("base" = the shipped library)
args=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then lib=finetoolsflexstructures.jl_b200/libfsgpu.so; else lib=variants/libfsgpu_$v.so; fi
  echo "== $v: $(FSGPU_LIB=$PWD/$lib python scripts/run_op.py $args 2>&1 | tail -1)"
done
