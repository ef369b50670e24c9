#!/bin/bash
# Builds library variants with different -D tuning flags into csrc/variants/ (git-ignored; they travel with gpurun).
# usage: bash profiles/build_variants.sh name1 "flags1" name2 "flags2" ...
cd "$(dirname "$0")/../nvalchemi-toolkit-ops_b200/csrc"
mkdir -p variants
pids=()
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags -o variants/lib_$name.so nvnl_api.cu 2> variants/$name.log && echo "built $name" || echo "FAILED $name" ) &
  pids+=($!)
  if [ ${#pids[@]} -ge 6 ]; then wait ${pids[0]}; pids=("${pids[@]:1}"); fi
done
wait
