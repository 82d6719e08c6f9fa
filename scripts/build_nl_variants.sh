#!/bin/bash
# Tuning experiments for the nonlinear kernel: builds build_variants/librl4_<name>.so with extra -D flags.
# usage: scripts/build_nl_variants.sh name1 "-DFLAG=1 ..." name2 "..." ...
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC"
C=rl4afcs_b200/csrc
mkdir -p build_variants
[ -f build_variants/runtime.o ] && [ build_variants/runtime.o -nt $C/runtime.cu ] || nvcc $F -c $C/runtime.cu -o build_variants/runtime.o &
[ -f build_variants/sp_kernels.o ] && [ build_variants/sp_kernels.o -nt $C/sp_kernels.cu ] && [ build_variants/sp_kernels.o -nt $C/sp_core.cuh ] && [ build_variants/sp_kernels.o -nt $C/rl4_math.cuh ] || nvcc $F -c $C/sp_kernels.cu -o build_variants/sp_kernels.o &
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc $F -Xptxas -v $flags -c $C/nl_kernels.cu -o build_variants/nl_$name.o > build_variants/nl_$name.log 2>&1 ) &
  names="$names $name"
done
wait
for name in $names; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/librl4_$name.so build_variants/runtime.o build_variants/sp_kernels.o build_variants/nl_$name.o
  echo "$name: $(grep -A2 'nl_run_kernelIfLi1ELb0' build_variants/nl_$name.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes stack frame, [0-9]* bytes spill stores, [0-9]* bytes spill loads' | tr '\n' ' ')"
done
