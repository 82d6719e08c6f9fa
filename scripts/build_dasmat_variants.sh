#!/bin/sh
# tuning builds of the translated-plant unit only (launch bounds of the stand-alone step kernel); linked against the current objects
# usage: scripts/build_dasmat_variants.sh NAME THREADS MIN_BLOCKS [extra -D flags]
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
N=$1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off \
  -DRL4_DASMAT_THREADS=$2 -DRL4_DASMAT_MIN_BLOCKS=$3 $4 $5 -Xptxas -v -c -o build_variants/dasmat_$N.o rl4afcs_b200/csrc/dasmat_plant.cu > build_variants/dasmat_$N.log 2>&1
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build_variants/librl4_$N.so rl4afcs_b200/build/runtime.o rl4afcs_b200/build/sp_kernels.o \
  rl4afcs_b200/build/nl_kernels.o rl4afcs_b200/build/step_kernels.o rl4afcs_b200/build/host_episode.o build_variants/dasmat_$N.o
echo built $N
