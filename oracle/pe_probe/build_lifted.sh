#!/bin/sh
# TEST INFRASTRUCTURE -- translate the reference's plant binary (build container only) and compile the CPU library
set -e
cd "$(dirname "$0")"
V=${1:-extended_input}
python ../../rl4afcs_b200/tools/lift_plant.py --variant "$V" --out ../_ref/lifted
gcc -O1 -fPIC -shared -ffp-contract=off -Wall -Wno-unused-label -Wno-unused-function \
    -DLIFT_GENERATED_INC="\"../_ref/lifted/citation_${V}_code.inc\"" -o ../_ref/libcitation_lifted_${V}.so citation_lifted.c -lm
