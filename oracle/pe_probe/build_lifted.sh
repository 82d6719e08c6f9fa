#!/bin/sh
# TEST INFRASTRUCTURE -- translate the reference's plant binary (build container only) and compile the CPU library
set -e
cd "$(dirname "$0")"
V=${1:-extended_input}
python lift.py --variant "$V"
gcc -O1 -fPIC -shared -ffp-contract=off -Wall -Wno-unused-label -Wno-unused-function \
    -DLIFT_GENERATED_INC="\"../_ref/lifted/citation_${V}_code.inc\"" -o ../_ref/libcitation_lifted_${V}.so citation_lifted.c -lm
