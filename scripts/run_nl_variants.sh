#!/bin/bash
# times scripts/prof_nl.py with every build_variants/librl4_*.so (tuning experiments; RL4AFCS_LIB selects the library)
cd "$(dirname "$0")/.."
for so in build_variants/librl4_*.so; do
  echo "== $so"
  RL4AFCS_LIB=$PWD/$so timeout 300 python scripts/prof_nl.py --steps 300 --warmup 50 "$@" 2>&1 | tail -1
done
