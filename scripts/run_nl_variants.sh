#!/bin/bash
# times a script with every build_variants/librl4_*.so (tuning experiments; RL4AFCS_LIB selects the library)
# usage: scripts/run_nl_variants.sh [script.py] [args...]     default script: scripts/prof_nl.py --steps 300 --warmup 50
cd "$(dirname "$0")/.."
script=scripts/prof_nl.py
if [[ "$1" == *.py ]]; then script=$1; shift; fi
for so in build_variants/librl4_*.so; do
  echo "== $so"
  if [ "$script" = scripts/prof_nl.py ] && [ $# -eq 0 ]; then set -- --steps 300 --warmup 50; fi
  RL4AFCS_LIB=$PWD/$so timeout 90 python $script "$@" 2>&1 | tail -6
done
