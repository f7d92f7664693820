#!/bin/bash
# Developer helper (GPU box): time the default library and every variant under summersph_b200/variants at N particles.
# usage: scripts/gpu_variants.sh N STEPS [variant names...]
N=${1:-4e6}; STEPS=${2:-2}; shift; shift
echo "== default"; python scripts/gpu_perf.py $N $STEPS 2>&1 | grep -A1 "^step" | grep -v "^--"
for v in "$@"; do
  echo "== $v"; SPH_B200_LIB=summersph_b200/variants/libsph_$v.so python scripts/gpu_perf.py $N $STEPS 2>&1 | grep -A1 "^step" | grep -v "^--"
done
