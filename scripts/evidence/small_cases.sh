#!/bin/bash
# step times of the small configurations (launch / sync bound) on the final build
mkdir -p gpurun_out
for n in 1e4 1e5 1e6; do timeout 60 python scripts/gpu_stats.py $n 6 2>&1 | grep "^step" | tail -n 2; done > gpurun_out/r2b_small_cases.log
cat gpurun_out/r2b_small_cases.log
