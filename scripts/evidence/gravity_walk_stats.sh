#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/gpu_stats.py 2e5 2 stats > gpurun_out/r3b_stats_small.log 2>&1; echo "small rc=$?"
tail -n 4 gpurun_out/r3b_stats_small.log
timeout 300 python scripts/gpu_stats.py 16e6 3 stats > gpurun_out/r3b_stats.log 2>&1; echo "rc=$?"
cat gpurun_out/r3b_stats.log
