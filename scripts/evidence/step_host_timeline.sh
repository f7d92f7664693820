#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gpu_io.py 16e6 3 > gpurun_out/r3f_io.log 2>&1; echo "rc=$?"
cat gpurun_out/r3f_io.log
