#!/bin/bash
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --durations=5 -k "far_field or 10k_steps or accretion") > gpurun_out/r3c_pytest.log 2>&1
tail -n 12 gpurun_out/r3c_pytest.log
timeout 300 python scripts/gpu_stats.py 16e6 4 > gpurun_out/r3c_perf.log 2>&1; echo "rc=$?"
cat gpurun_out/r3c_perf.log
