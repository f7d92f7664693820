#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -x --durations=3 -m gpu) > gpurun_out/r2k_pytest.log 2>&1
tail -4 gpurun_out/r2k_pytest.log
scripts/dd_scale.sh r2k 2000000 3 1 > gpurun_out/r2k_2M.log 2>&1
scripts/dd_scale.sh r2k 16000000 3 1 > gpurun_out/r2k_16M.log 2>&1
grep "last step" gpurun_out/r2k_2M.log gpurun_out/r2k_16M.log
