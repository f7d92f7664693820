#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -x --durations=5 -m gpu) > gpurun_out/r2j_pytest.log 2>&1
tail -5 gpurun_out/r2j_pytest.log
scripts/dd_scale.sh r2j 2000000 3 1 > gpurun_out/r2j_2M.log 2>&1
scripts/dd_scale.sh r2j 16000000 3 1 > gpurun_out/r2j_16M.log 2>&1
grep "last step" gpurun_out/r2j_2M.log gpurun_out/r2j_16M.log
