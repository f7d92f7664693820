#!/bin/bash
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=200
(time python -m pytest tests/test_domains.py tests/test_gpu_ics.py -q -x --durations=5) > gpurun_out/r2e_pytest.log 2>&1
tail -5 gpurun_out/r2e_pytest.log
scripts/dd_scale.sh r2e 2000000 2 1 2 4 > gpurun_out/r2e_scale_2M.log 2>&1
scripts/dd_scale.sh r2e 16000000 2 1 2 > gpurun_out/r2e_scale_16M.log 2>&1
cat gpurun_out/r2e_scale_2M.log gpurun_out/r2e_scale_16M.log | grep -v "^   counters" | cut -c1-600
