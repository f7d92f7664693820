#!/bin/bash
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=200
(time python -m pytest tests/test_domains.py tests/test_gpu_ics.py tests/test_multi_gpu.py -q -x --durations=5) > gpurun_out/r2i_pytest.log 2>&1
tail -5 gpurun_out/r2i_pytest.log
scripts/dd_scale.sh r2i 2000000 2 4 8 > gpurun_out/r2i_scale_2M.log 2>&1
grep -v "^   counters" gpurun_out/r2i_scale_2M.log | cut -c1-420
