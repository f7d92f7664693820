#!/bin/bash
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=200
(time python -m pytest tests/test_domains.py tests/test_multi_gpu.py tests/test_gpu_ics.py -q -x --durations=3) > gpurun_out/r2n_pytest.log 2>&1
tail -4 gpurun_out/r2n_pytest.log
