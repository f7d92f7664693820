#!/bin/bash
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=200
(time timeout 900 python -m pytest tests/test_domains.py tests/test_multi_gpu.py -q -x --durations=3) > gpurun_out/r3d_pytest_mg.log 2>&1
tail -n 8 gpurun_out/r3d_pytest_mg.log
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -q -x --durations=3) > gpurun_out/r3d_pytest_parity.log 2>&1
tail -n 8 gpurun_out/r3d_pytest_parity.log
