#!/bin/bash
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=200
(time timeout 900 python -m pytest tests/test_domains.py tests/test_multi_gpu.py tests/test_gpu_parity.py -q -x --durations=5 -k "recognise or resident or far_reuse or step_host or three_steps") > gpurun_out/r3h_pytest.log 2>&1
tail -n 14 gpurun_out/r3h_pytest.log
