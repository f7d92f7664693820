#!/bin/bash
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=300
(time python -m pytest tests -q -m gpu --durations=10) > gpurun_out/r2l_pytest.log 2>&1
tail -25 gpurun_out/r2l_pytest.log
