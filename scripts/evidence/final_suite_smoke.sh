#!/bin/bash
# last check of the round on one B200: smoke() and the full GPU suite on the final build
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/final_smoke.log
(time timeout 1500 python -m pytest tests -m gpu -q -x --durations=5) > gpurun_out/final_pytest_full.log 2>&1
tail -n 12 gpurun_out/final_pytest_full.log
