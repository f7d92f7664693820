#!/bin/bash
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=200
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py tests/test_domains.py -q -x --durations=5 -k "far or step_host or tree_reuse or 10k_steps") > gpurun_out/r3e_pytest.log 2>&1
tail -n 14 gpurun_out/r3e_pytest.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r3e_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3e_bench.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"])
print(d["stage_ms_per_step"])
print(d["roofline"]["launch_ms"], d["roofline"]["launches"], d["roofline"]["frac"], d["fp64"]["frac"], d["gravity_far_reuse"]["near_only_evaluations"])
print(d.get("config5", {}).get("ms_per_step"), d.get("config5", {}).get("stage_ms_per_step"))
PY
