#!/bin/bash
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=200
(time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py tests/test_domains.py -q -x --durations=5 -k "far or resident or recognise or step_host or 10k or accretion or sink") > gpurun_out/r3k_pytest.log 2>&1
tail -n 12 gpurun_out/r3k_pytest.log
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/r3k_bench.json 2> gpurun_out/r3k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3k_bench.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["cold"]["value"])
print(d["stage_ms_per_step"])
PY
