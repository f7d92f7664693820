#!/bin/bash
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --durations=5 -k "resident or step_host or far_field") > gpurun_out/r3g_pytest.log 2>&1
tail -n 14 gpurun_out/r3g_pytest.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r3g_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3g_bench.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"])
PY
