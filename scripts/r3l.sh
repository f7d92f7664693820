#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3l_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r3l_smoke.log
(time timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --durations=5) > gpurun_out/r3l_pytest.log 2>&1
tail -n 10 gpurun_out/r3l_pytest.log
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/r3l_bench.json 2> gpurun_out/r3l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3l_bench.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["cold"]["value"], d["e2e"]["cold"]["ms_per_step"])
print(d["stage_ms_per_step"])
PY
