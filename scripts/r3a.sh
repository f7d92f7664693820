#!/bin/bash
# round 3 (second half of round 2): far-field reuse of the gravity walk - bench at 16M, then the parity tests
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-config5 --no-e2e > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3a_bench.json"))
print(d["ms_per_step"], d["stage_ms_per_step"], d.get("state_hash"), d.get("state_sums"))
PY
(time timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --durations=5 -k "far_field or 10k or accretion or tree_reuse or sink_creation or pool") > gpurun_out/r3a_pytest.log 2>&1
tail -n 15 gpurun_out/r3a_pytest.log
