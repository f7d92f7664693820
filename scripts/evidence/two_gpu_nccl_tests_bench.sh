#!/bin/bash
# 2-GPU lease: NCCL cases of the multi-rank tests, then the driver's scaling command at N = 2
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=300
(time timeout 900 python -m pytest tests/test_domains.py tests/test_multi_gpu.py -q -x --durations=5 -k "nccl or recognise or far_reuse") > gpurun_out/r3i_pytest.log 2>&1
tail -n 10 gpurun_out/r3i_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 4 --warmup 3 --no-config5 > gpurun_out/r3i_bench_2gpu.json 2> gpurun_out/r3i_bench_2gpu.err; echo "bench rc=$?"
tail -n 3 gpurun_out/r3i_bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3i_bench_2gpu.json"))
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["resident_hits"], d["e2e"]["cold"]["value"])
print(d["stage_ms_per_step"]); print(d.get("multi_gpu_check")); print(d.get("fallback"))
PY
