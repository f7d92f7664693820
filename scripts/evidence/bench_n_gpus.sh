#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r3m_bench_${N}gpu.json 2> gpurun_out/r3m_bench_${N}gpu.err; echo "bench rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/r3m_bench_${N}gpu.err | tail -n 3 | cut -c1-300
python - $N <<'PY'
import json, sys
N = sys.argv[1]
line = [l for l in open(f"gpurun_out/r3m_bench_{N}gpu.json").read().splitlines() if l.startswith("{")][0]
d = json.loads(line)
print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["resident_hits"], d["e2e"]["cold"]["value"])
print(d["stage_ms_per_step"]); print(d.get("multi_gpu_check")); print(d.get("fallback")); print(d.get("per_rank_ms_per_step"))
print("config5", d["config5"].get("ms_per_step"), d["config5"].get("error"))
PY
