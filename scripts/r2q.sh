#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r2q_bench_${N}gpu_dd.json 2> gpurun_out/r2q_bench_${N}gpu_dd.err; echo "dd rc=$?"
tail -n 2 gpurun_out/r2q_bench_${N}gpu_dd.err | cut -c1-300
