#!/bin/bash
# 8-GPU lease: the driver's scaling command for N = 8 (domain form, with the 64M sub-record)
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 500 $TR bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r2h_bench_${N}gpu.json 2> gpurun_out/r2h_bench_${N}gpu.err ) 2>&1 | grep real; echo "rc=$?"
tail -n 3 gpurun_out/r2h_bench_${N}gpu.err | cut -c1-400
tail -n 1 gpurun_out/r2h_bench_${N}gpu.json | cut -c1-300
