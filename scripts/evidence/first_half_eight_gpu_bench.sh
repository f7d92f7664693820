#!/bin/bash
# 8-GPU lease: the driver's scaling command (domain form) and the replicated form beside it
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r2p_bench_${N}gpu_dd.json 2> gpurun_out/r2p_bench_${N}gpu_dd.err; echo "dd rc=$?"
timeout 200 $TR bench.py --gpus $N --steps 8 --warmup 4 --no-config5 --no-e2e --no-check > gpurun_out/r2p_bench_${N}gpu_dd_long.json 2> gpurun_out/r2p_bench_${N}gpu_dd_long.err; echo "dd long rc=$?"
timeout 200 $TR bench.py --gpus $N --steps 8 --warmup 4 --no-config5 --no-e2e --no-check --decomposition 0 > gpurun_out/r2p_bench_${N}gpu_rep.json 2> gpurun_out/r2p_bench_${N}gpu_rep.err; echo "rep rc=$?"
for f in gpurun_out/r2p_*.err; do tail -n 2 $f | cut -c1-200; done
