#!/bin/bash
# 2-GPU lease: both multi-rank forms of bench.py at 16M (and the domain form at 2M)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 3 --warmup 3 --particles 2000000 --no-config5 --no-e2e > gpurun_out/r2f_bench_2M_dd.json 2> gpurun_out/r2f_bench_2M_dd.err; echo "2M dd rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 4 --warmup 3 --no-config5 > gpurun_out/r2f_bench_16M_dd.json 2> gpurun_out/r2f_bench_16M_dd.err; echo "16M dd rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 4 --warmup 3 --no-config5 --decomposition 0 --no-check > gpurun_out/r2f_bench_16M_rep.json 2> gpurun_out/r2f_bench_16M_rep.err; echo "16M rep rc=$?"
tail -2 gpurun_out/r2f_*.err | cut -c1-300
