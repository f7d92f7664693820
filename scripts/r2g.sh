#!/bin/bash
# 2-GPU lease: domain tests over NCCL, then the domain form of bench.py at 2M and 16M
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=200
(time python -m pytest tests/test_domains.py -q -x --durations=5) > gpurun_out/r2o_pytest.log 2>&1
tail -3 gpurun_out/r2o_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 3 --warmup 3 --particles 2000000 --no-config5 --no-e2e --no-check > gpurun_out/r2o_bench_2M_dd.json 2> gpurun_out/r2o_bench_2M_dd.err; echo "2M dd rc=$?"
timeout 400 $TR bench.py --gpus 2 --steps 4 --warmup 3 --no-config5 --no-e2e --no-check > gpurun_out/r2o_bench_16M_dd.json 2> gpurun_out/r2o_bench_16M_dd.err; echo "16M dd rc=$?"
for f in gpurun_out/r2o_*.err; do echo $f; tail -n 2 $f | cut -c1-300; done
