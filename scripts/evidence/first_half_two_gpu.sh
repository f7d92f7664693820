#!/bin/bash
# Round-2 helper (GPU box with >= 2 GPUs): the NCCL forms of the multi-rank tests, then bench lines for both multi-rank forms.
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=200
(time python -m pytest tests/test_multi_gpu.py tests/test_domains.py tests/test_gpu_ics.py -q --durations=5) > gpurun_out/r2d_pytest.log 2>&1
tail -5 gpurun_out/r2d_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 2 --particles 2000000 --no-config5 > gpurun_out/r2d_bench_2M_dd.json 2> gpurun_out/r2d_bench_2M_dd.err; echo "2M dd rc=$?"
timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 2 --particles 2000000 --no-config5 --decomposition 0 > gpurun_out/r2d_bench_2M_rep.json 2> gpurun_out/r2d_bench_2M_rep.err; echo "2M rep rc=$?"
timeout 900 $TR bench.py --gpus 2 --steps 4 --warmup 3 --no-config5 > gpurun_out/r2d_bench_16M_dd.json 2> gpurun_out/r2d_bench_16M_dd.err; echo "16M dd rc=$?"
timeout 900 $TR bench.py --gpus 2 --steps 4 --warmup 3 --no-config5 --decomposition 0 > gpurun_out/r2d_bench_16M_rep.json 2> gpurun_out/r2d_bench_16M_rep.err; echo "16M rep rc=$?"
timeout 900 python bench.py --gpus 1 --steps 4 --warmup 3 --no-config5 --no-cpu-baseline > gpurun_out/r2d_bench_16M_1.json 2> gpurun_out/r2d_bench_16M_1.err; echo "16M 1gpu rc=$?"
tail -3 gpurun_out/*.err
