#!/bin/bash
# round-2 final evidence on one B200: full GPU suite, the default bench line, the reference arm, the ncu launch list of the
# bench command and one --set full capture of the walk kernels of a steady step at 16M
mkdir -p gpurun_out
export SPH_TEST_RANK_TIMEOUT=300
(time timeout 1500 python -m pytest tests -m gpu -q -x --durations=8) > gpurun_out/r3j_pytest_full.log 2>&1
tail -n 14 gpurun_out/r3j_pytest_full.log
timeout 900 python bench.py > gpurun_out/r3j_bench_1gpu.json 2> gpurun_out/r3j_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3j_bench_ref.json 2> gpurun_out/r3j_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config5 --no-e2e"
$CMD > gpurun_out/r3j_plain.json 2> gpurun_out/r3j_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launch_list_bench.csv $CMD > gpurun_out/r3j_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_gravity|k_density|k_force' -s 27 -c 9 -o gpurun_out/prof_r3_16M -f $CMD > gpurun_out/r3j_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_r3_16M.ncu-rep gpurun_out/r3_launch_list_bench.csv
tail -n 3 gpurun_out/r3j_ncu_full.log
