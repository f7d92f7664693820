#!/bin/bash
# 1-GPU evidence for profiles/: drift table on the GPU, the default bench line (with the 64M sub-record and the CPU baseline),
# the ncu launch list of the bench command and one --set full capture of the three walk kernels at 16M.
mkdir -p gpurun_out
python scripts/drift_report.py gpurun_out/r2_drift_disc10k.md > gpurun_out/r2m_drift.log 2>&1; echo "drift rc=$?"
python bench.py --steps 4 --warmup 3 > gpurun_out/r2m_bench_1gpu.json 2> gpurun_out/r2m_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config5"
$CMD > gpurun_out/r2m_plain.json 2> gpurun_out/r2m_plain.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launch_list_bench.csv $CMD > gpurun_out/r2m_ncu_list.log 2>&1; echo "ncu list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config5 --no-e2e"
$CMD2 > gpurun_out/r2m_plain2.json 2> gpurun_out/r2m_plain2.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gravity|k_density|k_force' -s 27 -c 4 -o gpurun_out/prof_r2_16M -f $CMD2 > gpurun_out/r2m_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/prof_r2_16M.ncu-rep gpurun_out/r2_launch_list_bench.csv
tail -n 3 gpurun_out/r2m_ncu_full.log
