#!/bin/bash
# 1-GPU box: ncu launch list of the bench command on the final tree (2^24 rows; the last prove is the serial-schedule one)
mkdir -p gpurun_out
timeout 340 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_2e24_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_2e24_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02_bench_2e24_launches.csv
python tools/launch_summary.py gpurun_out/r02_bench_2e24_launches.csv --last-prove | head -12
