#!/bin/bash
# round 2, GPU call 1: full GPU test suite, the default bench (with the full-size CPU baseline), pipe counters
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02_box.txt; nproc >> gpurun_out/r02_box.txt; free -g >> gpurun_out/r02_box.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench1.err
M=sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed.sum,sm__pipe_fmaheavy_cycles_active.sum,sm__pipe_fma_cycles_active.sum,sm__pipe_alu_cycles_active.sum,sm__cycles_active.sum,sm__cycles_elapsed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
LOG=20 timeout 600 ncu --metrics $M --clock-control none -k regex:'k_probe_imad$|k_accumulate' -c 12 --csv --log-file gpurun_out/r02_pipe_counters.csv python tools/pipe_counters.py > gpurun_out/r02_pipe_counters.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r02_pipe_counters.log
