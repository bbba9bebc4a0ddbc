#!/bin/bash
# lane-cooperative segment sums: tests, small-circuit numbers, the 8-GPU-shard size on one GPU, sanitizer
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not 2_24" > gpurun_out/r02_pytest18.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest18.log
timeout 300 python tools/cfg_small.py > gpurun_out/r02_small18.json 2> gpurun_out/r02_small18.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_small18.json').read().strip().splitlines()[-1])
c2=d['cfg2_batch256']; print('cfg1 prove ms', round(d['cfg1']['prove_s']*1e3,3), d['cfg1']['proof_bytes_equal_cpu_oracle'], '| cfg2 batch ms', round(c2['batch_s']*1e3,2), c2['all_256_proofs_sha256_equal_cpu_oracle'], '| stream ms', round(c2['stream']['batch_s']*1e3,2))
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --log-rows 21 > gpurun_out/r02_bench18_2e21.json 2> gpurun_out/r02_bench18.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench18_2e21.json').read().strip().splitlines()[-1])
print('2^21 value', round(d['value']*1e3,2), 'serial', round(d['serial_schedule_s']*1e3,2), {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()})
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02_sanitizer_racecheck.log
