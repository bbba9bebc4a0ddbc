#!/bin/bash
# 1-GPU call: window-size sweep at the per-rank MSM size of an 8-GPU 2^24 prove (2^21 points), sanitizer passes
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --log-rows 21"
for c in 17 18 19 20; do FB_MSM_TABLE_C=$c timeout 300 $B > gpurun_out/r02_sweep21_c$c.json 2> gpurun_out/r02_sweep21_c$c.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_sweep21_c$c.json').read().strip().splitlines()[-1])
    print('c=$c', 'value', round(d['value']*1e3,2), 'serial', round(d['serial_schedule_s']*1e3,2), d['config']['msm']['digits_per_scalar'], {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()})
except Exception as e:
    print('c=$c failed', e)
PY
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_sanitizer_racecheck.log
