#!/bin/bash
# task-length sweep: wave quantisation of the accumulate kernels (256-add tasks = 3 ms CTAs, 5-12 waves)
mkdir -p gpurun_out
for lg in 22 24; do for tl in 5 6 7 8; do
FB_MSM_TASK_LOG=$tl timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --log-rows $lg > gpurun_out/r02_tasksweep_${lg}_t$tl.json 2> gpurun_out/r02_tasksweep.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_tasksweep_${lg}_t$tl.json').read().strip().splitlines()[-1])
    print('2^$lg task_log=$tl', 'value', round(d['value']*1e3,2), 'serial', round(d['serial_schedule_s']*1e3,2), 'sha', d.get('proof_sha256_ok'), {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()})
except Exception as e:
    print('2^$lg tl=$tl failed', e)
PY
done; done
