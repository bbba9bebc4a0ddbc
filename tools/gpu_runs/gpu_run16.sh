#!/bin/bash
# window sweep at the per-rank size of a 4-GPU 2^24 prove (2^22 points) on one GPU
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --log-rows 22"
for c in 17 19 20; do FB_MSM_TABLE_C=$c timeout 300 $B > gpurun_out/r02_sweep22_c$c.json 2> gpurun_out/r02_sweep22_c$c.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_sweep22_c$c.json').read().strip().splitlines()[-1])
    print('c=$c W=%d' % d['config']['msm']['digits_per_scalar'], 'value', round(d['value']*1e3,2), 'serial', round(d['serial_schedule_s']*1e3,2), {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()})
except Exception as e:
    print('c=$c failed', e)
PY
done
