#!/bin/bash
lg=$1; shift
for t in "$@"; do
  FB_MSM_TASK_LOG=$t python bench.py --log-rows $lg --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('task_log=$t', 'lg=$lg', round(j['value'], 5), round(j['serial_schedule_s'], 5), {k: round(v['ms_per_prove'], 2) for k, v in j['kernel_ms'].items()})
"
done
