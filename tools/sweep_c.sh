#!/bin/bash
# window-size sweep for the MSM window tables: FB_MSM_TABLE_C in $2.. at 2^$1 rows
lg=$1; shift
for c in "$@"; do
  FB_MSM_TABLE_C=$c python bench.py --log-rows $lg --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('c=$c', 'lg=$lg', round(j['value'], 5), round(j['serial_schedule_s'], 5), {k: round(v['ms_per_prove'], 2) for k, v in j['kernel_ms'].items()})
"
done
