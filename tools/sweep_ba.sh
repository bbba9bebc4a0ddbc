#!/bin/bash
# batch-affine rounds on/off at 2^$1 rows
lg=$1
for ba in 0 1; do
  FB_MSM_BA=$ba python bench.py --log-rows $lg --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ba=$ba', 'lg=$lg', round(j['value'], 5), round(j['serial_schedule_s'], 5), {k: round(v['ms_per_prove'], 2) for k, v in j['kernel_ms'].items()}, j['proof_verifies'], j['pk_hbm_bytes'])
"
done
