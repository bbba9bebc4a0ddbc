#!/bin/bash
# round 2, GPU call 2: radix sort + new validation tests; A/B of the sort and of batch-affine G2; launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest2.log
tail -15 gpurun_out/r02_pytest2.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 $B > gpurun_out/r02_bench2_radix.json 2> gpurun_out/r02_bench2_radix.err; echo "bench radix rc=$?"
FB_MSM_SORT=atomic timeout 600 $B > gpurun_out/r02_bench2_atomic.json 2> gpurun_out/r02_bench2_atomic.err; echo "bench atomic rc=$?"
FB_MSM_BA=3 timeout 600 $B > gpurun_out/r02_bench2_ba3.json 2> gpurun_out/r02_bench2_ba3.err; echo "bench ba3 rc=$?"
FB_MSM_BA=3 FB_MSM_BA_ROUNDS=1 timeout 600 $B > gpurun_out/r02_bench2_ba3r1.json 2> gpurun_out/r02_bench2_ba3r1.err; echo "bench ba3r1 rc=$?"
for f in radix atomic ba3 ba3r1; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench2_$f.json').read().strip().splitlines()[-1])
    print('$f', 'value', round(d['value']*1e3,2), 'e2e', round(d['e2e']['value']*1e3,2), 'sha_ok', d.get('proof_sha256_ok'), {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()})
except Exception as e:
    print('$f', 'failed', e)
PY
done
SERIAL=1 LOG=24 REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_prove24_launches.csv python tools/prove_once.py > gpurun_out/r02_prove24_ncu.log 2>&1; echo "ncu rc=$?"
