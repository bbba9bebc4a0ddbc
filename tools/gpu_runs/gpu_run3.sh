#!/bin/bash
# GPU call 3 (1 GPU): radix sort with shared-memory atomic ranking vs the counting sort
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py -m gpu -x -q -k "not 2_24 and not 2_20" > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest3.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
timeout 600 $B > gpurun_out/r02_bench3_radix.json 2> gpurun_out/r02_bench3_radix.err; echo "bench radix rc=$?"
for f in radix; do python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench3_$f.json').read().strip().splitlines()[-1])
print('$f', 'value', round(d['value']*1e3,2), 'e2e', round(d['e2e']['value']*1e3,2), 'sha_ok', d.get('proof_sha256_ok'), {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()})
PY
done
SERIAL=1 LOG=20 REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_radix|k_bucket_offsets|k_scan' --csv --log-file gpurun_out/r02_sort20_launches.csv python tools/prove_once.py > gpurun_out/r02_sort20_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02_sort20_launches.csv
