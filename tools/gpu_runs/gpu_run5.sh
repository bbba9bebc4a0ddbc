#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_gpu_zz_golden.py -m gpu -x -q -k "not 2_24 and not 2_20" > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest5.log
for P in 64 128; do FB_BATCH_P=$P timeout 300 python tools/cfg_small.py > gpurun_out/r02_cfg_small_b_P$P.json 2> gpurun_out/r02_cfg_small_b_P$P.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_cfg_small_b_P$P.json').read().strip().splitlines()[-1])
    c=d['cfg2_batch256']; print('P=$P', 'batch_s', round(c['batch_s']*1e3,2), 'ms_per_proof', round(c['ms_per_proof'],4), 'sha', c['all_256_proofs_sha256_equal_cpu_oracle'], 'cfg1 prove ms', round(d['cfg1']['prove_s']*1e3,3), d['cfg1']['proof_bytes_equal_cpu_oracle'])
except Exception as e:
    print('P=$P failed', e); print(open('gpurun_out/r02_cfg_small_b_P$P.err').read()[-1500:])
PY
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --log-rows 20 > gpurun_out/r02_bench5_2e20.json 2> gpurun_out/r02_bench5_2e20.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench5_2e20.json').read().strip().splitlines()[-1])
print('2^20 value', round(d['value']*1e3,3), 'e2e', round(d['e2e']['value']*1e3,3), 'pageable', round(d['e2e']['pageable']['value']*1e3,3), 'sha_ok', d.get('proof_sha256_ok'), {k:round(v['ms_per_prove'],3) for k,v in d['kernel_ms'].items()})
PY
