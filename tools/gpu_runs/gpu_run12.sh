#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_gpu_zz_golden.py -m gpu -x -q -k "batch or cfg or golden" > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest12.log
for P in 32 64 128; do FB_BATCH_P=$P timeout 300 python tools/cfg_small.py > gpurun_out/r02_small12_P$P.json 2> gpurun_out/r02_small12.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_small12_P$P.json').read().strip().splitlines()[-1])
    c2=d['cfg2_batch256']; print('P=$P', 'cfg1 prove ms', round(d['cfg1']['prove_s']*1e3,3), d['cfg1']['proof_bytes_equal_cpu_oracle'], '| cfg2 batch ms', round(c2['batch_s']*1e3,2), c2['all_256_proofs_sha256_equal_cpu_oracle'])
except Exception as e:
    print('P=$P failed', e); print(open('gpurun_out/r02_small12.err').read()[-800:])
PY
done
