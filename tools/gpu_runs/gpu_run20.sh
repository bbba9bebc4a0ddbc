#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_gpu_zz_golden.py -m gpu -x -q -k "not 2_24 and not 2_20" > gpurun_out/r02_pytest20.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest20.log
for i in 1 2 3; do timeout 300 python tools/cfg_small.py > gpurun_out/r02_small20_$i.json 2> gpurun_out/r02_small20.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_small20_$i.json').read().strip().splitlines()[-1])
c2=d['cfg2_batch256']; print('cfg1 prove ms', round(d['cfg1']['prove_s']*1e3,3), d['cfg1']['proof_bytes_equal_cpu_oracle'], '| cfg2 batch ms', round(c2['batch_s']*1e3,2), c2['all_256_proofs_sha256_equal_cpu_oracle'], '| stream ms', round(c2['stream']['batch_s']*1e3,2), c2['stream']['proofs_equal_batch'])
PY
done
