#!/bin/bash
# final 1-GPU validation of the tree: full GPU suite, smoke, the bench line and the reference arm as the driver runs them
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest14.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest14.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke14.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke14.log
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench14.json 2> gpurun_out/r02_bench14.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench14.json').read().strip().splitlines()[-1])
print('ours value', d['value'], 'e2e', d['e2e']['value'], 'pageable', d['e2e']['pageable']['value'], 'sha_ok', d['proof_sha256_ok'], 'roofline frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['proof_bytes_equal_gpu'])
print({k:(v.get('value') or v.get('prove_s') or v.get('ms_per_proof') or v.get('ms')) for k,v in d['also'].items()})
print({k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()}, 'setup', d['setup_s'], 'load', d['key_load_s'])
PY
