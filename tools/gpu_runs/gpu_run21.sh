#!/bin/bash
# the bench line of the final tree, as the driver runs it
mkdir -p gpurun_out
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench21.json 2> gpurun_out/r02_bench21.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench21.json').read().strip().splitlines()[-1])
print('ours value', d['value'], 'e2e', d['e2e']['value'], 'pageable', d['e2e']['pageable']['value'], 'sha_ok', d['proof_sha256_ok'], 'roofline frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['proof_bytes_equal_gpu'])
print({k:(v.get('value') or v.get('prove_s') or v.get('ms_per_proof') or v.get('ms')) for k,v in d['also'].items()}, d['also']['cfg2_batch256'].get('stream'))
PY
