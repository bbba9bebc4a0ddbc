#!/bin/bash
# 8-GPU call on the final tree: 2^24 rows at N=8 (golden digest on every rank)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_2e24_n8.json 2> gpurun_out/r02_bench_2e24_n8.err; echo "bench 2^24 n8 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_2e24_n8.json').read().strip().splitlines()[-1])
print('n8', 'value', round(d['value']*1e3,2), 'e2e', round(d['e2e']['value']*1e3,2), 'pageable', round(d['e2e']['pageable']['value']*1e3,2), 'sha_ok', d.get('proof_sha256_ok'), d['config']['msm'], 'setup', round(d['setup_s'],1))
print('   kernel_ms', {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()}, 'serial', round(d['serial_schedule_s']*1e3,1))
PY
