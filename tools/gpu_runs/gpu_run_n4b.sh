#!/bin/bash
# 4-GPU call: 2^24 at N=4 with the retuned window model (c = 20 at 2^22-point shards)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 4 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_2e24_n4.json 2> gpurun_out/r02_bench_2e24_n4.err; echo "bench n4 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_2e24_n4.json').read().strip().splitlines()[-1])
print('n4', 'value', round(d['value']*1e3,2), 'e2e', round(d['e2e']['value']*1e3,2), 'sha_ok', d.get('proof_sha256_ok'), d['config']['msm'])
print('   kernel_ms', {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()}, 'serial', round(d['serial_schedule_s']*1e3,1))
PY
