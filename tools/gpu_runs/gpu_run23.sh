#!/bin/bash
# 2-GPU box: full GPU suite incl. the NCCL test (device witness read in place, collective path), N=2 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest23.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest23.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2e24_n2.json 2> gpurun_out/r02_bench_2e24_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_2e24_n2.json').read().strip().splitlines()[-1])
print('n2', 'value', round(d['value']*1e3,2), 'e2e', round(d['e2e']['value']*1e3,2), 'sha_ok', d.get('proof_sha256_ok'))
PY
