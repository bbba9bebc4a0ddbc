#!/bin/bash
# 2-GPU box: NCCL tests + N=2 bench with the overlapped NTT exchanges, and with FB_DIST_NO_OVERLAP=1 beside it
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_pytest24.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest24.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench24_ov.json 2> gpurun_out/r02_bench24_ov.err; echo "bench ov rc=$?"
FB_DIST_NO_OVERLAP=1 timeout 300 $TR --nproc-per-node 2 --master-port 29572 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench24_noov.json 2> gpurun_out/r02_bench24_noov.err; echo "bench noov rc=$?"
for f in ov noov; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench24_$f.json').read().strip().splitlines()[-1])
    print('$f', 'value', round(d['value']*1e3,2), 'e2e', round(d['e2e']['value']*1e3,2), 'sha_ok', d.get('proof_sha256_ok'), d.get('stage_ms'))
except Exception as e:
    print('$f failed', e); print(open('gpurun_out/r02_bench24_$f.err').read()[-2000:])
PY
done
