#!/bin/bash
# 4-GPU call: the NCCL test on the final tree, 2^24 at N=4 and N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_pytest_n4box.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_n4box.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 4 --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_2e24_n4.json 2> gpurun_out/r02_bench_2e24_n4.err; echo "bench n4 rc=$?"
timeout 900 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2e24_n2.json 2> gpurun_out/r02_bench_2e24_n2.err; echo "bench n2 rc=$?"
for f in 2e24_n4 2e24_n2; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_$f.json').read().strip().splitlines()[-1])
    print('$f', 'value', round(d['value']*1e3,2), 'e2e', round(d['e2e']['value']*1e3,2), 'sha_ok', d.get('proof_sha256_ok'), 'verifies', d.get('proof_verifies'), 'setup_s', round(d['setup_s'],1), d['config']['msm'])
    print('   kernel_ms', {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()}, 'serial', round(d['serial_schedule_s']*1e3,1))
except Exception as e:
    print('$f failed', e); print(open('gpurun_out/r02_bench_$f.err').read()[-2500:])
PY
done
