#!/bin/bash
# 2-GPU call: NCCL test (collective fb_prove, sharded setup), 2^24 at N=2 with the golden digest, rollup-size smoke at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2e24_n2.json 2> gpurun_out/r02_bench_2e24_n2.err; echo "bench 2^24 n2 rc=$?"
timeout 1500 $TR --master-port 29512 bench.py --gpus 2 --rows 35695616 --steps 2 --warmup 3 > gpurun_out/r02_bench_cfg5_n2.json 2> gpurun_out/r02_bench_cfg5_n2.err; echo "bench cfg5 n2 rc=$?"
for f in 2e24_n2 cfg5_n2; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_$f.json').read().strip().splitlines()[-1])
    print('$f', 'value', round(d['value']*1e3,2), 'e2e', round(d['e2e']['value']*1e3,2), 'sha_ok', d.get('proof_sha256_ok'), 'verifies', d.get('proof_verifies'), 'setup_s', round(d['setup_s'],1), 'load_s', round(d['key_load_s'],1), 'circuit_s', round(d['circuit_s'],1), d['config']['msm'])
    print('   kernel_ms', {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()}, 'serial', round(d['serial_schedule_s']*1e3,1), d.get('stage_ms_serial'))
except Exception as e:
    print('$f failed', e); print(open('gpurun_out/r02_bench_$f.err').read()[-2500:])
PY
done
