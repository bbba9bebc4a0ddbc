#!/bin/bash
# 8-GPU call: 2^24 rows at N=8 (golden digest on every rank, serial-schedule stage/kernel breakdown) and
# BASELINE configs[4]: 35,695,616 rows (domain 2^26) on 8 x B200
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r02_box_n8.txt; nproc >> gpurun_out/r02_box_n8.txt; free -g | head -2 >> gpurun_out/r02_box_n8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_2e24_n8.json 2> gpurun_out/r02_bench_2e24_n8.err; echo "bench 2^24 n8 rc=$?"
timeout 1500 $TR --master-port 29522 bench.py --gpus 8 --rows 35695616 --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg5_n8.json 2> gpurun_out/r02_bench_cfg5_n8.err; echo "bench cfg5 n8 rc=$?"
for f in 2e24_n8 cfg5_n8; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_$f.json').read().strip().splitlines()[-1])
    print('$f', 'value', round(d['value']*1e3,2), 'e2e', round(d['e2e']['value']*1e3,2), 'pageable', round(d['e2e']['pageable']['value']*1e3,2), 'sha_ok', d.get('proof_sha256_ok'), 'verifies', d.get('proof_verifies'), 'setup_s', round(d['setup_s'],1), 'load_s', round(d['key_load_s'],1), 'circuit_s', round(d['circuit_s'],1), d['config']['msm'])
    print('   kernel_ms', {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()}, 'serial', round(d['serial_schedule_s']*1e3,1), d.get('stage_ms_serial'))
    print('   stage_ms', d.get('stage_ms'))
except Exception as e:
    print('$f failed', e); print(open('gpurun_out/r02_bench_$f.err').read()[-2500:])
PY
done
