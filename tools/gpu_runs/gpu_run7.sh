#!/bin/bash
# 1-GPU call: what the driver runs at round end (full GPU suite, smoke, bench with its arguments, the reference arm),
# plus the ncu passes bench.py's roofline cites
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest7.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest7.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke7.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke7.log
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err ) 2>&1 | grep real; echo "bench rc=$?"
( time timeout 1700 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench7_ref.json 2> gpurun_out/r02_bench7_ref.err ) 2>&1 | grep real; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench7.json').read().strip().splitlines()[-1])
print('ours value', d['value'], 'e2e', d['e2e']['value'], 'sha_ok', d['proof_sha256_ok'], 'roofline frac', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['proof_bytes_equal_gpu'])
print({k:(v.get('value') or v.get('prove_s') or v.get('ms_per_proof') or v.get('ms')) for k,v in d['also'].items()})
r=json.loads(open('gpurun_out/r02_bench7_ref.json').read().strip().splitlines()[-1])
print('ref value', r['value'], 'steps', r['steps'], 'warmup', r['warmup'], 'sha_ok', r['proof_sha256_ok'], 'verifies', r['cpu_baseline']['proof_verifies'], 'params_ok', r['cpu_baseline']['params_sha256_ok'], 'setup_s', r['cpu_baseline']['setup_s'], r['cpu_baseline']['all_prove_s'])
PY
SERIAL=1 LOG=24 REPS=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_accumulate --csv --log-file gpurun_out/r02_acc24_traffic.csv python tools/prove_once.py > gpurun_out/r02_acc24_traffic.log 2>&1; echo "ncu traffic rc=$?"
SERIAL=1 LOG=20 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_accumulate.*FqCfg' -c 1 -o gpurun_out/r02_acc20_full python tools/prove_once.py > gpurun_out/r02_acc20_full.log 2>&1; echo "ncu full rc=$?"
