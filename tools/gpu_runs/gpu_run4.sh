#!/bin/bash
# GPU call 4 (1 GPU): batched proves of small circuits (one set of launches per chunk of proofs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_gpu_msm.py tests/test_gpu_zz_golden.py -m gpu -x -q -k "not 2_24 and not 2_20" > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest4.log
for P in 32 64 128 256; do echo "FB_BATCH_P=$P"; FB_BATCH_P=$P timeout 300 python tools/cfg_small.py > gpurun_out/r02_cfg_small_P$P.json 2> gpurun_out/r02_cfg_small_P$P.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_cfg_small_P$P.json').read().strip().splitlines()[-1])
    c=d['cfg2_batch256']; print('P=$P', 'batch_s', round(c['batch_s']*1e3,2), 'ms_per_proof', round(c['ms_per_proof'],4), 'sha', c['all_256_proofs_sha256_equal_cpu_oracle'], 'cfg1 prove ms', round(d['cfg1']['prove_s']*1e3,3), d['cfg1']['proof_bytes_equal_cpu_oracle'])
except Exception as e:
    print('P=$P failed', e); print(open('gpurun_out/r02_cfg_small_P$P.err').read()[-1500:])
PY
done
echo slots; FB_BATCH_MODE=slots timeout 300 python tools/cfg_small.py > gpurun_out/r02_cfg_small_slots.json 2>&1; tail -c 600 gpurun_out/r02_cfg_small_slots.json
FB_BATCH_P=64 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_cfg2_batch_launches.csv python tools/cfg_small.py > gpurun_out/r02_cfg2_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02_cfg2_batch_launches.csv | head -30
