#!/bin/bash
# full GPU suite on the current tree; configs[0]/[1] with the defaults; launch list of a batched chunk
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest11.log
timeout 300 python tools/cfg_small.py --cpu > gpurun_out/r02_cfg_small_final.json 2> gpurun_out/r02_cfg_small_final.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_cfg_small_final.json').read().strip().splitlines()[-1])
c2=d['cfg2_batch256']; print('cfg1 prove ms', round(d['cfg1']['prove_s']*1e3,3), d['cfg1']['proof_bytes_equal_cpu_oracle'], d['cfg1'].get('cpu_baseline',{}).get('value'), '| cfg2 batch ms', round(c2['batch_s']*1e3,2), c2['all_256_proofs_sha256_equal_cpu_oracle'], c2.get('cpu_baseline',{}).get('value'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_cfg2_batch_launches.csv python tools/cfg_small.py > gpurun_out/r02_cfg2_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/r02_cfg2_batch_launches.csv > gpurun_out/r02_cfg2_batch_launches.txt; head -24 gpurun_out/r02_cfg2_batch_launches.txt
