#!/bin/bash
# window / task-length sweep for the small circuits (single-prove latency of cfg 1 against batch throughput of cfg 2)
mkdir -p gpurun_out
for c in 8 9 10; do for tl in 4 5 6; do
FB_MSM_TABLE_C=$c FB_MSM_TASK_LOG=$tl timeout 300 python tools/cfg_small.py > gpurun_out/r02_small_sweep_c${c}_t$tl.json 2> gpurun_out/r02_small_sweep.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_small_sweep_c${c}_t$tl.json').read().strip().splitlines()[-1])
    c2=d['cfg2_batch256']; print('c=$c tl=$tl', 'cfg1 prove ms', round(d['cfg1']['prove_s']*1e3,3), d['cfg1']['proof_bytes_equal_cpu_oracle'], '| cfg2 batch ms', round(c2['batch_s']*1e3,2), c2['all_256_proofs_sha256_equal_cpu_oracle'])
except Exception as e:
    print('c=$c tl=$tl failed', e)
PY
done; done
