#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prove.py -m gpu -x -q -k "stream or batch or cfg" > gpurun_out/r02_pytest15.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest15.log
timeout 300 python tools/cfg_small.py > gpurun_out/r02_small15.json 2> gpurun_out/r02_small15.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_small15.json').read().strip().splitlines()[-1])
c2=d['cfg2_batch256']; print('cfg1 prove ms', round(d['cfg1']['prove_s']*1e3,3), '| cfg2 batch ms', round(c2['batch_s']*1e3,2), c2['all_256_proofs_sha256_equal_cpu_oracle'], '| stream', c2.get('stream'))
PY
LOG=13 N=64 timeout 300 python tools/stream_probe.py > gpurun_out/r02_stream_probe_2e13.json 2>&1; tail -1 gpurun_out/r02_stream_probe_2e13.json
