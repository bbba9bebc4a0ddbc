#!/bin/bash
# warp-per-bucket gather for small problems: tests, then single-prove latency (cfg 1) and batch throughput (cfg 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prove.py tests/test_gpu_zz_golden.py -m gpu -x -q -k "not 2_24 and not 2_20" > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest10.log
for c in 0 8 10; do for P in 64 128; do
if [ $c = 0 ]; then unset FB_MSM_TABLE_C; else export FB_MSM_TABLE_C=$c; fi
FB_BATCH_P=$P timeout 300 python tools/cfg_small.py > gpurun_out/r02_small10_c${c}_P$P.json 2> gpurun_out/r02_small10.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_small10_c${c}_P$P.json').read().strip().splitlines()[-1])
    c2=d['cfg2_batch256']; print('c=$c P=$P', 'cfg1 prove ms', round(d['cfg1']['prove_s']*1e3,3), d['cfg1']['proof_bytes_equal_cpu_oracle'], '| cfg2 batch ms', round(c2['batch_s']*1e3,2), c2['all_256_proofs_sha256_equal_cpu_oracle'])
except Exception as e:
    print('c=$c P=$P failed', e); print(open('gpurun_out/r02_small10.err').read()[-800:])
PY
done; done
unset FB_MSM_TABLE_C
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --log-rows 12 > gpurun_out/r02_bench10_2e12.json 2>gpurun_out/r02_bench10.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench10_2e12.json').read().strip().splitlines()[-1]);print('2^12 value ms', d['value']*1e3, 'sha', d['proof_sha256_ok'])"
