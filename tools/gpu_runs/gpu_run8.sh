#!/bin/bash
# 1-GPU call: after removing the batch-affine experiment and relaxing the top-digit rule for small circuits
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not 2_24" > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest8.log
for P in 64 128 256; do FB_BATCH_P=$P timeout 300 python tools/cfg_small.py > gpurun_out/r02_cfg_small_c_P$P.json 2> gpurun_out/r02_cfg_small_c_P$P.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02_cfg_small_c_P$P.json').read().strip().splitlines()[-1])
    c=d['cfg2_batch256']; print('P=$P', 'batch_s', round(c['batch_s']*1e3,2), 'ms_per_proof', round(c['ms_per_proof'],4), 'sha', c['all_256_proofs_sha256_equal_cpu_oracle'], 'cfg1 prove ms', round(d['cfg1']['prove_s']*1e3,3), d['cfg1']['proof_bytes_equal_cpu_oracle'])
except Exception as e:
    print('P=$P failed', e); print(open('gpurun_out/r02_cfg_small_c_P$P.err').read()[-1500:])
PY
done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench8.json').read().strip().splitlines()[-1])
print('2^24 value', round(d['value']*1e3,2), 'e2e', round(d['e2e']['value']*1e3,2), 'sha_ok', d.get('proof_sha256_ok'), 'traffic', d['roofline']['traffic'], {k:round(v['ms_per_prove'],2) for k,v in d['kernel_ms'].items()})
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_target.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_racecheck.log
SERIAL=1 LOG=20 REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_accumulate -s 1 -c 1 -o gpurun_out/r02_acc20_full python tools/prove_once.py > gpurun_out/r02_acc20_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/r02_acc20_full.ncu-rep
