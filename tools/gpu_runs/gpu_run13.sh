#!/bin/bash
# ncu full captures of the NTT passes at 2^24 (one strided DIF pass, the contiguous mid pass, one strided DIT pass)
mkdir -p gpurun_out
SERIAL=1 LOG=24 REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 3 -c 5 -o gpurun_out/r02_ntt24_full python tools/prove_once.py > gpurun_out/r02_ntt24_full.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/r02_ntt24_full.ncu-rep; tail -3 gpurun_out/r02_ntt24_full.log
