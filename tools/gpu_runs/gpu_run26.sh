#!/bin/bash
# 1-GPU box: the full GPU suite on the final tree (the virtual-rank H-pipeline test now runs both exchange schedules)
mkdir -p gpurun_out
timeout 232 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest26.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest26.log
