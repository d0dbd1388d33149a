#!/bin/bash
TAG=${1:-r2e}
mkdir -p gpurun_out
for cfg in "regs 0" "regs 1" "stream 0" "stream 1"; do set -- $cfg; echo "== TFK_SOFTMAX=$1 TFK_BN_FROM_Y=$2"; TFK_SOFTMAX=$1 TFK_BN_FROM_Y=$2 timeout 300 python tools/debug_c4.py linear 2>&1 | tail -22; done
echo "== pytest -m gpu (no -x)"; timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_${TAG}.log 2>&1; echo exit=$?; tail -8 gpurun_out/pytest_${TAG}.log
