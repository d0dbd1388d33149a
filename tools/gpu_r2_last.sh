#!/bin/bash
# what the driver runs at round end, on HEAD: the GPU suite and smoke()
TAG=${1:-r2w}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_${TAG}.log 2>&1; echo exit=$?; tail -3 gpurun_out/pytest_${TAG}.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
