#!/bin/bash
# parity tests + C4 (batch-norm + dropout) and C2 bench lines
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench c4"; timeout 600 python bench.py --config c4 --no-cpu-baseline > gpurun_out/bench_c4_${1:-t}.json 2> gpurun_out/bench_c4_${1:-t}.err; tail -2 gpurun_out/bench_c4_${1:-t}.err; cut -c1-2600 gpurun_out/bench_c4_${1:-t}.json
echo "== bench c2"; timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_${1:-t}.json 2> gpurun_out/bench_c2_${1:-t}.err; tail -2 gpurun_out/bench_c2_${1:-t}.err; cut -c1-300 gpurun_out/bench_c2_${1:-t}.json
