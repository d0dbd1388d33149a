#!/bin/bash
# quicker GPU visit: GEMM self-test, parity tests, smoke, bench line
TAG=${1:-q}
mkdir -p gpurun_out
echo "== selftest_gemm"; timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm ${2:-quick} > gpurun_out/selftest_${TAG}.log 2>&1; echo exit=$?; grep -E "FAIL|BENCH|fused bwd|selftest_gemm:" gpurun_out/selftest_${TAG}.log | head -30
echo "== selftest_gemm, every tile as two half-width tiles"; TFK_GEMM_HALF_TILES=all timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm quick > gpurun_out/selftest_half_${TAG}.log 2>&1; echo exit=$?; grep -E "FAIL|selftest_gemm:" gpurun_out/selftest_half_${TAG}.log | head -10
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
