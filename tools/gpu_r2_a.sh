#!/bin/bash
# round 2, visit A: GEMM self-test, the whole GPU suite (with the printed error figures), smoke, bench lines C2 / C4
TAG=${1:-r2a}
mkdir -p gpurun_out
echo "== selftest_gemm quick"; timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm quick > gpurun_out/selftest_${TAG}.log 2>&1; echo exit=$?; grep -E "FAIL|selftest_gemm:" gpurun_out/selftest_${TAG}.log | head -10
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rP > gpurun_out/pytest_${TAG}.log 2>&1; echo exit=$?; tail -5 gpurun_out/pytest_${TAG}.log; grep -E "flip-controlled|timed mode|bf16 log-lik|c5 bf16|gradient errors" gpurun_out/pytest_${TAG}.log | cut -c1-1500
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench c2"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_${TAG}.json 2> gpurun_out/bench_c2_${TAG}.err; echo exit=$?; tail -3 gpurun_out/bench_c2_${TAG}.err; cut -c1-3500 gpurun_out/bench_c2_${TAG}.json
echo "== bench c2 200 steps"; timeout 900 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_c2_200_${TAG}.json 2> gpurun_out/bench_c2_200_${TAG}.err; echo exit=$?; cut -c1-1200 gpurun_out/bench_c2_200_${TAG}.json
echo "== bench c4"; timeout 900 python bench.py --config c4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_${TAG}.json 2> gpurun_out/bench_c4_${TAG}.err; echo exit=$?; tail -3 gpurun_out/bench_c4_${TAG}.err; cut -c1-3000 gpurun_out/bench_c4_${TAG}.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-1200
