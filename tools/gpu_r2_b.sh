#!/bin/bash
# round 2, visit B: bf16x3 4-tile staging + batch-norm partial sums in the dgrad epilogue; C4 kernel list; layer-0 ncu capture
TAG=${1:-r2b}
mkdir -p gpurun_out
echo "== selftest_gemm quick"; timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm quick > gpurun_out/selftest_${TAG}.log 2>&1; echo exit=$?; grep -E "FAIL|selftest_gemm:" gpurun_out/selftest_${TAG}.log | head -10
echo "== selftest_gemm bench (2cta)"; timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm bench 1 1 > gpurun_out/selftest_bench_${TAG}.log 2>&1; grep -E "BENCH" gpurun_out/selftest_bench_${TAG}.log | cut -c1-160
echo "== selftest_gemm l0"; timeout 120 ./tfkaldi_b200/csrc/build/selftest_gemm l0 | grep BENCH | cut -c1-160
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -rP > gpurun_out/pytest_${TAG}.log 2>&1; echo exit=$?; tail -5 gpurun_out/pytest_${TAG}.log
echo "== bench c4"; timeout 900 python bench.py --config c4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_${TAG}.json 2> gpurun_out/bench_c4_${TAG}.err; echo exit=$?; tail -3 gpurun_out/bench_c4_${TAG}.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_c4_${TAG}.json"))
print("c4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_mode"]["value"] if d.get("parity_mode") else None, d["roofline"]["per_step_us_by_kernel_class"], d["hbm_kernels"].get("bn"))
PY
echo "== bench c2"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_${TAG}.json 2> gpurun_out/bench_c2_${TAG}.err; echo exit=$?; python - <<PY
import json
d=json.load(open("gpurun_out/bench_c2_${TAG}.json"))
print("c2", d["value"], d["timing"]["windows_ms_per_step"], d["e2e"]["value"], d["parity_mode"]["value"], d["parity_mode"]["windows_ms_per_step"], d["roofline"]["per_step_us_by_kernel_class"])
PY
echo "== ncu launch list c4"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_c4_${TAG}.csv python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_c4_${TAG}.log 2>&1; echo exit=$?
echo "== ncu full: layer-0 forward / wgrad / hidden forward, stand-alone"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tfk_gemm2 -s 6 -c 1 -o gpurun_out/prof_l0fwd_${TAG} -f ./tfkaldi_b200/csrc/build/selftest_gemm l0 4 > gpurun_out/ncu_l0_${TAG}.log 2>&1; echo exit=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tfk_gemm2 -s 13 -c 1 -o gpurun_out/prof_l0wgrad_${TAG} -f ./tfkaldi_b200/csrc/build/selftest_gemm l0 4 >> gpurun_out/ncu_l0_${TAG}.log 2>&1; echo exit=$?
echo "== ncu full: the non-GEMM kernels of a C4 step"
timeout 900 ncu --set full --clock-control none -k regex:"bn_|colsum|softmax_ce|adam_kernel|split_f32|accum_loss" -s 120 -c 26 -o gpurun_out/prof_c4_small_${TAG} -f python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_c4_full_${TAG}.log 2>&1; echo exit=$?
ls -la gpurun_out | tail -12
