#!/bin/bash
TAG=${1:-r2f}
mkdir -p gpurun_out
echo "== selftest_gemm quick"; timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm quick > gpurun_out/selftest_${TAG}.log 2>&1; echo exit=$?; grep -E "FAIL|selftest_gemm:" gpurun_out/selftest_${TAG}.log | head -10
echo "== selftest_gemm quick, generic epilogue"; TFK_GEMM_GENERIC_EPILOGUE=1 timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm quick 1 1 2>&1 | grep -E "FAIL|selftest_gemm:" | head -5
echo "== selftest_gemm full 2cta"; timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm full 1 1 2>&1 | grep -E "FAIL|BENCH|fused bwd|selftest_gemm:" | cut -c1-160
echo "== selftest quick + l0, A-stationary on"; TFK_GEMM_A_RESIDENT=1 timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm quick 1 1 2>&1 | grep -E "FAIL|selftest_gemm:" | head -5; TFK_GEMM_A_RESIDENT=1 timeout 120 ./tfkaldi_b200/csrc/build/selftest_gemm l0 | grep BENCH | cut -c1-160
echo "== selftest_gemm l0"; timeout 120 ./tfkaldi_b200/csrc/build/selftest_gemm l0 | grep BENCH | cut -c1-160
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_${TAG}.log 2>&1; echo exit=$?; tail -5 gpurun_out/pytest_${TAG}.log
echo "== bench c4"; timeout 900 python bench.py --config c4 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_${TAG}.json 2> gpurun_out/bench_c4_${TAG}.err; echo exit=$?; tail -3 gpurun_out/bench_c4_${TAG}.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_c4_${TAG}.json"))
print("c4", d["value"], d["ms_per_step"], d["e2e"]["value"], d["parity_mode"]["value"] if d.get("parity_mode") else None, d["roofline"]["per_step_us_by_kernel_class"], d["hbm_kernels"].get("bn"))
PY
echo "== bench c2"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c2_${TAG}.json 2> gpurun_out/bench_c2_${TAG}.err; echo exit=$?; tail -3 gpurun_out/bench_c2_${TAG}.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_c2_${TAG}.json"))
print("c2", d["value"], d["timing"]["windows_ms_per_step"], d["e2e"]["value"], d["parity_mode"]["value"], d["parity_mode"]["windows_ms_per_step"], d["roofline"]["per_step_us_by_kernel_class"], d["roofline"]["frac_of_burst_peak"], d["clocks"])
PY
echo "== ncu full: layer-0 forward with bits"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tfk_gemm2 -s 6 -c 1 -o gpurun_out/prof_l0fwd_${TAG} -f ./tfkaldi_b200/csrc/build/selftest_gemm l0 4 > gpurun_out/ncu_l0_${TAG}.log 2>&1; echo exit=$?
