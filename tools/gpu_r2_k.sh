#!/bin/bash
# A/B on ONE box: PDL for the small kernels, shuffled bias; C4 / C2 lines
TAG=${1:-r2k}
mkdir -p gpurun_out
echo "== selftest_gemm quick"; timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm quick 1 1 2>&1 | grep -E "FAIL|selftest_gemm:" | head -5
for sh in 0 1; do echo "== selftest l0 TFK_GEMM_BIAS_SHFL=$sh"; TFK_GEMM_BIAS_SHFL=$sh timeout 120 ./tfkaldi_b200/csrc/build/selftest_gemm l0 | grep BENCH | cut -c1-160; done
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_${TAG}.log 2>&1; echo exit=$?; tail -3 gpurun_out/pytest_${TAG}.log
one() { # cfg name env...
  cfg=$1; name=$2; shift; shift
  env "$@" timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_${cfg}_${name}_${TAG}.json 2> gpurun_out/bench_${cfg}_${name}_${TAG}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${cfg}_${name}_${TAG}.json"))
print("$cfg $name", "%.3e" % d["value"], d["timing"]["windows_ms_per_step"], d["roofline"]["per_step_us_by_kernel_class"])
PY
}
for rep in 1 2; do
one c4 pdl1_shfl1 TFK_X=1
one c4 pdl0_shfl1 TFK_PDL_SMALL=0
one c2 pdl1_shfl1 TFK_X=1
one c2 pdl1_shfl0 TFK_GEMM_BIAS_SHFL=0
one c2 pdl0_shfl1 TFK_PDL_SMALL=0
done
