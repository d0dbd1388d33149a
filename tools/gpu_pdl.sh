#!/bin/bash
# experiment: programmatic dependent launch on/off - GEMM bench, then parity tests + bench line with it on
mkdir -p gpurun_out
for pdl in 0 1; do
  echo "== selftest_gemm bench, TFK_PDL=$pdl"
  TFK_PDL=$pdl timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm bench 1 1 > gpurun_out/pdl_$pdl.log 2>&1; echo exit=$?
  grep -E "BENCH|fused bwd" gpurun_out/pdl_$pdl.log | cut -c1-150
done
echo "== bench TFK_PDL=0"; TFK_PDL=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err; cut -c1-260 gpurun_out/bench_pdl0.json
bash tools/gpu_quick.sh ${1:-pdl} full
