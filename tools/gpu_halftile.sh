#!/bin/bash
# experiment: GEMM bench (CTA-pair kernel only) under the three half-tile policies
mkdir -p gpurun_out
for pol in 0 all auto; do
  echo "== TFK_GEMM_HALF_TILES=$pol"
  TFK_GEMM_DEBUG=1 TFK_GEMM_HALF_TILES=$pol timeout 300 ./tfkaldi_b200/csrc/build/selftest_gemm bench 1 1 > gpurun_out/halftile_$pol.log 2>&1
  grep -E "BENCH|fused bwd|tfk gemm" gpurun_out/halftile_$pol.log | sort | uniq -c | sort -k2 | cut -c1-200 | head -40
done
