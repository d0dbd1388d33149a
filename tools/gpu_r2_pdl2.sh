#!/bin/bash
# round 2: programmatic dependent launch on the small main-stream kernels, again, now that the strip kernels are one wave
TAG=${1:-r2p}
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'P'
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
print(sys.argv[2], round(d["value"]), round(d["ms_per_step"], 4), d["timing"]["windows_ms_per_step"], "e2e", round(d["e2e"]["value"]))
P
}
for rep in 1 2; do
  for pdl in 0 1; do
    TFK_PDL_SMALL=$pdl timeout 600 python bench.py --config c4 --steps 50 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_c4_pdl${pdl}_${rep}_${TAG}.json 2>/dev/null
    show gpurun_out/bench_c4_pdl${pdl}_${rep}_${TAG}.json "c4 pdl_small=$pdl rep$rep"
  done
done
for pdl in 0 1; do
  TFK_PDL_SMALL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_c2_pdl${pdl}_${TAG}.json 2>/dev/null
  show gpurun_out/bench_c2_pdl${pdl}_${TAG}.json "c2 pdl_small=$pdl"
done
echo "== tests under TFK_PDL_SMALL=1"
TFK_PDL_SMALL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_feeder.py -m gpu -q -x > gpurun_out/pytest_pdl1_${TAG}.log 2>&1; echo exit=$?; tail -2 gpurun_out/pytest_pdl1_${TAG}.log
