#!/bin/bash
# round 2: batch-norm strip kernels in one wave (A/B on one box against the library built from the previous kernels.cu)
TAG=${1:-r2s}
mkdir -p gpurun_out
echo "== pytest: batch-norm paths"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zfullsize.py tests/test_gpu_feeder.py -m gpu -q -x -k "bn or trajectory or single_step or c4 or fused or feeder or forward_eval or unit_entry" > gpurun_out/pytest_${TAG}.log 2>&1; echo exit=$?; tail -3 gpurun_out/pytest_${TAG}.log
L=tfkaldi_b200/libtfkaldi_b200.so
cp $L /tmp/new.so
for rep in 1 2; do
  for which in new prev; do
    if [ $which = prev ]; then cp tfkaldi_b200/libtfkaldi_b200_prev.so $L; else cp /tmp/new.so $L; fi
    timeout 600 python bench.py --config c4 --steps 50 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_c4_${which}${rep}_${TAG}.json 2>/dev/null
    python - gpurun_out/bench_c4_${which}${rep}_${TAG}.json $which$rep <<'P'
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
print(sys.argv[2], "c4", round(d["value"]), d["ms_per_step"], d["timing"]["windows_ms_per_step"], d["hbm_kernels"]["bn"])
P
  done
done
cp /tmp/new.so $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/launches_c4_${TAG}.csv python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_c4_${TAG}.log 2>&1; echo exit=$?
python tools/ncu_summarize.py launches gpurun_out/launches_c4_${TAG}.csv | head -12
