#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench line, ncu launch list, ncu full capture of the GEMM.
# usage: tools/gpu_round.sh [tag]      (outputs under gpurun_out/)
TAG=${1:-r1}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json
echo "== bench: other lines (C4, C5 decode, feeder from Kaldi archives)"
for cfg in c4 c5 feed; do
  timeout 600 python bench.py --config $cfg --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${cfg}_${TAG}.json 2> gpurun_out/bench_${cfg}_${TAG}.err
  cut -c1-400 gpurun_out/bench_${cfg}_${TAG}.json
done
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_bench_${TAG}.log | cut -c1-300
echo "== ncu full (gemm)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tfk_gemm -s 42 -c 6 -o gpurun_out/prof_gemm_${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-300
ls -la gpurun_out | head -30
