#!/bin/bash
# bench lines for C2 / C4 / C5 + ncu launch list + ncu full capture of the CTA-pair GEMM
TAG=${1:-p}
mkdir -p gpurun_out
for cfg in c2 c4 c5; do
  echo "== bench $cfg"; timeout 900 python bench.py --config $cfg > gpurun_out/bench_${cfg}_${TAG}.json 2> gpurun_out/bench_${cfg}_${TAG}.err; tail -2 gpurun_out/bench_${cfg}_${TAG}.err; cut -c1-1800 gpurun_out/bench_${cfg}_${TAG}.json
done
echo "== bench c2 bf16x3"; timeout 900 python bench.py --precision bf16x3 --no-cpu-baseline > gpurun_out/bench_c2x3_${TAG}.json 2>/dev/null; cut -c1-400 gpurun_out/bench_c2x3_${TAG}.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
echo "== ncu full (gemm2 + adam + softmax)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tfk_gemm2|adam_kernel|softmax_ce" -s 60 -c 18 -o gpurun_out/prof_${TAG} -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log | cut -c1-200
ls -la gpurun_out | tail -12
