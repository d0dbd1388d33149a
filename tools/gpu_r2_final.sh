#!/bin/bash
# round 2, closing visit: what the driver runs (tests, smoke, both bench arms) + every other bench line + ncu evidence of the final build
TAG=${1:-r2z}
mkdir -p gpurun_out
if [ -z "$SKIP_SELFTEST" ]; then echo "== selftest_gemm full"; timeout 400 ./tfkaldi_b200/csrc/build/selftest_gemm full > gpurun_out/selftest_full_${TAG}.log 2>&1; echo exit=$?; grep -E "FAIL|BENCH|fused bwd|selftest_gemm:" gpurun_out/selftest_full_${TAG}.log | cut -c1-170 | tail -24; fi
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_${TAG}.log 2>&1; echo exit=$?; tail -3 gpurun_out/pytest_${TAG}.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; cut -c1-400 gpurun_out/bench_ref_${TAG}.json
echo "== bench (driver flags)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_${TAG}.json 2> gpurun_out/bench_c2_${TAG}.err; echo exit=$?; tail -2 gpurun_out/bench_c2_${TAG}.err; cut -c1-600 gpurun_out/bench_c2_${TAG}.json
echo "== bench (default flags)"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_c2_default_${TAG}.json 2>/dev/null; cut -c1-300 gpurun_out/bench_c2_default_${TAG}.json
for cfg in c4 feed; do echo "== bench $cfg"; timeout 900 python bench.py --config $cfg --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${cfg}_${TAG}.json 2> gpurun_out/bench_${cfg}_${TAG}.err; echo exit=$?; tail -2 gpurun_out/bench_${cfg}_${TAG}.err; cut -c1-500 gpurun_out/bench_${cfg}_${TAG}.json; done
for prec in bf16 bf16x3; do echo "== bench c5 $prec"; timeout 900 python bench.py --config c5 --precision $prec --steps 100 > gpurun_out/bench_c5_${prec}_${TAG}.json 2> gpurun_out/bench_c5_${prec}_${TAG}.err; echo exit=$?; tail -2 gpurun_out/bench_c5_${prec}_${TAG}.err; cut -c1-700 gpurun_out/bench_c5_${prec}_${TAG}.json; done
echo "== ncu launch lists (c2, c4)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/launches_c2_${TAG}.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_c2_${TAG}.log 2>&1; echo exit=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file gpurun_out/launches_c4_${TAG}.csv python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_c4_${TAG}.log 2>&1; echo exit=$?
# the .ncu-rep files stay on the box (gpurun_out/ is capped at 64 MiB): only their summaries travel
echo "== ncu full: one C2 step (gemm, adam, softmax)"
timeout 900 ncu --set full --clock-control none -k regex:"tfk_gemm2|adam_kernel|softmax_ce" -s 80 -c 24 -o /tmp/prof_c2_${TAG} -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_c2_full_${TAG}.log 2>&1; echo exit=$?
python tools/ncu_summarize.py full /tmp/prof_c2_${TAG}.ncu-rep > gpurun_out/ncu_full_summary_c2_${TAG}.txt 2>&1; echo exit=$?
python tools/ncu_traffic.py /tmp/prof_c2_${TAG}.ncu-rep c2:bf16:8192 "ncu --set full --clock-control none, bench.py --steps 3 --warmup 3 (C2, bf16), visit ${TAG}: 14 consecutive tfk_gemm2_kernel launches = one step" > gpurun_out/ncu_traffic_${TAG}.json 2> gpurun_out/ncu_traffic_${TAG}.err; echo exit=$?
echo "== ncu full: the strip / small kernels of a C4 step"
timeout 900 ncu --set full --clock-control none -k regex:"bn_|colsum|softmax_ce|split_f32|accum_loss" -s 60 -c 20 -o /tmp/prof_c4_small_${TAG} -f python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_c4_full_${TAG}.log 2>&1; echo exit=$?
python tools/ncu_summarize.py full /tmp/prof_c4_small_${TAG}.ncu-rep > gpurun_out/ncu_full_summary_c4_small_${TAG}.txt 2>&1; echo exit=$?
du -sh gpurun_out; ls -la gpurun_out | tail -12
