#!/bin/bash
# strip kernels with early loads; C4 launch list; C4/C2 lines
TAG=${1:-r2l}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_${TAG}.log 2>&1; echo exit=$?; tail -3 gpurun_out/pytest_${TAG}.log
one() { cfg=$1; name=$2; shift; shift
  env "$@" timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_${cfg}_${name}_${TAG}.json 2> gpurun_out/bench_${cfg}_${name}_${TAG}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${cfg}_${name}_${TAG}.json"))
print("$cfg $name", "%.3e" % d["value"], d["timing"]["windows_ms_per_step"], d["roofline"]["per_step_us_by_kernel_class"], d["hbm_kernels"].get("bn"))
PY
}
one c4 default TFK_X=1
one c4 gemmpdl0 TFK_PDL=0
one c2 default TFK_X=1
one c2 gemmpdl0 TFK_PDL=0
one c4 default2 TFK_X=1
echo "== ncu launch list c4"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_c4_${TAG}.csv python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline --no-parity-mode > gpurun_out/ncu_c4_${TAG}.log 2>&1; echo exit=$?
python tools/ncu_summarize.py launches gpurun_out/launches_c4_${TAG}.csv | tail -12
