#!/bin/bash
# multi-GPU visit (gpurun --gpus N): data-parallel parity tests (incl. full size) and the bench line under the transports
N=${1:-2}; TAG=${2:-r2dp}
mkdir -p gpurun_out
echo "== pytest tests/test_gpu_dp.py"; timeout 1500 python -m pytest tests/test_gpu_dp.py -m gpu -q -rP > gpurun_out/pytest_dp_${N}_${TAG}.log 2>&1; echo exit=$?; tail -6 gpurun_out/pytest_dp_${N}_${TAG}.log; grep "full-size data parallel" gpurun_out/pytest_dp_${N}_${TAG}.log
run() {  # name, env...
  name=$1; shift
  echo "== bench --gpus $N [$name]"
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_${N}gpu_${name}_${TAG}.json 2> gpurun_out/bench_${N}gpu_${name}_${TAG}.err
  echo "exit=$?"; tail -2 gpurun_out/bench_${N}gpu_${name}_${TAG}.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${N}gpu_${name}_${TAG}.json") if l.startswith("{")][-1])  # (NCCL prints its version banner to stdout first)
    print("$name", "value %.3e" % d["value"], d["timing"]["windows_ms_per_step"], d["timing"]["per_step_ms_in_an_extra_window"], "e2e", d["e2e"]["windows_ms_per_step"], d["roofline"]["per_step_us_by_kernel_class"])
except Exception as e:
    print("no json", e)
PY
}
run serial TFK_X=1
run serial_nccl_small TFK_DP_SMALL=nccl
run serial_again TFK_X=1
run serial_nccl_small_again TFK_DP_SMALL=nccl
if [ "$N" != "2" ]; then run allreduce TFK_DP_MODE=allreduce; fi
echo "== bench --gpus 1 (same box)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_1gpu_${TAG}.json 2>/dev/null; python - <<PY
import json
d=json.load(open("gpurun_out/bench_1gpu_${TAG}.json")); print("1gpu value %.3e" % d["value"], d["timing"]["windows_ms_per_step"])
PY
