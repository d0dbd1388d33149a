#!/bin/bash
# last 2-GPU check of the final build: the data-parallel tests (quick ones) and the bench line, whose e2e leg now goes through tfk_last_loss
TAG=${1:-r2q}
mkdir -p gpurun_out
echo "== pytest tests/test_gpu_dp.py (without the full-size test)"; timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x -k "not full_size" > gpurun_out/pytest_dp_${TAG}.log 2>&1; echo exit=$?; tail -3 gpurun_out/pytest_dp_${TAG}.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_2gpu_${TAG}.json 2> gpurun_out/bench_2gpu_${TAG}.err; echo exit=$?; tail -2 gpurun_out/bench_2gpu_${TAG}.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_2gpu_${TAG}.json") if l.startswith("{")][-1])
print("2gpu value %.4e" % d["value"], d["timing"]["windows_ms_per_step"], "e2e %.4e" % d["e2e"]["value"], d["e2e"].get("windows_ms_per_step"))
PY
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity-mode > gpurun_out/bench_1gpu_${TAG}.json 2>/dev/null; python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_1gpu_${TAG}.json") if l.startswith("{")][-1]); print("1gpu value %.4e" % d["value"], d["timing"]["windows_ms_per_step"], "e2e %.4e" % d["e2e"]["value"])
PY
