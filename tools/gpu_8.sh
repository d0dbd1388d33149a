#!/bin/bash
# 8-GPU visit (gpurun --gpus 8): the bench line in the default data-parallel mode (GEMM -> reduce-scatter and
# update -> all-gather fused over NVLink peer memory) and, for comparison, with NCCL all-reduce.
N=${1:-8}
mkdir -p gpurun_out
for mode in fused allreduce; do
  echo "== bench --gpus $N, TFK_DP_MODE=$mode"
  if [ $mode = fused ]; then unset TFK_DP_MODE; else export TFK_DP_MODE=$mode; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu_${mode}.json 2> gpurun_out/bench_${N}gpu_${mode}.err
  echo "exit=$?"; tail -3 gpurun_out/bench_${N}gpu_${mode}.err | cut -c1-300; cut -c1-1500 gpurun_out/bench_${N}gpu_${mode}.json
done
