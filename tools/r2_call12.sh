#!/bin/bash
# 8 GPUs at the defaults the driver will run (micro-batch 2, GEMM SM reserve 8 while the all-reduce overlaps)
set -u; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== bench N=8 defaults"; timeout 200 $T --nproc-per-node 8 --master-port 29614 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2_bench_n8_r8.json 2> gpurun_out/bench_n8d.err; echo "rc=$?"; tail -c 600 gpurun_out/r2_bench_n8_r8.json; grep -E "b200w|bench.py:|Error|NCCL WARN|memory" gpurun_out/bench_n8d.err | tail -5
