#!/bin/bash
# 8 GPUs: the default data-parallel path (overlapped per-matrix bf16 all-reduce, NCCL capped at 16 CTAs, GEMMs on SMs-16)
# and the same without the SM partition. Each leg is bounded: a hang costs 8 x its timeout.
set -u; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== bench N=8 default"; NCCL_DEBUG=WARN timeout 170 $T --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?"; tail -c 600 gpurun_out/r2_bench_n8.json; grep -E "b200w:|bench.py:|Error|NCCL WARN" gpurun_out/bench_n8.err | tail -3
echo "== bench N=8 no SM partition (NCCL default CTAs)"; B200W_AR_SM_RESERVE=0 timeout 170 $T --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_bench_n8_nopart.json 2> gpurun_out/bench_n8b.err; echo "rc=$?"; tail -c 300 gpurun_out/r2_bench_n8_nopart.json; grep -E "b200w:|bench.py:|Error" gpurun_out/bench_n8b.err | tail -3
