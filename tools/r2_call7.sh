#!/bin/bash
# 8 GPUs, micro-batch 2 (T = 8192 per GEMM): does it fit next to the 13.5 GB wire copy and NCCL's buffers, and what does it buy?
set -u; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(for i in $(seq 1 40); do nvidia-smi --query-gpu=index,memory.used,memory.total --format=csv,noheader,nounits | tr '\n' ';'; echo; sleep 3; done) > gpurun_out/r2_n8_mb2_mem.txt 2>&1 &
MON=$!
echo "== bench N=8 micro-batch 2"; timeout 200 $T --nproc-per-node 8 --master-port 29613 bench.py --gpus 8 --steps 3 --warmup 3 --micro-batch 2 > gpurun_out/r2_bench_n8_mb2.json 2> gpurun_out/bench_n8c.err; echo "rc=$?"; tail -c 500 gpurun_out/r2_bench_n8_mb2.json; grep -E "b200w|bench.py:|Error|NCCL WARN|memory" gpurun_out/bench_n8c.err | tail -5
kill $MON 2>/dev/null
sort -t, -k2 -n -r gpurun_out/r2_n8_mb2_mem.txt | head -2
