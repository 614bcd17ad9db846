#!/bin/bash
# final build at 4 GPUs (driver-style: defaults)
set -u; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== bench N=4"; timeout 220 $T --nproc-per-node 4 --master-port 29632 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/r2_bench_n4_final.json 2> gpurun_out/bench_n4f.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n4_final.json')); print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')}); print(d['roofline']['frac'], d.get('per_rank'))"; grep -E "b200w|Error|NCCL WARN" gpurun_out/bench_n4f.err | tail -5
