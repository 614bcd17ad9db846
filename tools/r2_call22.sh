#!/bin/bash
# final build at 8 GPUs (driver-style: defaults)
set -u; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== bench N=8"; timeout 110 $T --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/bench_n8f.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n8_final.json')); print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')}); print(d['roofline']['frac'], d.get('per_rank'))"; grep -E "b200w|Error|NCCL WARN" gpurun_out/bench_n8f.err | tail -5
