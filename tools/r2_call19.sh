#!/bin/bash
# 2 GPUs: multi-rank parity tests with the wire-copy epilogue, then the bench at N = 2 (per-rank GEMM rates in the line)
set -u; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== multi-GPU tests"; timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
echo "== bench N=2"; timeout 240 $T --nproc-per-node 2 --master-port 29621 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2_bench_n2_v2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_v2.json')); print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')}); print(d['roofline']['frac'], d.get('per_rank'))"; grep -E "b200w|Error|NCCL WARN" gpurun_out/bench_n2.err | tail -5
