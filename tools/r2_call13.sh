#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== tests touched by the wire-copy epilogue and the pipelined rmsnorm_bwd"; timeout 900 python -m pytest tests/test_ops.py tests/test_gemm.py tests/test_engine.py tests/test_multi_gpu.py tests/test_opt.py tests/test_falcon_train.py -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -8
echo "== ops probe (includes ~20 us host call + sync overhead per launch)"; timeout 200 python tools/ops_probe.py | tee gpurun_out/r2_ops_probe_v1.json
echo "== rmsnorm_bwd under ncu"; timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:rmsnorm_bwd_kernel -c 6 python tools/ops_probe.py 2>&1 | grep -E "rmsnorm_bwd_kernel|gpu__time|dram__bytes" | head -24
echo "== bench N=1"; timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-decode > gpurun_out/r2_bench_n1_v16.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_v16.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['roofline']['frac'], d['roofline']['achieved'])"; tail -3 gpurun_out/bench.err
