#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gemm.py tests/test_engine.py -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -2
run() { TOKENS=8192 timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$4" -c 1 -f -o "gpurun_out/r2_k_$1_mb2b" python tools/one_kernel.py "$3" 3 > "gpurun_out/ncu_$1.log" 2>&1; tail -1 "gpurun_out/ncu_$1.log"; }
run gemm_wgrad_acc gemm_bf16 gemm_wgrad_acc 2
run gemm_fwd gemm_bf16 gemm_fwd 2
run gemm_dgrad gemm_bf16 gemm_dgrad 2
for i in 1 2; do timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-decode > gpurun_out/bench_b2.json 2> gpurun_out/bench.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_b2.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"; done
