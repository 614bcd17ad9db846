#!/bin/bash
# dQ-kernel v9 gate: parity first; the bench only counts if parity is green.
set -u
mkdir -p gpurun_out
python tools/run_gpu_tests.py tests/test_attention.py tests/test_engine.py tests/test_ops.py > gpurun_out/v9_tests.log 2>&1
RC=$?
grep -E "green|FAIL|failing|Error" gpurun_out/v9_tests.log | tail -8
if [ $RC -eq 0 ]; then GREEN=1; else GREEN=0; fi
echo "parity green: $GREEN"
timeout 300 python tools/perf_probe.py --only attn --out gpurun_out/probe_attn_v9.json 2>&1 | grep -E "attention|clocks"
if [ "$GREEN" = "1" ]; then
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_v9b.json 2> gpurun_out/bench_v9b.err
  python -c "
import json; b=json.load(open('gpurun_out/bench_v9b.json')); print({k:b[k] for k in ('value','ms_per_step','clocks')}, b['roofline']['achieved'])"
  tail -3 gpurun_out/bench_v9b.err
else
  echo "SKIPPING bench: parity not green"
fi
tools/ab_gemm.sh ab_prev/libb200w_prev.so 2
