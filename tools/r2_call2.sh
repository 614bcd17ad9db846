#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench n1"; timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_v12.json 2> gpurun_out/bench.err; tail -c 1200 gpurun_out/r2_bench_n1_v12.json; tail -3 gpurun_out/bench.err
