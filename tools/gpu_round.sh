#!/bin/bash
# One GPU-box session: driver-equivalent checks + bench + ncu evidence. Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench N=1"; timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
echo "== ncu launch list (same command, 1 sequence per step to bound replay time)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3400 -c 1000 --csv \
  --log-file gpurun_out/launches.csv python bench.py --gpus 1 --steps 1 --warmup 3 --per-device-batch 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches.csv
echo "== ncu --set full: gemm, attention fwd, attention bwd (4-layer model, same widths)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 60 -c 3 -f -o gpurun_out/prof_gemm \
  python bench.py --gpus 1 --steps 1 --warmup 3 --per-device-batch 1 --layers 2 --no-cpu > gpurun_out/ncu_gemm.log 2>&1; tail -1 gpurun_out/ncu_gemm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 8 -c 2 -f -o gpurun_out/prof_attn \
  python bench.py --gpus 1 --steps 1 --warmup 3 --per-device-batch 1 --layers 2 --no-cpu > gpurun_out/ncu_attn.log 2>&1; tail -1 gpurun_out/ncu_attn.log
ls -la gpurun_out | head -30
