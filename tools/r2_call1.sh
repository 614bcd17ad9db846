#!/bin/bash
# round-2 GPU call 1: the per-stage dK/dV build on hardware, stress, sanitizer, decode + N=1 baselines, timelines
set -u; mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== stress"; timeout 240 python tools/stress_attn.py --iters 8000 --noise 2>&1 | tail -3
echo "== memcheck"; timeout 240 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/r2_memcheck.txt 2>&1; tail -4 gpurun_out/r2_memcheck.txt
echo "== racecheck"; timeout 240 compute-sanitizer --tool racecheck python tools/sanitize_small.py > gpurun_out/r2_racecheck.txt 2>&1; tail -4 gpurun_out/r2_racecheck.txt
echo "== decode"; timeout 240 python tools/decode_bench.py > gpurun_out/r2_decode_v0.json 2> gpurun_out/decode.err; cat gpurun_out/r2_decode_v0.json; tail -2 gpurun_out/decode.err
echo "== bench n1"; timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_v11.json 2> gpurun_out/bench.err; tail -c 900 gpurun_out/r2_bench_n1_v11.json; tail -2 gpurun_out/bench.err
echo "== full-step timeline"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20241 -c 6747 --csv \
  --log-file gpurun_out/r2_step_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_step.log 2>&1; wc -l gpurun_out/r2_step_launches.csv
echo "== decode timeline"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 700 --csv --log-file gpurun_out/r2_decode_launches.csv \
  python tools/decode_bench.py > gpurun_out/ncu_decode.log 2>&1; wc -l gpurun_out/r2_decode_launches.csv
run() { timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$4" -c 1 -f -o "gpurun_out/r2_k_$1" python tools/one_kernel.py "$3" 3 > "gpurun_out/ncu_$1.log" 2>&1; tail -1 "gpurun_out/ncu_$1.log"; }
run gemm_wgrad_acc gemm_bf16 gemm_wgrad_acc 2
run gemm_fwd gemm_bf16 gemm_fwd 2
run gemm_dgrad gemm_bf16 gemm_dgrad 2
ls -la gpurun_out/*.ncu-rep
