#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== decode (cluster split-K, separate merge)"; timeout 240 python bench.py --decode-only > gpurun_out/r2_decode_v3.json 2> gpurun_out/decode.err; cat gpurun_out/r2_decode_v3.json; tail -2 gpurun_out/decode.err
echo "== bench N=1 micro-batch 2"; timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --no-decode --micro-batch 2 > gpurun_out/r2_bench_n1_mb2.json 2> gpurun_out/bench.err; tail -c 700 gpurun_out/r2_bench_n1_mb2.json; tail -3 gpurun_out/bench.err
echo "== bench N=1 micro-batch 1 (same box)"; timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --no-decode > gpurun_out/r2_bench_n1_v13.json 2> gpurun_out/bench.err; tail -c 700 gpurun_out/r2_bench_n1_v13.json; tail -3 gpurun_out/bench.err
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -8
echo "== full-step timeline"; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_step_launches.csv python tools/step_timeline.py > gpurun_out/ncu_step.log 2>&1; grep STEP_TIMELINE gpurun_out/ncu_step.log; wc -l gpurun_out/r2_step_launches.csv
