#!/bin/bash
# Round-end validation on one B200: what the driver runs (pytest -m gpu, smoke, bench both arms) + the sanitizer
# pass over the round-2 kernels + the launch timeline of one full step at the default micro-batch.
set -u; mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== memcheck (tiny shapes)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r2_memcheck_v2.log 2>&1; echo "rc=$?"; grep -E "ok|ERROR SUMMARY|Invalid|Error" gpurun_out/r2_memcheck_v2.log | tail -14
echo "== bench N=1 as the driver runs it"; timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_final.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']['frac'], d['roofline']['achieved'], d['roofline']['share_of_step']); print(d['decode']['ms_per_step'], d['decode']['roofline']['frac'], d['decode']['prefill']['tokens_per_s']); print({k: d['cpu_baseline'][k] for k in ('value','cores','spread_max_over_min')})"; tail -3 gpurun_out/bench.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/r2_bench_ref_final.json; tail -2 gpurun_out/bench_ref.err
echo "== full-step timeline (micro-batch 2)"; timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_step_launches_mb2.csv python tools/step_timeline.py > gpurun_out/ncu_step.log 2>&1; grep STEP_TIMELINE gpurun_out/ncu_step.log; wc -l gpurun_out/r2_step_launches_mb2.csv
