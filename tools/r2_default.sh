#!/bin/bash
set -u; mkdir -p gpurun_out
timeout 230 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_default.json')); print({k: d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','clocks')}); print(d['e2e'], d['roofline']['frac'], d['roofline']['traffic']); print(d['decode']['ms_per_step'], d['decode']['roofline']['frac']); print({k: d['cpu_baseline'][k] for k in ('value','cores','timed_wall_s')})"; tail -2 gpurun_out/bench.err | cut -c1-200
