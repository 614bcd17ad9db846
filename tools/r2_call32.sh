#!/bin/bash
# activation recomputation at Llama-2-7B size: memory and step time, micro-batch 2 and 4 (NOT the benchmark configuration)
set -u; mkdir -p gpurun_out
for mb in 2 4; do
  timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu --no-decode --recompute --micro-batch $mb > gpurun_out/bench_rec_mb$mb.json 2> gpurun_out/bench.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_rec_mb$mb.json')); print('recompute mb', $mb, d['value'], d['ms_per_step'], 'device_gb', d['device_gb'], 'gemm frac', d['roofline']['frac'], d['clocks']['sm_mhz'], d['loss'])"; tail -2 gpurun_out/bench.err | cut -c1-200
done
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --no-decode > gpurun_out/bench_norec.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_norec.json')); print('no recompute mb 2', d['value'], d['ms_per_step'], 'device_gb', d['device_gb'], d['clocks']['sm_mhz'], d['loss'])"
