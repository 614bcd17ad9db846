#!/bin/bash
# band budget sweep (MB of re-read operand per band), one box
set -u; mkdir -p gpurun_out
for mb in 34 24 44 17 34; do
  echo "-- band budget $mb MB"
  for k in gemm_fwd gemm_wgrad_acc; do B200W_GEMM_BAND_MB=$mb TOKENS=8192 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:gemm_bf16 -s 2 -c 1 python tools/one_kernel.py $k 3 2>&1 | grep -E "dram__bytes|gpu__time" | awk -v k=$k '{printf "%s %s %s %s; ", k, $1, $3, $2} END {print ""}'; done
  B200W_GEMM_BAND_MB=$mb timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-decode > gpurun_out/bench_band.json 2> gpurun_out/bench.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_band.json')); print('budget', $mb, d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
