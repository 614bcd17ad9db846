#!/bin/bash
# same box: bands off / bands on without the long-K square waves / bands on with them
set -u; mkdir -p gpurun_out
q() { for k in gemm_fwd gemm_dgrad gemm_wgrad_acc; do TOKENS=8192 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:gemm_bf16 -s 2 -c 1 python tools/one_kernel.py $k 3 2>&1 | grep -E "dram__bytes|gpu__time" | awk -v k=$k '{printf "%s %s %s; ", k, $3, $2} END {print ""}'; done; }
b() { timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-decode > gpurun_out/bench_x.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); print('   bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"; }
echo "-- bands off"; export B200W_GEMM_RASTER_BANDS=0; q; b
echo "-- bands on, long-K square off"; export B200W_GEMM_RASTER_BANDS=1 B200W_GEMM_LONGK_SQUARE=0; q; b
echo "-- bands on, long-K square on"; export B200W_GEMM_LONGK_SQUARE=1; q; b
echo "-- bands on, long-K square off (again)"; export B200W_GEMM_LONGK_SQUARE=0; b
