#!/bin/bash
# banded tile raster: correctness, DRAM bytes per launch (ncu) and the step (bench), bands on vs off, one box
set -u; mkdir -p gpurun_out
echo "== gemm tests"; timeout 600 python -m pytest tests/test_gemm.py tests/test_engine.py -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -2
for on in 1 0; do for k in gemm_fwd gemm_dgrad gemm_wgrad_acc; do
  echo "-- bands=$on $k"; B200W_GEMM_RASTER_BANDS=$on TOKENS=8192 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16 -s 2 -c 1 python tools/one_kernel.py $k 3 2>&1 | grep -E "dram__bytes|gpu__time" | awk '{printf "%s %s %s; ", $1, $3, $2} END {print ""}'
done; done
for on in 1 0 1 0; do
  B200W_GEMM_RASTER_BANDS=$on timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-decode > gpurun_out/bench_bands$on.json 2> gpurun_out/bench.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_bands$on.json')); print('bands', $on, d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
done
