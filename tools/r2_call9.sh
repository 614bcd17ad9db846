#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -6
echo "== decode, tiled weights"; timeout 240 python bench.py --decode-only > gpurun_out/r2_decode_v5_tiled.json 2> gpurun_out/decode.err; cut -c1-420 gpurun_out/r2_decode_v5_tiled.json; tail -2 gpurun_out/decode.err
echo "== decode, row-major weights through TMA boxes (same box)"; B200W_DECODE_TILED=0 timeout 240 python bench.py --decode-only > gpurun_out/r2_decode_v5_rowmajor.json 2> gpurun_out/decode.err; cut -c1-420 gpurun_out/r2_decode_v5_rowmajor.json; tail -2 gpurun_out/decode.err
echo "== bench N=1 (bias/ReLU epilogue build)"; timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-decode > gpurun_out/r2_bench_n1_v15.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_v15.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']['frac'], d['roofline']['achieved'])"; tail -3 gpurun_out/bench.err
echo "== decode timeline"; B200W_PROFILE_DECODE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_decode_launches_v5.csv python bench.py --decode-only > gpurun_out/ncu_decode.log 2>&1; wc -l gpurun_out/r2_decode_launches_v5.csv
