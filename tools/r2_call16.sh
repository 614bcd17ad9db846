#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== infer tests"; timeout 600 python -m pytest tests/test_infer.py tests/test_infer_round2.py tests/test_server.py tests/test_server_round2.py tests/test_worker_opt.py tests/test_worker_falcon.py -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -3
echo "== decode (default = fitted split)"; B200W_DEBUG_SPLITS=1 timeout 240 python bench.py --decode-only > gpurun_out/r2_decode_v7.json 2> gpurun_out/decode.err; cut -c1-330 gpurun_out/r2_decode_v7.json; grep "decode GEMM" gpurun_out/decode.err | sort | uniq -c | head -4
echo "== decode timeline"; B200W_PROFILE_DECODE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_decode_launches_v7.csv python bench.py --decode-only > gpurun_out/ncu_decode.log 2>&1; wc -l gpurun_out/r2_decode_launches_v7.csv
echo "== bench N=1 default (decode first, cpu baseline)"; timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/r2_bench_n1_v17.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_v17.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']['frac'], d['roofline']['achieved']); print(d['decode']['ms_per_step'], d['decode']['roofline']['frac']); print(d.get('cpu_baseline'))"; tail -3 gpurun_out/bench.err
