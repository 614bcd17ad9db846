#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== decode default (carve-out max, fitted split)"; B200W_DEBUG_SPLITS=1 timeout 240 python bench.py --decode-only > gpurun_out/r2_decode_v8.json 2> gpurun_out/decode.err; cut -c1-330 gpurun_out/r2_decode_v8.json; grep "decode GEMM" gpurun_out/decode.err | sort | uniq -c | head -4
for sp in 8 6 5; do
  B200W_DECODE_SPLITS=$sp timeout 240 python bench.py --decode-only > gpurun_out/dec_sp.json 2> gpurun_out/decode.err
  python -c "
import json; d=json.load(open('gpurun_out/dec_sp.json')); print('splits', $sp, 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
done
echo "== infer tests"; timeout 600 python -m pytest tests/test_infer.py tests/test_infer_round2.py tests/test_server.py tests/test_server_round2.py -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -3
echo "== ncu decode gemm"; B200W_PROFILE_DECODE=1 timeout 300 ncu --profile-from-start off --set full --clock-control none -k regex:"gemm_decode" -c 4 -f -o gpurun_out/r2_k_decode_v8 python bench.py --decode-only > gpurun_out/ncu_decode_full.log 2>&1; tail -1 gpurun_out/ncu_decode_full.log
