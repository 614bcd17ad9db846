#!/bin/bash
# same-box sweep of the decode GEMM's split-K cluster size (Falcon-7B [dense | 4h_to_h]: 36 tiles, K = 22720)
set -u; mkdir -p gpurun_out
for sp in 3 2 4 5 6 7 8 3; do
  B200W_DEBUG_SPLITS=1 B200W_DECODE_SPLITS=$sp timeout 240 python bench.py --decode-only > gpurun_out/dec_sp.json 2> gpurun_out/decode.err
  python -c "
import json; d=json.load(open('gpurun_out/dec_sp.json')); print('splits', $sp, 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
done
grep "K=22720" gpurun_out/decode.err | sort | uniq -c | head -3
