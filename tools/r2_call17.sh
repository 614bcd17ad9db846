#!/bin/bash
set -u; mkdir -p gpurun_out
for sp in 6 4 8 5 6 4; do
  B200W_DEBUG_SPLITS=1 B200W_DECODE_SPLITS=$sp timeout 240 python bench.py --decode-only > gpurun_out/dec_sp.json 2> gpurun_out/decode.err
  python -c "
import json; d=json.load(open('gpurun_out/dec_sp.json')); print('splits', $sp, 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
done
grep "K=22720" gpurun_out/decode.err | sort | uniq -c | head -3
