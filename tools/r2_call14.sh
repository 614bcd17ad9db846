#!/bin/bash
# same-box A/B: decode split-K sized to the clusters the chip holds at once vs the widest split
set -u; mkdir -p gpurun_out
echo "== infer tests"; timeout 600 python -m pytest tests/test_infer.py tests/test_infer_round2.py tests/test_server.py tests/test_server_round2.py tests/test_gemm.py -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -4
echo "== decode, split fitted"; B200W_DEBUG_SPLITS=1 timeout 240 python bench.py --decode-only > gpurun_out/r2_decode_v6_fit.json 2> gpurun_out/decode.err; cut -c1-330 gpurun_out/r2_decode_v6_fit.json; grep "decode GEMM" gpurun_out/decode.err | sort | uniq -c | head -12
echo "== decode, widest split (previous)"; B200W_DECODE_SPLIT_FIT=0 timeout 240 python bench.py --decode-only > gpurun_out/r2_decode_v6_widest.json 2> gpurun_out/decode.err; cut -c1-330 gpurun_out/r2_decode_v6_widest.json
echo "== decode, split fitted (again)"; timeout 240 python bench.py --decode-only > gpurun_out/r2_decode_v6_fit2.json 2> gpurun_out/decode.err; cut -c1-330 gpurun_out/r2_decode_v6_fit2.json
