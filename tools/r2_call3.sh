#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== the two failing tests + attention after the dK/dV barrier removal"
timeout 600 python -m pytest tests/test_opt.py tests/test_infer_round2.py tests/test_attention.py tests/test_onchip_state.py tests/test_engine.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "opt grad|passed|failed|FAILED|Error|assert|rel_err|greedy|blocks" | tail -40
echo "== stress"; timeout 240 python tools/stress_attn.py --iters 6000 --noise 2>&1 | tail -2
echo "== attention A/B (prev = call-2 build, cur = dK/dV without the 512-thread barrier)"
timeout 500 bash tools/ab_attn.sh ab_prev/libb200w_prev.so 2 2>&1 | grep -vE "^==PROF|ncu-rep" | tail -12
echo "== decode timeline (2 steps)"
B200W_PROFILE_DECODE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_decode_launches_v1.csv python bench.py --decode-only > gpurun_out/ncu_decode.log 2>&1; wc -l gpurun_out/r2_decode_launches_v1.csv
echo "== ncu full: decode gemm + tc attention"
B200W_PROFILE_DECODE=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_decode|decode_attn_tc" -c 6 -f -o gpurun_out/r2_k_decode python bench.py --decode-only > gpurun_out/ncu_decode_full.log 2>&1; tail -2 gpurun_out/ncu_decode_full.log
