#!/bin/bash
# Attention what-if variants (timing only) + the pair row-max exchange (real numerics) against the current build, same box
set -u; mkdir -p gpurun_out
bash tools/ab_variants.sh "noexp noother onecta nodp pairmax" 1 2>&1 | tail -8
echo "== correctness of the pairmax build"; cp runbooks_b200/libb200w.so /tmp/cur_keep.so; cp ab_prev/lib_pairmax.so runbooks_b200/libb200w.so
timeout 600 python -m pytest tests/test_attention.py tests/test_onchip_state.py tests/test_engine.py tests/test_opt.py -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -4
cp /tmp/cur_keep.so runbooks_b200/libb200w.so
