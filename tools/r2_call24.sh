#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== correctness (delta pre-multiplied by the scale)"; timeout 600 python -m pytest tests/test_attention.py tests/test_onchip_state.py tests/test_engine.py tests/test_opt.py tests/test_falcon_train.py -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -3
bash tools/ab_variants.sh "before" 2 2>&1 | tail -6
