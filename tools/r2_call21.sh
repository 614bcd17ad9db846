#!/bin/bash
set -u; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_opt.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "opt-125m|passed|failed|Error" | tail -8
