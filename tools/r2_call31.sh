#!/bin/bash
set -u; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -E "device bytes|passed|failed|Error|assert" | tail -12
