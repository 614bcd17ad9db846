#!/bin/bash
# the driver's exact GPU test command (serial, one process)
set -u
( time timeout 320 python -m pytest tests/ -x -q -m gpu --durations=8 -p no:cacheprovider ) 2>&1 | tail -22
nvidia-smi --query-gpu=memory.used --format=csv,noheader
