#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== falcon fine-tune + opt gradients (verbose)"; timeout 900 python -m pytest tests/test_falcon_train.py tests/test_worker_falcon.py tests/test_opt.py "tests/test_ops.py::test_gelu_and_its_backward" -m gpu -q -x -s -p no:cacheprovider 2>&1 | grep -v "^$" | tail -60
echo "== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -12
