#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== opt gradients (verbose)"; timeout 600 python -m pytest tests/test_opt.py::test_opt_gradients_match_hf -m gpu -q -s --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -40
echo "== falcon"; timeout 900 python -m pytest tests/test_falcon_train.py tests/test_worker_falcon.py "tests/test_ops.py::test_gelu_and_its_backward" -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v "^$" | tail -40
