#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
