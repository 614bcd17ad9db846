#!/bin/bash
set -u; mkdir -p gpurun_out
echo "== pytest -m gpu (4 processes, one file per process at a time, no -x)"
timeout 900 python -m pytest tests -m gpu -q -n 4 --dist loadfile -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/r2_gputests.txt | tail -25
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== decode bench"; timeout 240 python tools/decode_bench.py > gpurun_out/r2_decode_v1.json 2> gpurun_out/decode.err; cat gpurun_out/r2_decode_v1.json; tail -2 gpurun_out/decode.err
echo "== bench n1"; timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_n1_v12.json 2> gpurun_out/bench.err; tail -c 1200 gpurun_out/r2_bench_n1_v12.json; tail -3 gpurun_out/bench.err
echo "== decode timeline"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 450 --csv --log-file gpurun_out/r2_decode_launches_v1.csv \
  python tools/decode_bench.py > gpurun_out/ncu_decode.log 2>&1; wc -l gpurun_out/r2_decode_launches_v1.csv
