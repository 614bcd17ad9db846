#!/bin/bash
# Same-box A/B of two builds of the library (numbers from different boxes differ by their power-capped
# clocks). Runs on the GPU box's scratch copy only: the shipped library is swapped and restored.
#   usage: tools/ab_gemm.sh <other.so> [rounds]
# To build <other.so>: git show <rev>:runbooks_b200/csrc/gemm.cu > /tmp/prev.cu; nvcc (flags of runbooks_b200/build.py)
# -Irunbooks_b200/csrc -Iinclude -c /tmp/prev.cu -o /tmp/prev.o; link it with the other objects of
# runbooks_b200/build/ into a second .so inside the repo tree so that it travels to the GPU box
# (round 1 used rev e6ad6b9).
set -u
OTHER=$1; ROUNDS=${2:-2}
LIB=runbooks_b200/libb200w.so
mkdir -p gpurun_out
cp $LIB /tmp/cur.so
for i in $(seq 1 $ROUNDS); do
  cp $OTHER $LIB; echo "== prev $i"; timeout 200 python tools/perf_probe.py --only gemm --out gpurun_out/ab_prev_$i.json | grep -E "clocks" 
  cp /tmp/cur.so $LIB; echo "== cur $i"; timeout 200 python tools/perf_probe.py --only gemm --out gpurun_out/ab_cur_$i.json | grep -E "clocks"
done
cp /tmp/cur.so $LIB
python - <<'PY'
import json, glob
def load(tag):
    runs = [json.load(open(f)) for f in sorted(glob.glob(f"gpurun_out/ab_{tag}_*.json"))]
    return runs
prev, cur = load("prev"), load("cur")
keys = [k for k in cur[0] if k.startswith("gemm")]
best = lambda runs, k: max(r[k]["tflops"] for r in runs)
print(f"{'shape':30s} {'prev':>8s} {'cur':>8s}  cur/prev")
for k in keys:
    a, b = best(prev, k), best(cur, k)
    print(f"{k:30s} {a:8.1f} {b:8.1f}  {b / a:6.3f}")
print("clocks prev", [r["clocks"]["sm_mhz"] for r in prev], "cur", [r["clocks"]["sm_mhz"] for r in cur])
PY
