#!/bin/bash
# Same-box per-kernel A/B of the attention kernels: ncu launch durations (serialised, cold cache,
# --clock-control none) of two library builds, interleaved. Then one --set full capture of the dQ kernel.
#   usage: tools/ab_attn.sh <other.so> [rounds]
# To build <other.so>: git show <rev>:runbooks_b200/csrc/attention.cu > /tmp/prev.cu; nvcc (flags of runbooks_b200/build.py)
# -Irunbooks_b200/csrc -Iinclude -c /tmp/prev.cu -o /tmp/prev.o; link it with the other objects of
# runbooks_b200/build/ into a second .so inside the repo tree so that it travels to the GPU box
# (round 1 used rev c3287da).
set -u
OTHER=$1; ROUNDS=${2:-2}
LIB=runbooks_b200/libb200w.so
mkdir -p gpurun_out
cp $LIB /tmp/cur.so
one() { timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_ --csv \
          --log-file gpurun_out/ab_attn_$1.csv python tools/perf_probe.py --only attn --out gpurun_out/ab_attn_probe_$1.json > gpurun_out/ab_attn_$1.log 2>&1; }
for i in $(seq 1 $ROUNDS); do
  cp $OTHER $LIB; one prev_$i
  cp /tmp/cur.so $LIB; one cur_$i
done
cp /tmp/cur.so $LIB
python - <<'PY'
import csv, glob, collections, statistics
def load(tag):
    d = collections.defaultdict(list)
    for f in sorted(glob.glob(f"gpurun_out/ab_attn_{tag}_*.csv")):
        rows = [r for r in csv.reader(open(f, errors="ignore")) if len(r) > 5]
        hdr = next(r for r in rows if "Kernel Name" in r)
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        for r in rows[rows.index(hdr) + 1:]:
            v = float(r[vi].replace(",", "")); u = r[ui]
            v = v / 1e3 if u in ("ns", "nsecond") else v * (1e3 if u in ("ms", "msecond") else 1.0)
            d[r[ki].split("(")[0]].append(v)
    return d
p, c = load("prev"), load("cur")
print(f"{'kernel':28s} {'prev us':>9s} {'cur us':>9s} cur/prev   (median over launches; n)")
for k in sorted(c):
    a = statistics.median(p[k]) if p.get(k) else float('nan'); b = statistics.median(c[k])
    print(f"{k:28s} {a:9.1f} {b:9.1f} {b / a:7.3f}   n={len(c[k])}")
PY
for k in attn_bwd_dq attn_bwd_dkdv attn_fwd; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/v9_$k \
    python tools/perf_probe.py --only attn --out /tmp/x.json > gpurun_out/ncu_v9_$k.log 2>&1; tail -1 gpurun_out/ncu_v9_$k.log
done
ls -la gpurun_out/*.ncu-rep
