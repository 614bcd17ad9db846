#!/bin/bash
# Same-box per-kernel timing of several library builds (ab_prev/lib_<name>.so, tools/build_variant.sh) against the
# current one: ncu launch durations (serialised, cold cache, --clock-control none) of tools/perf_probe.py --only attn,
# ROUNDS interleaved rounds.   usage: tools/ab_variants.sh "<name> <name> ..." [rounds] [kernel regex]
set -u
NAMES=$1; ROUNDS=${2:-2}; REGEX=${3:-attn_}
LIB=runbooks_b200/libb200w.so
mkdir -p gpurun_out; rm -f gpurun_out/abv_*.csv
cp $LIB /tmp/cur.so
one() { timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$REGEX --csv \
          --log-file gpurun_out/abv_$1.csv python tools/perf_probe.py --only attn --out gpurun_out/abv_probe.json > gpurun_out/abv_$1.log 2>&1; }
for i in $(seq 1 $ROUNDS); do
  for n in $NAMES; do cp ab_prev/lib_$n.so $LIB; one ${n}_$i; done
  cp /tmp/cur.so $LIB; one cur_$i
done
cp /tmp/cur.so $LIB
python - "$NAMES" <<'PY'
import csv, glob, collections, statistics, sys
names = sys.argv[1].split() + ["cur"]
def load(tag):
    d = collections.defaultdict(list)
    for f in sorted(glob.glob(f"gpurun_out/abv_{tag}_*.csv")):
        rows = [r for r in csv.reader(open(f, errors="ignore")) if len(r) > 5]
        hdr = next((r for r in rows if "Kernel Name" in r), None)
        if hdr is None:
            continue
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        for r in rows[rows.index(hdr) + 1:]:
            v = float(r[vi].replace(",", "")); u = r[ui]
            v = v / 1e3 if u in ("ns", "nsecond") else v * (1e3 if u in ("ms", "msecond") else 1.0)
            d[r[ki].split("(")[0].split("::")[-1]].append(v)
    return d
data = {n: load(n) for n in names}
kernels = sorted(data["cur"])
print(f"{'kernel':26s}" + "".join(f"{n:>12s}" for n in names) + "   (median us over launches)")
for k in kernels:
    print(f"{k:26s}" + "".join(f"{statistics.median(data[n][k]):12.1f}" if data[n].get(k) else f"{'-':>12s}" for n in names))
PY
