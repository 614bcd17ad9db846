"""Aggregates an ncu gpu__time_duration launch list (csv) by kernel -> profiles/*.txt"""
import collections, csv, re, sys
src, dst, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
lines = [l for l in open(src) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    ns = v * 1000 if u in ("usecond", "us") else (v if u in ("nsecond", "ns") else v * 1e6)
    k = re.sub(r"\(.*", "", row["Kernel Name"]); k = re.sub(r"^void ", "", k)
    k = k.replace("b200w::(anonymous namespace)::", "").replace("unnamed>::", "")
    agg[k][0] += 1; agg[k][1] += ns
tot = sum(v[1] for v in agg.values())
out = [note, f"total {tot / 1e6:.2f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare SHARES)"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{t / 1e6:9.3f} ms {100 * t / tot:5.1f}% n={n:4d} avg {t / n / 1e3:9.1f} us  {k[:110]}")
open(dst, "w").write("\n".join(out) + "\n"); print("\n".join(out))
