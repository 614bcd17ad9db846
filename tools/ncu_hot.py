"""Top stalled SASS instructions of a kernel from an ncu report (source page)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
iS, iN, iI = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[1:]):
    try: s = int(r[iS])
    except: continue
    data.append((s, k, r))
tot = sum(d[0] for d in data)
print(f"total samples {tot}")
for s, k, r in sorted(data, reverse=True)[:top]:
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    ctx = rows[1 + max(0, k - 1)][iN].strip()[:60]
    print(f"{100 * s / tot:5.1f}% #{k:5d} {r[iN].strip()[:70]:70s} exec={r[iI]:>8s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}   prev: {ctx}")
