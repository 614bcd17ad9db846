"""Per-CUDA-source-line cost of a kernel from an ncu report captured with -lineinfo and --import-source on:
instructions executed (warp level) and warp-stall samples, top N lines. Inlined helpers (ptx.cuh) appear
under their own file.   python tools/ncu_lines.py <report.ncu-rep> [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines, cur_file, hdr = [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        iA, iS, iI = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    elif hdr and len(r) > iI and r[iA] == "-":          # a CUDA line (SASS rows carry an address)
        try:
            lines.append((int(r[iI]), int(r[iS]), cur_file, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot_i, tot_s = sum(l[0] for l in lines), sum(l[1] for l in lines)
print(f"{rep}: {tot_i:,} warp instructions, {tot_s:,} stall samples attributed to {len(lines)} source lines")
print(f"{'instr %':>8s} {'samples %':>9s}  location                       source")
for i, s, f, n, src in sorted(lines, reverse=True)[:top]:
    print(f"{100 * i / tot_i:8.1f} {100 * s / max(1, tot_s):9.1f}  {f + ':' + str(n):30s} {src[:95]}")
