"""Per-kernel SASS census of libb200w.so: how many tcgen05 MMA (UTCHMMA / UTCQMMA), TMA load / store
(UTMALDG / UTMASTG / UTMAREDG), TMEM load / store (LDTM / STTM), legacy tensor-core (HMMA) and
programmatic-dependent-launch instructions each kernel holds. Run here (no GPU):
    python tools/sass_census.py > profiles/r02_sass_census.txt
tests/test_abi.py asserts the same facts on the whole library."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "runbooks_b200", "libb200w.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "ACQBULK",
             "PREEXIT", "SYNCS", "MUFU.EX2", "ATOM", "RED", "STL", "LDL"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                          text=True).stdout.splitlines()
names = re.findall(r"Function : (\S+)", sass)
pretty = dict(zip(names, demangle)) if len(names) == len(demangle) else {}
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and "/*" in line:
        for mn in MNEMONICS:
            if re.search(r"(?<![A-Z])" + re.escape(mn) + r"\b", line):
                counts[cur][mn] += 1
        counts[cur]["_total"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
print(f"SASS census of {os.path.relpath(LIB, ROOT)} (sm_100a), {len(counts)} kernels; columns = instruction counts")
print("  UTCHMMA = tcgen05.mma (kind::f16)   UTMALDG / UTMASTG = cp.async.bulk.tensor load / store   LDTM / STTM = tcgen05.ld / st")
print("  HMMA = legacy mma.sync (must be 0)   STL / LDL = local-memory spills\n")
hdr = f"{'kernel':78s} {'instr':>6s} " + " ".join(f"{m[:8]:>8s}" for m in MNEMONICS)
print(hdr)
tot = collections.Counter()
for k, c in counts.items():
    name = re.sub(r"b200w::|\(anonymous namespace\)::|<unnamed>::", "", pretty.get(k, k)).replace("void ", "")
    depth, cut = 0, len(name)
    for i, ch in enumerate(name):            # drop the argument list, keep the template arguments
        depth += ch == "<"
        depth -= ch == ">"
        if ch == "(" and depth == 0:
            cut = i
            break
    name = re.sub(r"\((int|bool)\)", "", name[:cut])
    print(f"{name[:78]:78s} {c['_total']:6d} " + " ".join(f"{c[m]:8d}" for m in MNEMONICS))
    tot.update(c)
print(f"{'TOTAL':78s} {tot['_total']:6d} " + " ".join(f"{tot[m]:8d}" for m in MNEMONICS))
