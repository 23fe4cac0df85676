"""cuobjdump -sass of the shipped library -> per-kernel counts of the Blackwell-native mnemonics (profiles/r02_sass_summary.md).
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor loads, UTCBAR = tcgen05.commit, HMMA = legacy mma.sync."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "carefree-learn_b200", "libb200_cflearn.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), stdout=subprocess.PIPE, text=True).stdout.split("\n")
names = dict(zip(re.findall(r"Function : (\S+)", out), demangle))
pats = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "USETMAXREG", "HMMA", "FFMA2", "MUFU"]
rows = []
cur, counts = None, None
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, counts))
        cur, counts = m.group(1), collections.Counter()
        continue
    if cur:
        for p in pats:
            if re.search(r"\b" + p + r"\b|\b" + p + r"\.", line):
                counts[p] += 1
        if re.search(r"^\s+/\*[0-9a-f]{4,6}\*/\s+\S", line):
            counts["instr"] += 1
if cur:
    rows.append((cur, counts))
print("# SASS summary of libb200_cflearn.so (sm_100a) -- `python tools/sass_summary.py`\n")
print("| kernel | instr | " + " | ".join(pats) + " |")
print("|---|---:|" + "---:|" * len(pats))
tot = collections.Counter()
for fn, c in sorted(rows, key=lambda r: -r[1]["UTCHMMA"]):
    nm = names.get(fn, fn).replace("b200::", "").replace("void ", "")
    nm = (nm.split(">(")[0] + ">") if ">(" in nm else re.sub(r"\(.*", "", nm)
    nm = nm.replace("(int)", "").replace("(bool)", "")
    print(f"| `{nm}` | {c['instr']} | " + " | ".join(str(c[p]) for p in pats) + " |")
    tot.update(c)
print(f"| **total** | {tot['instr']} | " + " | ".join(str(tot[p]) for p in pats) + " |")
