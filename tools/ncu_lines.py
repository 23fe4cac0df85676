import csv, re, subprocess, sys, collections
rep, func, cubin = sys.argv[1], sys.argv[2], sys.argv[3]
src = sys.argv[4] if len(sys.argv) > 4 else None
KSTART = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout.split("\n")
# collect instructions of function with line numbers
ins = []; cur = None; inside = False
for l in dis:
    if l.startswith("\t.section") or l.startswith(".section") or "//-----" in l:
        inside = (func in l) if "//-----" in l else inside
        continue
    if not inside: continue
    if "//## File" in l:
        locs = re.findall(r'"([^"]+)", line (\d+)', l)
        # outermost location inside the kernel's own file (last one in the chain that is in the .cu), else innermost
        cu = [int(n) for f, n in locs if f.endswith(".cu")]
        cur = cu[0] if cu and len(cu) == 1 else (cu[-1] if cu else -int(locs[0][1]))
        # prefer the deepest .cu line that is >= kernel start (role code), keep helper lines otherwise
        big = [n for n in cu if n >= KSTART]
        if big: cur = big[0]
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append((cur, m.group(2).strip()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[1]; data = rows[2:]; ix = {k: i for i, k in enumerate(h)}
print("disasm instrs", len(ins), "ncu instrs", len(data))
st = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
mism = 0
agg = collections.defaultdict(lambda: collections.Counter())
tot = 0
for i, r in enumerate(data):
    if i >= len(ins): break
    op_n = r[ix["Source"]].strip().split()[0:2]
    op_d = ins[i][1].split()[0:2]
    if op_n[:1] != op_d[:1] and not (op_n and op_n[0].startswith("@")): mism += 1
    line = ins[i][0]
    n = int(r[ix["# Samples"]]); tot += n
    agg[line]["samples"] += n
    agg[line]["inst"] += int(r[ix["Instructions Executed"]])
    for k in st:
        v = int(r[ix[k]])
        if v: agg[line][k] += v
print("opcode mismatches", mism, "total samples", tot)
srcl = open(src).read().split("\n") if src else None
for line, c in sorted(agg.items(), key=lambda x: -x[1]["samples"])[:int(__import__("os").environ.get("NCU_LINES_TOP", "45"))]:
    top = [(k[6:], v) for k, v in c.most_common(6) if k.startswith("stall_")][:3]
    print(f"{line:5d} {c['samples']:6d} {100*c['samples']/tot:5.1f}% inst {c['inst']:9d} {top}  | {srcl[line-1].strip()[:90] if srcl and line and line > 0 else ''}")
