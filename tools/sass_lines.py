"""Join an ncu SASS source page (per-instruction execution counts) with nvdisasm -g line info of the same kernel:
   python tools/sass_lines.py ncu_sass.csv nvdisasm_all.txt '<mangled kernel name>' source.cu [cells]
Prints the source lines with the most executed warp-instructions, split FP64 / other."""
import csv, re, sys, collections
ncu, dis, kern, src = sys.argv[1:5]
cells = float(sys.argv[5]) if len(sys.argv) > 5 else None
rows = list(csv.reader(open(ncu)))
hdr = rows[1]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
insts = [(r[iS].strip(), int(r[iE] or 0), int(r[iW] or 0)) for r in rows[2:] if len(r) > iE]
lines = []
cur, on = None, False
for l in open(dis):
    if l.startswith(".text."):
        on = l.strip() == f".text.{kern}:"
        continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m: lines.append(cur)
print(f"ncu instructions {len(insts)}, nvdisasm instructions {len(lines)}")
n = min(len(insts), len(lines))
tot = collections.Counter(); fp = collections.Counter(); st = collections.Counter()
for k in range(n):
    s, e, w = insts[k]
    t = s.split()
    op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "")).split(".")[0]
    tot[lines[k]] += e; st[lines[k]] += w
    if op in ("DFMA", "DMUL", "DADD", "DSETP"): fp[lines[k]] += e
text = open(src).read().splitlines()
allx = sum(tot.values()); alls = sum(st.values())
sc = 32 / cells if cells else 1
print("   line    total  fp64  other  stall%   source")
for (f, ln), e in tot.most_common(60):
    t = text[ln - 1].strip()[:110] if f == src.split("/")[-1] and ln <= len(text) else f
    print(f"{f[:14]:14s}:{ln:5d} {e*sc:7.1f} {fp[(f,ln)]*sc:6.1f} {(e-fp[(f,ln)])*sc:6.1f} {100*st[(f,ln)]/alls:6.1f}   {t}")
# region summary for csi_fused.cu (line ranges of the current source; pass REGIONS="name:lo-hi,..." to override)
import os
reg = os.environ.get("REGIONS")
if reg:
    print("\nregion            total   fp64  other  stall%")
    for item in reg.split(","):
        name, rng = item.split(":"); lo, hi = map(int, rng.split("-"))
        keys = [k for k in tot if k and k[0] == src.split("/")[-1] and lo <= k[1] <= hi]
        e = sum(tot[k] for k in keys); f_ = sum(fp[k] for k in keys); w = sum(st[k] for k in keys)
        print(f"{name:16s} {e*sc:7.1f} {f_*sc:6.1f} {(e-f_)*sc:6.1f} {100*w/alls:6.1f}")
    keys = [k for k in tot if not k or k[0] != src.split("/")[-1]]
    e = sum(tot[k] for k in keys); f_ = sum(fp[k] for k in keys); w = sum(st[k] for k in keys)
    print(f"{'other files':16s} {e*sc:7.1f} {f_*sc:6.1f} {(e-f_)*sc:6.1f} {100*w/alls:6.1f}")
