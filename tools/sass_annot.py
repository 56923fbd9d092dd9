"""Annotated listing of a kernel: ncu SASS page (per-instruction counts) joined with nvdisasm -g line info.
   python tools/sass_annot.py ncu_sass.csv nvdisasm.txt 'mangled substring' warps_per_launch out.txt"""
import csv, re, sys, collections
ncu, dis, kern, W, out = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
rows = list(csv.reader(open(ncu))); hdr = rows[1]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
insts = [(r[iS].strip(), int(r[iE] or 0), int(r[iW] or 0)) for r in rows[2:] if len(r) > iE]
lines, cur, on = [], None, False
for l in open(dis):
    if l.startswith(".text."):
        on = kern in l; continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l): lines.append(cur)
assert len(lines) == len(insts), (len(lines), len(insts))
tot = 0
with open(out, "w") as f:
    for k, (s, e, w) in enumerate(insts):
        f.write(f"{k:5d} {e / W:6.3f} {w:5d} {lines[k][0][:12]}:{lines[k][1]:<5d} {s}\n"); tot += e
print("instructions per warp per tile", tot / W)
