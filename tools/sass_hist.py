"""Opcode histogram of one kernel from an ncu source page: ncu -i rep --page source --csv --print-source sass > x.csv; python tools/sass_hist.py x.csv [cells]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
cells = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
iS, iE, iT, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
hist, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iT: continue
    src = r[iS].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    n = int(r[iE] or 0)
    hist[op] += n; tot += n
    samp[op] += int(r[iSamp] or 0)
st = sum(samp.values())
print(f"total warp-instructions {tot}" + (f" = {tot*32/cells:.0f} thread-instr per cell-update (32-lane)" if cells else ""))
for op, n in hist.most_common(40):
    print(f"{op:12s} {n:12d} {100*n/tot:6.2f}%   stall samples {100*samp[op]/st:5.1f}%" + (f"   {n*32/cells:7.1f}/cell" if cells else ""))
