"""Static look at one kernel of a cubin / object: python tools/sass_static.py file.o 'ILb1ELb0ELb0ELb0EE' [--upto-atom]
Prints the instruction count and opcode histogram of the kernel whose mangled name contains the pattern; with --upto-atom only
the instructions laid out before the first RED/ATOM (the IEEE re-pass of the fused kernel starts there), which is a fair static
proxy of the FAST pass because its code is straight-line."""
import collections, re, subprocess, sys, tempfile, os
obj, pat = sys.argv[1], sys.argv[2]
upto = "--upto-atom" in sys.argv
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    obj = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
dis = subprocess.run(["nvdisasm", "-c", obj], capture_output=True, text=True).stdout
on, hist, n = False, collections.Counter(), 0
for l in dis.splitlines():
    if l.startswith(".text."):
        on = pat in l
        continue
    if not on: continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if not m: continue
    t = m.group(1).split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    if upto and op in ("RED", "ATOM", "ATOMG", "REDG"): break
    hist[op] += 1; n += 1
fp = sum(hist[k] for k in ("DFMA", "DMUL", "DADD"))
print(f"{n} instructions, {fp} FP64 (DFMA/DMUL/DADD), {n - fp} other")
print("  ".join(f"{k} {v}" for k, v in hist.most_common(24)))
