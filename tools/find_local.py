"""Which source lines keep a specialised kernel's workspace in local memory?  Compiles one generated
spec_<name>.cu to PTX and lists, per kernel, the lines whose local loads/stores are not at a constant
offset of the frame (dynamic indexing) and the lines of loops left rolled ('.pragma "nounroll"').
usage: python tools/find_local.py pendulum5 [kernel-substring]"""
import os, re, subprocess, sys
from collections import Counter

here = os.path.dirname(os.path.abspath(__file__))
csrc = os.path.join(here, "..", "trep_b200", "csrc")
name = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
ptx = "/tmp/find_local_%s.ptx" % name
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-ptx",
                os.path.join(csrc, "gen", "spec_%s.cu" % name), "-o", ptx, "-I", os.path.join(here, "..", "include")],
               check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = open(ptx).read().split("\n")
files = {}
for l in lines:
    m = re.match(r'\s*\.file\s+(\d+)\s+"([^"]+)"', l)
    if m: files[int(m.group(1))] = os.path.basename(m.group(2))
kern, cur, depot = None, None, None
dyn, roll, tot = Counter(), Counter(), Counter()
for l in lines:
    m = re.match(r"\.(?:visible )?\.?entry (\S+)\(", l)
    if m:
        k = re.search(r"trepb\d+([a-z0-9_]*kernel)", m.group(1)); kern = k.group(1) if k else m.group(1)[:30]
    m = re.match(r"\s*\.loc\s+(\d+)\s+(\d+)(.*)", l)
    if m: cur = (files.get(int(m.group(1))), int(m.group(2)), tuple(re.findall(r"inlined_at \d+ (\d+)", m.group(3))))
    if want not in (kern or ""): continue
    if "ld.local" in l or "st.local" in l:
        tot[kern] += 1
        if not re.search(r"\[%rd\d+(\+\d+)?\]", l) or True:
            a = re.search(r"\[(%rd\d+)(\+\d+)?\]", l)
            dyn[(kern, cur, a.group(1) if a else "?")] += 1
    if '"nounroll"' in l: roll[(kern, cur)] += 1
for k, v in tot.items(): print("local accesses", k, v)
bases = Counter()
for (k, c, b), v in dyn.items(): bases[(k, b)] += v
print("by base register:", dict(bases))
main = {k: max((b for (kk, b) in bases if kk == k), key=lambda b: bases[(k, b)]) for k in tot}
for (k, c, b), v in sorted(dyn.items(), key=lambda x: str(x[0])):
    if b != main[k]: print("dynamic", k, c, v)
for k, v in roll.items(): print("rolled loop", k, v)
