"""Attribute the per-SASS-instruction counters of an ncu report to source lines.
  python tools/sass_lines.py <report.ncu-rep> <cubin-name e.g. coop_puppet> <kernel-substring> [top N]
Joins `ncu --page source --csv` (per SASS address: instructions executed, stall samples) with
`nvdisasm -g` of the cubin inside trep_b200/libtrepb.so (address -> file:line of the innermost inlined
frame) and prints instructions / samples per source line, per line range (function) and per opcode."""
import csv, os, re, subprocess, sys, tempfile
from collections import Counter, defaultdict

rep, cub, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "..", "trep_b200", "libtrepb.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", cub + ".sm_100a.cubin", os.path.abspath(so)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub + ".sm_100a.cubin")], capture_output=True, text=True).stdout
line_of = {}
cur, infn = None, False
for l in sass.split("\n"):
    if l.startswith(".text."):
        infn = kern in l
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
per_line, per_op, samp_line = Counter(), Counter(), Counter()
stall_line = defaultdict(Counter)
tot = tots = 0
for r in rows[2:]:
    if len(r) <= isamp or not r[ia].startswith("0x"): continue
    a = int(r[ia], 16)
    if base is None: base = a
    off = a - base
    n = int(r[ii] or 0); s = int(r[isamp] or 0)
    loc, txt = line_of.get(off, (None, r[1].strip()))
    per_line[loc] += n; samp_line[loc] += s
    op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
    per_op[op.split(".")[0]] += n
    for c in stall_cols:
        v = int(r[c] or 0)
        if v: stall_line[loc][hdr[c]] += v
    tot += n; tots += s
print("total warp instructions %d, samples %d" % (tot, tots))
print("--- by opcode")
for op, n in per_op.most_common(25): print("  %-12s %6.2f %%" % (op, 100.0 * n / tot))
print("--- by source line (instr %, samples %, top stalls)")
for loc, n in sorted(per_line.items(), key=lambda x: -samp_line[x[0]])[:top]:
    st = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in stall_line[loc].most_common(3))
    print("  %-34s instr %5.2f %%  samples %5.2f %%   %s" % ("%s:%s" % loc if loc else "?", 100.0 * n / tot, 100.0 * samp_line[loc] / max(tots, 1), st))
# ranges of trepb_coop_math.cuh (functions)
if len(sys.argv) > 5:
    ranges = [tuple(x.split(":")) for x in sys.argv[5].split(",")]
    print("--- by line range of trepb_coop_math.cuh")
    for name, lo, hi in ranges:
        lo, hi = int(lo), int(hi)
        n = sum(v for k, v in per_line.items() if k and k[0] == "trepb_coop_math.cuh" and lo <= k[1] <= hi)
        s = sum(v for k, v in samp_line.items() if k and k[0] == "trepb_coop_math.cuh" and lo <= k[1] <= hi)
        print("  %-24s instr %5.2f %%  samples %5.2f %%" % (name, 100.0 * n / tot, 100.0 * s / max(tots, 1)))
    files = Counter(); sf = Counter()
    for k, v in per_line.items(): files[k[0] if k else None] += v
    for k, v in samp_line.items(): sf[k[0] if k else None] += v
    for f, v in files.most_common(): print("  file %-28s instr %5.2f %% samples %5.2f %%" % (f, 100.0 * v / tot, 100.0 * sf[f] / max(tots, 1)))
