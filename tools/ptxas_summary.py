"""Summarise the ptxas -v logs of trep_b200/csrc/build: kernel, registers, stack frame, spills."""
import glob, os, re, sys

def main(pattern="*"):
    here = os.path.dirname(os.path.abspath(__file__))
    for f in sorted(glob.glob(os.path.join(here, "..", "trep_b200", "csrc", "build", pattern + ".log"))):
        txt = open(f).read()
        cur = None
        rows = []
        for line in txt.split("\n"):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                name = m.group(1)
                k = re.search(r"trepb\d*(?:5coopk)?\d+([a-z0-9_]*kernel)", name)
                cur = [k.group(1) if k else name[:40], None, None, None]
                rows.append(cur)
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur[1] is None:
                cur[1] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m and cur[2] is None:
                cur[2] = int(m.group(1))
        if rows:
            print(os.path.basename(f).replace(".cu.o.log", ""))
            for r in rows:
                print("    %-22s regs %3s  stack %6d  spill st/ld %5d/%5d" % (r[0], r[2], r[1][0], r[1][1], r[1][2]))

if __name__ == "__main__":
    main(*sys.argv[1:])
