"""One line per metric of a bench.py JSON line (headline + secondary)."""
import json, sys
for path in sys.argv[1:]:
    for l in open(path):
        l = l.strip()
        if not l.startswith("{"): continue
        d = json.loads(l)
        print("%s  n_gpus=%s" % (path, d.get("n_gpus")))
        def row(m):
            r = m.get("roofline") or {}
            print("  %-118s %.4g %s  frac %s %s" % (m["metric"][:118], m["value"], m.get("unit", ""), ("%.3f" % r["frac"]) if "frac" in r else "-", r.get("bound", "")))
        row(d)
        if "e2e" in d: print("  e2e %.4g" % d["e2e"]["value"])
        for m in d.get("secondary", []): row(m)
