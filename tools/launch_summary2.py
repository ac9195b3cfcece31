"""Like launch_summary.py but grouped by (kernel, grid size): python tools/launch_summary2.py x.csv [kernel-substring]"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
key = sys.argv[2] if len(sys.argv) > 2 else ""
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ki, vi, ui, gi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("Grid Size")
t = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi or key not in r[ki]:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v if r[ui] in ("us", "usecond") else v * 1000.0
    name = re.sub(r"\(.*", "", r[ki])[-40:] + " " + r[gi]
    n, s = t.get(name, (0, 0.0))
    t[name] = (n + 1, s + v)
tot = sum(s for _, s in t.values())
print("%d launches, %.1f us" % (sum(n for n, _ in t.values()), tot))
for name, (n, s) in sorted(t.items(), key=lambda kv: -kv[1][1])[:40]:
    print("  %5d x %9.2f us  %9.1f us %5.1f %%  %s" % (n, s / n, s, 100 * s / tot, name))
