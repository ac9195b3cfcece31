"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py x.csv"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
t = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v if r[ui] in ("us", "usecond") else v * 1000.0 if r[ui] in ("ms", "msecond") else v
    name = re.sub(r"\(.*", "", r[ki])[:90]
    n, s = t.get(name, (0, 0.0))
    t[name] = (n + 1, s + v)
tot = sum(s for _, s in t.values())
print("%d launches, %.1f us in kernels" % (sum(n for n, _ in t.values()), tot))
for name, (n, s) in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print("  %6d x %9.2f us avg  %10.1f us  %5.1f %%  %s" % (n, s / n, s, 100 * s / tot, name))
