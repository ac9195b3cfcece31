"""Top stalled SASS instructions of one kernel of an ncu report (source page): python tools/ncu_hot.py rep kernel-regex [N]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# the export may hold several launches of the kernel: keep the first block
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
start = hdr_i[0]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
hdr = rows[start]
body = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
texec = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
print("instructions (SASS lines): %d, warp instructions executed: %d, samples: %d" % (len(body), texec, tot))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[col[h]] or 0) for r in body) for h in stall_cols}
print("samples by reason:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(tot, 1)) for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:N]
for i in sorted(idx):
    r = body[i]
    s = int(r[col["# Samples"]] or 0)
    top = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print("%5d %5.2f%% exec %9s  %-70s %s" % (i, 100.0 * s / max(tot, 1), r[col["Instructions Executed"]], r[col["Source"]].strip()[:70],
                                             " ".join("%s:%d" % (n, v) for v, n in top if v)))
