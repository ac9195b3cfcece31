"""Loops of one kernel in a cuobjdump -sass listing: python tools/sass_loops.py obj.o kernel-substring [min_len]
Prints, for every backward branch whose body holds >= min_len instructions, the body length and its opcode histogram."""
import collections, re, subprocess, sys
obj, key = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 150
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("\n")
ins = []
on = False
for line in txt:
    if "Function :" in line:
        on = key in line
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
print("instructions:", len(ins))
for i, (a, s) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?(?:\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)", s)
    if not m:
        continue
    t = int(m.group(1), 16)
    if t in addr and addr[t] < i and i - addr[t] >= minlen:
        body = ins[addr[t]:i + 1]
        c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in body)
        print("loop %#x..%#x: %d instr: %s" % (t, a, len(body), ", ".join("%s %d" % kv for kv in c.most_common(18))))
