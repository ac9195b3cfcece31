#!/usr/bin/env python
"""Compiles one .cu of serstacker_b200/csrc with -Xptxas -v and prints one line per kernel: registers, stack, spills, shared memory.
Usage: python tools/ptxas_summary.py ssk_fused_tma.cu [filter]"""
import re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "serstacker_b200", "csrc", sys.argv[1])
flt = sys.argv[2] if len(sys.argv) > 2 else ""
extra = os.environ.get("SSK_NVCC_EXTRA", "").split()
r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xptxas", "-v"] + extra +
                   ["-c", src, "-o", "/tmp/_ptxas_summary.o"], capture_output=True, text=True)
txt = r.stderr
if r.returncode:
    print(txt); sys.exit(1)
cur = None
for line in txt.split("\n"):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::|ssk::|void ", "", cur)
        cur = re.sub(r"\(.*$", "", cur)
        stack = None
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur and stack is None:
        stack = m.groups()
    m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, \d+ bytes cumulative stack size)?(?:, (\d+) bytes smem)?", line)
    if m and cur:
        if flt in cur:
            print("%-60s regs %3s  smem %6s  stack %4s  spill st/ld %s/%s" % (cur, m.group(1), m.group(3) or "0", stack[0], stack[1], stack[2]))
        cur = None
