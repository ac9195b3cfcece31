"""Builds serstacker_b200/libssk.so (CUDA kernels + C ABI) for sm_100a with nvcc, in-tree.

Usage: python -m serstacker_b200.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# SSK_BUILD_TAG=<tag> (with SSK_NVCC_EXTRA=-D...): a side build of the same ABI into libssk_<tag>.so for A/B runs (SSK_LIB selects it)
TAG = os.environ.get("SSK_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))
LIB = os.path.join(HERE, "libssk" + ("_" + TAG if TAG else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17"] + os.environ.get("SSK_NVCC_EXTRA", "").split() + [
         "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    hs.append(os.path.join(HERE, "..", "include", "ssk.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    srcp = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(srcp)
            and os.path.getmtime(obj) > _headers_mtime()):
        return obj, False
    cmd = [NVCC] + FLAGS + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, True


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in res]
    rebuilt = any(ch for _, ch in res)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]   # NCCL is dlopen'ed at run time (ssk_multi.cu)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("libssk.so:", "rebuilt" if rebuilt else "up to date", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
