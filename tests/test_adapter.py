"""The C++ adapter (serstacker_b200/host/ssk_adapter.h: the reference's class names over the C ABI) compiles with
g++ alone, links libssk.so, and - on a GPU - registers and stacks an analytic scene through those classes."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "cpp", "_build")


def _build():
    from serstacker_b200 import build as b
    b.build()
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "adapter_smoke")
    libdir = os.path.join(ROOT, "serstacker_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "serstacker_b200", "host"), os.path.join(ROOT, "tests", "cpp", "adapter_smoke.cc"),
           "-o", exe, "-L", libdir, "-lssk", "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_adapter_compiles_and_links():
    exe = _build()
    r = subprocess.run([exe, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_adapter_registers_and_stacks(gpu):
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
