"""CPU: the C-ABI library loads and exports every symbol include/ssk.h declares; entry points fail loudly
(no CPU fallback) when there is no GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ssk.h")).read()
    return sorted(set(re.findall(r"SSK_API\s+[\w\s\*]+?\b(ssk_\w+)\s*\(", src)))


def test_library_builds_and_loads():
    from serstacker_b200 import build
    build.build(verbose=False)
    from serstacker_b200 import capi
    assert capi.lib.ssk_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    from serstacker_b200 import capi
    names = _declared()
    assert len(names) > 40
    lib = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libssk.so does not export %s" % n
    assert set(names) == set(capi.EXPORTED), set(names) ^ set(capi.EXPORTED)


def test_only_declared_symbols_are_visible():
    from serstacker_b200 import capi
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    extra = {s for s in exported if s.startswith("ssk_")} - set(_declared())
    assert not extra, extra


def test_no_oracle_import_in_product_code():
    pkg = os.path.join(ROOT, "serstacker_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), os.path.join(dp, f)
                calls_cv = re.search(r"^\s*(import cv2|from cv2)", txt, re.M) or "#include <opencv" in txt
                # synth.py only generates synthetic inputs; the host adapter wraps cv::Mat when OpenCV headers exist
                assert not calls_cv or f in ("synth.py", "ssk_adapter.h"), "product code must not call OpenCV: %s" % f


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from serstacker_b200 import api, capi
    with pytest.raises(capi.SskError) as e:
        api.c_weigthed_average()
    assert e.value.code == capi.SSK_ERR_CUDA
    with pytest.raises(capi.SskError):
        api.c_frame_registration(api.registration_options())
    with pytest.raises(capi.SskError):
        api.compute_local_variance_map(np.zeros((16, 16), np.float32))
