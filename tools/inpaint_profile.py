"""Runs average_pyramid_inpaint on a 1920x1080 mono image with ~20 % holes a few times (for ncu / timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from serstacker_b200 import api

rng = np.random.default_rng(0)
src = rng.random((1080, 1920), dtype=np.float32)
mask = (rng.random((1080, 1920)) < 0.8).astype(np.uint8) * 255
mask[300:600, 500:1100] = 0
for i in range(4):
    t0 = time.perf_counter()
    out, om = api.average_pyramid_inpaint(src, mask)
    print("host call %.3f ms (includes 8.3 MB H2D + D2H of pageable memory)" % ((time.perf_counter() - t0) * 1e3), om.min())
