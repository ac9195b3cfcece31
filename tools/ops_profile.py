"""Runs the operators either side of the per-frame loop once each at their configuration sizes (for an ncu launch list):
debayer_nn2 on a 4096x3000 RGGB16 frame (config #3), unsharp_mask(1, 0.8) and average_pyramid_inpaint on 1920x1080 mono."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from serstacker_b200 import api

rng = np.random.default_rng(0)
raw = rng.integers(0, 65536, (3000, 4096)).astype(np.uint16)
img = rng.random((1080, 1920), dtype=np.float32)
mask = (rng.random((1080, 1920)) < 0.8).astype(np.uint8) * 255
mask[300:600, 500:1100] = 0
for _ in range(3):
    api.debayer_nn2(raw, 8)
    api.unsharp_mask(img, 1.0, 0.8)
    api.average_pyramid_inpaint(img, mask)
print("ok")
