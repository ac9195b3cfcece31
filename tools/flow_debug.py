"""Debug aid: where does the device eccflow leave the oracle?  (run on the GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import eccflow as oef
from serstacker_b200 import api
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_eccflow import _scene, _gpu_options

ref, cur, ident = _scene(270, 480, 1)
for kw in [dict(max_pyramid_level=0, max_iterations=1), dict(max_pyramid_level=0, max_iterations=3), dict(max_pyramid_level=1, max_iterations=1),
           dict(max_pyramid_level=3, max_iterations=3), dict(max_pyramid_level=8, max_iterations=3), dict()]:
    o = oef.registration_options(**kw)
    fo = oef.EccFlow(o); fo.set_reference_image(ref)
    fg = api.c_eccflow(_gpu_options(o)); fg.set_reference_image(ref)
    if not kw.get("max_pyramid_level"):
        for l, e in enumerate(fo.pyramid):
            D = fg.pyramid_image(4, l)
            rel = np.abs(D - e.D) / np.maximum(np.abs(e.D), 1e-30)
            print("  level", l, "D rel diff max per channel", rel.reshape(-1, 4).max(0))
    for init in (None, ident):
        want = fo.compute(cur, init); got = fg.compute(cur, init)
        d = np.abs(got - want).max(-1)
        print(kw, "init", "empty" if init is None else "ident", "levels", len(fo.pyramid), "max %.3g p999 %.3g mean %.3g  |uv| max %.3g" % (d.max(), np.quantile(d, .999), d.mean(), np.abs(fo.uv).max()))
