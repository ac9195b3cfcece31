"""Debug helper (not a test): per-trial trace of the GPU solver next to the oracle's, for one configuration."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, ".")
from oracle import ecc as oecc, transforms as otf
from serstacker_b200 import synth, api, capi

motion, method, maxlevel = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
frames, _, _ = synth.make_planet_sequence(320, 240, 4, 11 + motion, sigma_t=2.0, sigma_rot_deg=0.2 if motion else 0.0,
                                          sigma_scale=0.002 if motion in (2, 3, 4) else 0.0, dtype="f32")
kw = dict(maxlevel=maxlevel, minimum_image_size=16, epsx=0.05, max_iterations=30, update_step_scale=1.0)
ot = otf.create_image_transform(motion)
o = oecc.EccH(ot, method=method, **kw)
o.set_reference_image(frames[0], None)
gt = api.create_image_transform(motion)
g = api.c_ecch(gt, method=method, **kw)
g.set_reference_image(frames[0])
capi.check(capi.lib.ssk_ecch_set_trace(g._h, 256))
np.set_printoptions(precision=6, suppress=True, linewidth=200)
for f in frames[1:3]:
    ot.reset(); gt.set_parameters(ot.parameters())
    o.trace = []
    o.align(f, None)
    g.align(f)
    buf = np.zeros((256, 40), np.float32); n = C.c_int()
    capi.check(capi.lib.ssk_ecch_get_trace(g._h, buf.ctypes.data_as(C.POINTER(C.c_float)), 256, C.byref(n)))
    print("=== frame: oracle its", o.num_iterations, "gpu its", g.num_iterations())
    recs = []
    for lvl, tr in o.trace:
        for r in tr:
            recs.append((lvl, r))
    for i in range(max(n.value, len(recs))):
        if i < len(recs):
            lvl, r = recs[i]
            print("O lvl", lvl, "err %.8g" % r.get("err", 0), "newerr %.8g" % r.get("newerr", 0), "lam %g" % r.get("lam", 0), "eps %.6g" % r.get("eps", 0),
                  "cma", r.get("cma"), "newp", r.get("newp"), "p", r["p"], "dp", r["dp"].ravel(), "v", r.get("v", r.get("ep")).ravel())
        if i < n.value:
            b = buf[i]
            print("G lvl", int(b[0]), "err %.8g" % b[3], "newerr %.8g" % b[4], "lam %g" % b[5], "eps %.6g" % b[6], "n", b[7],
                  "p", b[8:16], "tq", b[16:24], "dp", b[24:32], "v", b[32:40])
    print("final O", ot.parameters(), "G", gt.parameters())
# Hp comparison (IC-LM)
if method == 3:
    import cv2
    e0 = o.pyramid[0]
    f = frames[1]
    ot.reset(); gt.set_parameters(ot.parameters())
    g.align(f)
    buf = np.zeros((256, 40), np.float32); n = C.c_int()
    capi.check(capi.lib.ssk_ecch_get_trace(g._h, buf.ctypes.data_as(C.POINTER(C.c_float)), 256, C.byref(n)))
    for i in range(n.value):
        if buf[i, 1] == 9:
            M = int(buf[i, 2]); Hg = buf[i, 4:4 + M * M].reshape(M, M).copy()
            print("oracle Hp\n", e0._Hp)
            print("rel diff\n", (Hg - e0._Hp) / np.abs(e0._Hp))
            ok, x = cv2.solve(e0._Hp, np.ones((M, 1), np.float32), flags=cv2.DECOMP_CHOLESKY)
            ok, y = cv2.solve(Hg, np.ones((M, 1), np.float32), flags=cv2.DECOMP_CHOLESKY)
            print("solve(ones) oracleHp", x.ravel(), "gpuHp", y.ravel())
            print("cond", np.linalg.cond(e0._Hp.astype(np.float64)))
            break
