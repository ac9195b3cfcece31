"""Generates the golden fixtures of tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

Two kinds of vectors:

  cv_*.npz     outputs of the real OpenCV (cv2) primitives the reference calls on this path (cv::remap,
               cv::pyrDown, cv::sepFilter2D, cv::invertAffineTransform, cv::solve(DECOMP_CHOLESKY), cv::erode).
               They pin oracle/cvmodel.py (the scalar restatement the CUDA kernels are written from) and the
               CUDA kernels themselves to the third-party library the reference links, independent of the
               cv2 build installed where the tests run.

  stack_*.npz  inputs + outputs of the oracle pipeline (oracle/pipeline.py) on small seeded sequences: per-frame
               warp parameters, rho, iteration counts, final average and mask.  The reference has no tests or
               golden vectors for this path and cannot be compiled here (DESIGN.md "Oracle"), so these are
               regression vectors of the restatement, not outputs of the reference binary: PARITY UNPINNED.

All inputs are stored in the files (no dependence on a random generator or on synth.py at test time).
"""
import os
import sys

import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ecc as oecc                 # noqa: E402
from oracle import pipeline as opl             # noqa: E402
from oracle import transforms as otf           # noqa: E402
from oracle import weights as ow               # noqa: E402
from serstacker_b200 import synth              # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def cv_vectors():
    rng = np.random.default_rng(20261017)
    d = {}
    # --- cv::remap (bilinear / bicubic, REPLICATE / REFLECT101 / CONSTANT) + the all-255 mask validity rule
    src = rng.random((45, 61)).astype(f32)
    M = np.array([[1.013, 0.021, -3.37], [-0.017, 0.991, 2.81]], f32)
    ys, xs = np.mgrid[0:45, 0:61].astype(f32)
    mapx = (M[0, 0] * xs + M[0, 1] * ys + M[0, 2]).astype(f32)
    mapy = (M[1, 0] * xs + M[1, 1] * ys + M[1, 2]).astype(f32)
    d["remap_src"], d["remap_mapx"], d["remap_mapy"] = src, mapx, mapy
    for iname, interp in (("linear", cv2.INTER_LINEAR), ("cubic", cv2.INTER_CUBIC)):
        for bname, border in (("replicate", cv2.BORDER_REPLICATE), ("reflect101", cv2.BORDER_REFLECT101),
                              ("constant", cv2.BORDER_CONSTANT)):
            d["remap_%s_%s" % (iname, bname)] = cv2.remap(src, mapx, mapy, interp, borderMode=border, borderValue=0)
        m255 = np.full(src.shape, 255, np.uint8)
        d["mask255_%s" % iname] = cv2.remap(m255, mapx, mapy, interp, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    d["erode5_src"] = ((rng.random((33, 47)) > 0.08) * 255).astype(np.uint8)
    d["erode5"] = cv2.erode(d["erode5_src"], np.ones((5, 5), np.uint8), borderType=cv2.BORDER_CONSTANT,
                            borderValue=255)
    # --- cv::pyrDown on odd / even / narrow sizes
    for i, (h, w) in enumerate(((64, 96), (45, 61), (7, 9), (33, 130))):
        a = rng.random((h, w)).astype(f32)
        d["pyr_src%d" % i] = a
        d["pyr_dst%d" % i] = cv2.pyrDown(a)   # BORDER_DEFAULT (REFLECT101), default size
        # as c_ecch::downscale_image calls it (ecc2.cc:972-982): dstsize from compute_next_pyramid_layer_size (ecc2.h:290-293)
        d["pyr_ecc%d" % i] = cv2.pyrDown(a, dstsize=(((w + 1) >> 1) & ~1, ((h + 1) >> 1) & ~1))
    # --- cv::sepFilter2D with the ECC gradient kernels (ecc2.cc:148-149), widths % 4 == 0
    a = rng.random((40, 64)).astype(f32)
    kd = np.array([1 / 12, -2 / 3, 0, 2 / 3, -1 / 12], f32)
    ks = np.array([0.25, 0.5, 0.25], f32)
    d["sep_src"] = a
    d["sep_gx"] = cv2.sepFilter2D(a, cv2.CV_32F, kd, ks, borderType=cv2.BORDER_REPLICATE)
    d["sep_gy"] = cv2.sepFilter2D(a, cv2.CV_32F, ks, kd, borderType=cv2.BORDER_REPLICATE)
    # --- small linear algebra
    A = rng.random((16, 2, 3)).astype(f32)
    A[:, 0, 0] += 1
    A[:, 1, 1] += 1
    d["inva_src"] = A
    d["inva_dst"] = np.stack([cv2.invertAffineTransform(a) for a in A])
    Hs, vs, xs_ = [], [], []
    for n in (2, 3, 4, 6, 8):
        B = rng.random((n + 3, n)).astype(f32)
        H = (B.T @ B).astype(f32)
        v = rng.random((n, 1)).astype(f32)
        ok, x = cv2.solve(H, v, flags=cv2.DECOMP_CHOLESKY)
        assert ok
        Hp = np.zeros((8, 8), f32)
        Hp[:n, :n] = H
        vp = np.zeros(8, f32)
        vp[:n] = v[:, 0]
        xp = np.zeros(8, f32)
        xp[:n] = x[:, 0]
        Hs.append(Hp), vs.append(vp), xs_.append(xp)
    d["chol_n"] = np.array([2, 3, 4, 6, 8], np.int32)
    d["chol_H"], d["chol_v"], d["chol_x"] = np.stack(Hs), np.stack(vs), np.stack(xs_)
    np.savez_compressed(os.path.join(OUT, "cv_primitives.npz"), **d)
    print("cv_primitives.npz: %d arrays" % len(d))


STACK_CASES = [
    # name, motion, method, interpolation, accumulation, size (w, h), nframes, ecch_max_level
    ("trans_iclm_linear_avg", otf.IMAGE_MOTION_TRANSLATION, oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM,
     cv2.INTER_LINEAR, opl.ACC_AVERAGE, (96, 64), 5, 0),
    ("affine_iclm_cubic_wavg", otf.IMAGE_MOTION_AFFINE, oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM,
     cv2.INTER_CUBIC, opl.ACC_WEIGHTED_AVERAGE, (160, 120), 5, -1),
    ("affine_ic_linear_wavg", otf.IMAGE_MOTION_AFFINE, oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL,
     cv2.INTER_LINEAR, opl.ACC_WEIGHTED_AVERAGE, (160, 120), 4, -1),
    ("affine_lm_cubic_avg", otf.IMAGE_MOTION_AFFINE, oecc.ECC_ALIGN_LM,
     cv2.INTER_CUBIC, opl.ACC_AVERAGE, (160, 120), 4, -1),
    ("trans_fa_linear_avg", otf.IMAGE_MOTION_TRANSLATION, oecc.ECC_ALIGN_FORWARD_ADDITIVE,
     cv2.INTER_LINEAR, opl.ACC_AVERAGE, (128, 96), 4, -1),
]


def stack_vectors():
    for k, (name, motion, method, interp, accm, (w, h), n, maxlvl) in enumerate(STACK_CASES):
        frames, _, _ = synth.make_planet_sequence(w, h, n, seed=100 + k, radius=min(w, h) * 0.3, sigma_t=1.5,
                                                  sigma_rot_deg=0.15, sigma_scale=0.002, blur_range=(0.7, 1.8),
                                                  dtype="u16")
        frames = [np.ascontiguousarray(f) for f in frames]
        so = opl.StackingOptions(accumulation_method=accm)
        so.registration.motion_type = motion
        so.registration.interpolation = interp
        so.registration.ecc.ecc_method = method
        so.registration.ecc.ecch_max_level = maxlvl
        rec = []
        ff = [opl.to_float_frame(f, 16) for f in frames]
        avg, mask, _, _ = opl.run_stacking(ff, so, collect=rec)
        _, w1 = ow.compute_local_variance_map(ff[1], so.sm_dscale, so.sm_kradius, so.sm_uscale)
        d = dict(frames=np.stack(frames), bpp=np.int32(16), motion=np.int32(motion), method=np.int32(method),
                 interpolation=np.int32(interp), weighted=np.int32(accm == opl.ACC_WEIGHTED_AVERAGE),
                 ecch_max_level=np.int32(maxlvl),
                 params=np.stack([np.pad(np.asarray(r["params"], f32), (0, 8 - len(r["params"]))) for r in rec]),
                 nparams=np.int32(len(rec[0]["params"])),
                 ok=np.array([r["ok"] for r in rec], np.int32),
                 rho=np.array([r.get("rho", 0.0) for r in rec], np.float64),
                 avg=avg.astype(f32), mask=mask.astype(np.uint8), w1_frame1=w1.astype(f32))
        np.savez_compressed(os.path.join(OUT, "stack_%s.npz" % name), **d)
        print("stack_%s.npz: %d frames ok=%s" % (name, n, d["ok"].tolist()))


if __name__ == "__main__":
    cv_vectors()
    stack_vectors()
