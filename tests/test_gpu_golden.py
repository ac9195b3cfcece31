"""GPU parity against the committed golden vectors (tests/golden/, see make_golden.py): the CUDA path through the
C ABI vs outputs of real OpenCV primitives and vs the oracle pipeline outputs stored with their inputs."""
import glob
import os

import numpy as np
import cv2
import pytest

from helpers import map_diff_px, rel_l2

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STACKS = sorted(glob.glob(os.path.join(GOLD, "stack_*.npz")))


@pytest.fixture(scope="module")
def cvp():
    return np.load(os.path.join(GOLD, "cv_primitives.npz"))


@pytest.mark.parametrize("iname,interp", [("linear", cv2.INTER_LINEAR), ("cubic", cv2.INTER_CUBIC)])
@pytest.mark.parametrize("bname,border", [("replicate", cv2.BORDER_REPLICATE), ("reflect101", cv2.BORDER_REFLECT101),
                                          ("constant", cv2.BORDER_CONSTANT)])
def test_remap_matches_opencv_golden(gpu, cvp, iname, interp, bname, border):
    from serstacker_b200 import api
    rmap = np.ascontiguousarray(np.dstack([cvp["remap_mapx"], cvp["remap_mapy"]]))
    got, mask = api.remap(None, rmap, cvp["remap_src"], want_mask=True, interpolation=interp, border_mode=border)
    want = cvp["remap_%s_%s" % (iname, bname)]
    if iname == "linear":
        assert np.array_equal(got, want)
    else:
        assert np.abs(got - want).max() <= 4e-7 * max(1.0, np.abs(want).max())
    # base_remap's mask: erode5x5(remap(all-255, interp, CONSTANT 0) >= 255), border value 255
    m = ((cvp["mask255_%s" % iname] >= 255) * 255).astype(np.uint8)
    m = cv2.erode(m, np.ones((5, 5), np.uint8), borderType=cv2.BORDER_CONSTANT, borderValue=255)
    assert np.array_equal(mask, m)


@pytest.mark.parametrize("i", range(4))
def test_pyramid_matches_opencv_golden(gpu, cvp, i):
    """Level 1 of the ECC reference pyramid (sigma 0, no normalisation) is exactly cv::pyrDown of level 0."""
    from serstacker_b200 import api
    src = cvp["pyr_src%d" % i]
    if min(src.shape) < 16:
        pytest.skip("below the ECC minimum image size")
    g = api.c_ecch(None, maxlevel=2, minimum_image_size=4, reference_smooth_sigma=0.0)
    g.set_reference_image(src)
    assert g.num_levels() == 2
    assert np.array_equal(g.reference_image(1), cvp["pyr_ecc%d" % i])


@pytest.mark.parametrize("path", STACKS, ids=[os.path.basename(p)[6:-4] for p in STACKS])
@pytest.mark.parametrize("max_batch", [1, 3])
def test_stacking_matches_golden(gpu, path, max_batch):
    from serstacker_b200 import api
    g = np.load(path)
    frames = [np.ascontiguousarray(f) for f in g["frames"]]
    motion, n = int(g["motion"]), int(g["nparams"])
    ro = api.registration_options(motion_type=motion, interpolation=int(g["interpolation"]),
                                  ecc=dict(ecc_method=int(g["method"]), ecch_max_level=int(g["ecch_max_level"])))
    p = api.c_image_stacking_pipeline(api.stack_options(registration=ro, accumulation_method=1 if int(g["weighted"]) else 0,
                                                        max_batch=max_batch))
    p.set_reference(frames[0], bpp=int(g["bpp"]))
    res = p.add_frames(frames)
    h, w = frames[0].shape[:2]
    for r, pw, ok in zip(res, g["params"], g["ok"]):
        assert bool(r["ok"]) == bool(ok)
        assert map_diff_px(motion, r["params"][:n], pw[:n], (w, h)) <= 1e-3      # north_star: 1e-3 px
    avg, mask = p.compute()
    m = (mask > 0) & (g["mask"] > 0)
    assert (mask > 0).sum() == (g["mask"] > 0).sum() or np.mean((mask > 0) != (g["mask"] > 0)) < 1e-3
    assert rel_l2(avg, g["avg"], m) <= 1e-4                                      # north_star: 1e-4 rel L2
