"""GPU parity: c_canvas_average (c_frame_accumulation.cc:264-445) through the C ABI against oracle/accumulation.py::CanvasAverage
(cv2.remap + the reference's running-mean update).  Bilinear / nearest samples are bit-exact against cv::remap, bicubic within
2 ulp; the tolerance on the running mean is 2e-6 relative, the masks must be identical."""
import numpy as np
import cv2
import pytest

from oracle.accumulation import CanvasAverage

pytestmark = pytest.mark.gpu
f32 = np.float32


def _frames(h, w, cn, n, seed):
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur(rng.random((h + 64, w + 64, cn)).astype(f32), (0, 0), 1.5)
    base = base.reshape(h + 64, w + 64, cn)
    out = []
    for i in range(n):
        dx, dy = rng.integers(-6, 7, 2)
        f = base[32 + dy:32 + dy + h, 32 + dx:32 + dx + w] + rng.standard_normal((h, w, cn)).astype(f32) * f32(0.01)
        out.append(np.ascontiguousarray(f[..., 0] if cn == 1 else f, dtype=f32))
    return out


@pytest.mark.parametrize("cn", [1, 3])
@pytest.mark.parametrize("wmode", ["none", "mask", "weights"])
@pytest.mark.parametrize("interp", [cv2.INTER_LINEAR, cv2.INTER_NEAREST, cv2.INTER_CUBIC])
def test_canvas_average_matches_oracle(gpu, cn, wmode, interp):
    from serstacker_b200 import api
    h, w, n = 96, 128, 9
    frames = _frames(h, w, cn, n, seed=cn * 10 + interp)
    rng = np.random.default_rng(3)
    o, g = CanvasAverage(interpolation=interp, canvas_size=(400, 340)), api.c_canvas_average(interpolation=interp, canvas_size=(400, 340))
    yy, xx = np.mgrid[0:h, 0:w].astype(f32)
    walk = [(0, 0)]
    for i in range(1, n):
        walk.append((walk[-1][0] + int(rng.integers(8, 22)), walk[-1][1] - int(rng.integers(5, 16))))   # drifts towards a corner
    for i, f in enumerate(frames):
        wts = None
        if wmode == "mask":
            wts = ((rng.random((h, w)) > 0.2) * 255).astype(np.uint8)
        elif wmode == "weights":
            wts = (rng.random((h, w)) - 0.15).astype(f32)
        if i == 0:
            assert o.add(f, wts) and g.add(f, wts)
        elif i == 4:
            assert o.add(f, wts) and g.add(f, wts)                  # no remap requested: lands on the last box
        else:
            a = 0.003 * i
            rmap = np.stack([xx * f32(np.cos(a)) - yy * f32(np.sin(a)) + f32(0.37 * i), xx * f32(np.sin(a)) + yy * f32(np.cos(a)) - f32(0.21 * i)], -1).astype(f32)
            x0, y0 = o.last_bbox[0] + walk[i][0] - walk[i - 1][0], o.last_bbox[1] + walk[i][1] - walk[i - 1][1]
            bbox = (x0, y0, w, h)
            ro, rg = o.add(f, wts, rmap, bbox), g.add(f, wts, rmap, bbox)
            assert ro == rg
        assert tuple(o.last_bbox) == g.last_bbox(), (i, o.last_bbox, g.last_bbox())
    assert g.accumulated_frames() == o.accumulated_frames
    assert g.accumulator_size()[:2] == (o.accumulator.shape[1], o.accumulator.shape[0])
    ao, mo = o.compute()
    ag, mg = g.compute()
    assert np.array_equal(mo, mg)
    tol = 2e-6 if interp != cv2.INTER_CUBIC else 2e-5
    assert np.abs(ag - ao).max() <= tol * max(1.0, float(np.abs(ao).max()))
    box = (o.last_bbox[0] - 10, o.last_bbox[1] - 7, 70, 50)
    ao2, mo2 = o.compute(box)
    ag2, mg2 = g.compute(rbbox=box)
    assert ao2.shape == ag2.shape and np.array_equal(mo2, mg2) and np.abs(ag2 - ao2).max() <= tol * max(1.0, float(np.abs(ao).max()))


def test_canvas_average_shifts_near_the_edges_and_rejects(gpu):
    """maintainCanvasBoundaries moves the content by 64 px when a box comes within 32 px of an edge; setCanvasSize; clear."""
    from serstacker_b200 import api
    h, w = 160, 200
    frames = _frames(h, w, 1, 6, seed=5)
    o, g = CanvasAverage(canvas_size=(420, 330)), api.c_canvas_average(canvas_size=(420, 330))
    yy, xx = np.mgrid[0:h, 0:w].astype(f32)
    ident = np.stack([xx, yy], -1)
    assert o.add(frames[0]) and g.add(frames[0])
    assert g.accumulator_size() == (420, 330, 1)
    for i, f in enumerate(frames[1:]):
        bbox = (o.last_bbox[0] - 45, o.last_bbox[1] + 38, w, h)
        assert o.add(f, None, ident + f32(0.25), bbox) == g.add(f, None, ident + f32(0.25), bbox)
        assert tuple(o.last_bbox) == g.last_bbox()
    ao, mo = o.compute()
    ag, mg = g.compute()
    assert np.array_equal(mo, mg) and np.abs(ag - ao).max() <= 2e-6
    assert g.add(frames[0], None, ident, (10000, 10000, w, h)) is False      # ROI is empty
    with pytest.raises(api.SskError):
        g.add(frames[0].astype(np.uint16))                                   # CV_32F frames only
    g.clear()
    assert g.accumulated_frames() == 0 and g.accumulator_size()[0] == 0
