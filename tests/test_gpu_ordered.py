"""GPU: stream-ordered call chains (ssk_set_stream_ordered, include/ssk.h).  Device-resident lpg -> GaussianBlur -> add and
jdr_derotate_and_add chains enqueued without per-call host waits must give the bits of the blocking calls on host arrays
(the reference's operator semantics, c_jdr_pipeline.cc:1184-1236 and the focus-stack loop of BASELINE config #5)."""
import ctypes as C
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(h, w, cn, seed):
    import cv2
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.random((h, w, cn)).astype(np.float32), (0, 0), 1.5 + 0.5 * seed)
    return np.ascontiguousarray(img.reshape(h, w, cn) if cn > 1 else img.reshape(h, w))


def test_focus_chain_stream_ordered_equals_blocking(gpu):
    import torch
    from serstacker_b200 import api, capi
    H, W = 260, 332
    frames = [_scene(H, W, 3, s) for s in range(5)]
    blocking = api.c_weigthed_average()
    for f in frames:
        blocking.add(f, api.gaussian_blur(api.lpg(f, k=6.0, p=2.0, dscale=0, uscale=0), 1.0))
    want, wmask = blocking.compute()

    dev = torch.device("cuda", 0)
    dfr = [torch.from_numpy(f).to(dev) for f in frames]
    wmap = torch.empty((H, W), dtype=torch.float32, device=dev)
    wblur = torch.empty((H, W), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    mw, mb = capi.device_mat(wmap.data_ptr(), H, W, np.float32), capi.device_mat(wblur.data_ptr(), H, W, np.float32)
    acc = api.c_weigthed_average()
    assert api.set_stream_ordered(True) is False
    try:
        for t in dfr:
            m = capi.device_mat(t.data_ptr(), H, W, np.float32, cn=3)
            capi.check(capi.lib.ssk_lpg(C.byref(m), 6.0, 2.0, 0, 0, C.byref(mw)))
            capi.check(capi.lib.ssk_gaussian_blur(C.byref(mw), 1.0, 1.0, C.byref(mb)))
            capi.check(capi.lib.ssk_acc_add(acc._h, C.byref(m), C.byref(mb), 0))
        got, gmask = acc.compute()                      # waits for the chain
        api.device_synchronize()
    finally:
        assert api.set_stream_ordered(False) is True
    assert acc.accumulated_frames() == len(frames)
    assert np.array_equal(got, want) and np.array_equal(gmask, wmask)


def test_jdr_chain_stream_ordered_equals_blocking(gpu):
    import torch
    from serstacker_b200 import api, capi
    from oracle import derotation as od
    from test_gpu_derotation import _jovian_frame
    size, center, axes = (400, 320), (201.3, 158.6), (120.0, 112.0, 120.0)
    target = (0.4, math.radians(2.5), math.radians(-8.0))
    dlons = [math.radians(v) for v in (-4.0, 0.0, 3.0, 5.5)]
    frames = [_jovian_frame(size, center, axes, (target[0] + dl, target[1], target[2]), seed=i) for i, dl in enumerate(dlons)]
    jov = api.c_jovian_derotation_remap()
    jov.set_reference_pose(size, center, axes, target)
    period = jov.rotation_period_sec

    def run(acc, mats):
        for i, (dl, f) in enumerate(zip(dlons, mats)):
            jov.derotate_and_add(acc, f, None, dl * period / (2 * math.pi), 1.0 / (1.0 + abs(dl) * 10), i == 1, True,
                                 lpg_k=2.0, lpg_p=2.0, lpg_dscale=1, lpg_uscale=3)

    blocking = api.c_weigthed_average()
    run(blocking, frames)
    want, wmask = blocking.compute()

    dev = torch.device("cuda", 0)
    dfr = [torch.from_numpy(f).to(dev) for f in frames]
    torch.cuda.synchronize()


    acc = api.c_weigthed_average()
    api.set_stream_ordered(True)
    try:
        d = lambda v, n: (C.c_double * n)(*[float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1)])
        for i, (dl, t) in enumerate(zip(dlons, dfr)):
            Rc = api.build_ellipsoid_rotation((target[0] + dl, target[1], target[2]))
            m = capi.device_mat(t.data_ptr(), size[1], size[0], np.float32)
            capi.check(capi.lib.ssk_jdr_derotate_and_add(acc._h, C.byref(m), None, d(center, 2), d(axes, 3), d(Rc, 9), d(jov.Rtarget, 9),
                                                         float(jov.ebox[2]), (C.c_int * 4)(*jov.crop_box), 1.0 / (1.0 + abs(dl) * 10), int(i == 1), 1,
                                                         2.0, 2.0, 1, 3))
        got, gmask = acc.compute()
    finally:
        api.set_stream_ordered(False)
    assert np.array_equal(got, want) and np.array_equal(gmask, wmask)
