"""Host geometry of c_jovian_derotation_remap / c_saturn_derotation_remap in libssk (ssk_build_ellipsoid_rotation,
ssk_ellipsoid_bbox: ellipsoid.h:47-71, ellipsoid.cc:16-84, 279-328) against the oracle restatement over numpy / cv2.eigen.
No device work: runs without a GPU."""
import math
import numpy as np

from oracle import derotation as od


def test_rotation_and_bbox_match_oracle():
    from serstacker_b200 import api
    rng = np.random.default_rng(5)
    worst_R, n_float_diff = 0.0, 0
    for it in range(300):
        pose = (rng.uniform(-math.pi, math.pi), rng.uniform(-0.6, 0.6), rng.uniform(-math.pi, math.pi))
        A = rng.uniform(40, 400)
        axes = (A, A * rng.uniform(0.8, 1.0), A * rng.uniform(0.9, 1.1))
        size = (int(rng.integers(300, 2000)), int(rng.integers(300, 2000)))
        center = (rng.uniform(0.3, 0.7) * size[0], rng.uniform(0.3, 0.7) * size[1])
        Ro = od.build_ellipsoid_rotation(*pose)
        Rg = api.build_ellipsoid_rotation(pose)
        worst_R = max(worst_R, float(np.abs(Ro - Rg).max()))
        eo = od.ellipsoid_bbox(center, *axes, Ro)
        co = od.ellipse_crop_box(eo, size)
        eg, cg = api.ellipsoid_bbox(size, center, axes, Ro)
        fo = np.array([eo[0][0], eo[0][1], eo[1][0], eo[1][1], eo[2]], np.float32)
        fg = np.array([eg[0][0], eg[0][1], eg[1][0], eg[1][1], eg[2]], np.float32)
        assert np.array_equal(fo[:2], fg[:2])
        # widths and angle: double arithmetic rounded to float; the 4 x 4 LU inverse of the reference against the closed form here
        assert np.all(np.abs(fo[2:4] - fg[2:4]) <= np.spacing(np.abs(fo[2:4]))), (fo, fg)
        da = abs(float(fo[4]) - float(fg[4]))
        assert min(da, abs(da - 180.0), abs(da - 360.0)) <= 1e-4, (fo, fg)
        n_float_diff += int(not np.array_equal(fo, fg))
        assert all(abs(a - b) <= 1 for a, b in zip(co, cg)), (co, cg)
    print("rotation max |d| = %.3g, bounding ellipses differing in a float member: %d of 300" % (worst_R, n_float_diff))
    assert worst_R <= 1e-15


def test_derotation_classes_mirror_reference_periods():
    from serstacker_b200 import api
    assert api.c_jovian_derotation_remap().rotation_period_sec == 9. * 3600 + 55. * 60 + 40.632
    assert api.c_saturn_derotation_remap().rotation_period_sec == 10 * 3600. + 33 * 60. + 38
