"""c_image_transform::eps / invert_and_compose through the C ABI (host arithmetic of libssk, csrc/ssk_xf_host.cu) against
oracle/transforms.py (c_image_transform.cc:136-139, 509-523, 736-833, 917-924, 1196-1205; c_image_transform.h:164-167, 312-318,
379-384): bit for bit, five motion types, random parameters and steps.  No GPU needed: the library loads without a device."""
import numpy as np
import pytest

from oracle import transforms as otf

f32 = np.float32


def _random_case(motion, rng):
    if motion == otf.IMAGE_MOTION_TRANSLATION:
        p, dp = rng.normal(0, 20, 2), rng.normal(0, 0.5, 2)
    elif motion == otf.IMAGE_MOTION_EUCLIDEAN:
        p, dp = [rng.normal(0, 20), rng.normal(0, 20), rng.normal(0, 0.05)], [rng.normal(0, 0.5), rng.normal(0, 0.5), rng.normal(0, 1e-3)]
    elif motion == otf.IMAGE_MOTION_SCALED_EUCLIDEAN:
        p = [rng.normal(0, 20), rng.normal(0, 20), rng.normal(0, 0.05), 1 + rng.normal(0, 0.02)]
        dp = [rng.normal(0, 0.5), rng.normal(0, 0.5), rng.normal(0, 1e-3), rng.normal(0, 1e-3)]
    elif motion == otf.IMAGE_MOTION_AFFINE:
        p = np.array([1, 0, 0, 0, 1, 0], float) + np.concatenate([rng.normal(0, 0.02, 2), rng.normal(0, 20, 1), rng.normal(0, 0.02, 2), rng.normal(0, 20, 1)])
        dp = np.concatenate([rng.normal(0, 1e-3, 2), rng.normal(0, 0.5, 1), rng.normal(0, 1e-3, 2), rng.normal(0, 0.5, 1)])
    else:
        p = np.array([1, 0, 0, 0, 1, 0, 0, 0], float) + np.concatenate([rng.normal(0, 0.02, 2), rng.normal(0, 20, 1), rng.normal(0, 0.02, 2),
                                                                          rng.normal(0, 20, 1), rng.normal(0, 1e-5, 2)])
        dp = np.concatenate([rng.normal(0, 1e-3, 2), rng.normal(0, 0.5, 1), rng.normal(0, 1e-3, 2), rng.normal(0, 0.5, 1), rng.normal(0, 1e-6, 2)])
    return np.asarray(p, f32), np.asarray(dp, f32)


@pytest.mark.parametrize("motion", [0, 1, 2, 3, 4])
def test_eps_and_invert_and_compose_match_oracle(motion):
    from serstacker_b200 import api
    rng = np.random.default_rng(100 + motion)
    for it in range(200):
        p, dp = _random_case(motion, rng)
        size = (int(rng.integers(32, 2000)), int(rng.integers(32, 2000)))
        o = otf.create_image_transform(motion)
        g = api.create_image_transform(motion)
        o.set_parameters(p)
        g.set_parameters(p)
        if motion in (otf.IMAGE_MOTION_EUCLIDEAN, otf.IMAGE_MOTION_SCALED_EUCLIDEAN) and it % 2:
            c = (f32(size[0] / 2), f32(size[1] / 2))              # rotation centre (set_center, c_image_transform.cc:300-310)
            o.set_center(c)
            g.t.aux[0], g.t.aux[1] = float(c[0]), float(c[1])
        assert g.eps(dp, size) == o.eps(dp, size), (motion, it)
        want = np.asarray(o.invert_and_compose(o.parameters(), dp), f32).reshape(-1)
        got = g.invert_and_compose(dp)
        assert np.array_equal(got, want), (motion, it, got, want)


def test_identity_step_and_bad_arguments():
    from serstacker_b200 import api, capi
    g = api.create_image_transform(3)
    g.set_parameters([1.01, 0.002, 3.5, -0.001, 0.99, -2.25])
    assert g.eps(np.zeros(6, f32), (640, 480)) == 0.0
    back = g.invert_and_compose(np.zeros(6, f32))                 # inverting twice: the parameters up to float rounding
    assert np.allclose(back, g.parameters(), rtol=0, atol=1e-5)
    with pytest.raises(capi.SskError):
        g.eps(np.zeros(5, f32), (640, 480))
    with pytest.raises(capi.SskError):
        g.invert_and_compose(np.zeros(8, f32))


@pytest.mark.parametrize("motion", [0, 1, 2, 3, 4])
def test_remap_points_matches_oracle(motion):
    """c_image_transform::remap(params, rpts, cpts) (c_image_transform.cc:232-249, 557-585, 1019-1033, 1294-1306)."""
    from serstacker_b200 import api
    rng = np.random.default_rng(300 + motion)
    for it in range(50):
        p, _ = _random_case(motion, rng)
        o = otf.create_image_transform(motion)
        g = api.create_image_transform(motion)
        o.set_parameters(p)
        g.set_parameters(p)
        if motion in (otf.IMAGE_MOTION_EUCLIDEAN, otf.IMAGE_MOTION_SCALED_EUCLIDEAN) and it % 2:
            o.set_center((f32(320), f32(240)))
            g.t.aux[0], g.t.aux[1] = 320.0, 240.0
        pts = rng.uniform(-50, 2000, (257, 2)).astype(f32)
        want = otf.remap_points(o, pts)
        got = g.remap_points(pts)
        assert np.array_equal(got, want), (motion, it, np.abs(got - want).max())
    assert g.remap_points(np.zeros((0, 2), f32)).shape == (0, 2)


@pytest.mark.parametrize("motion", [0, 3])
def test_remap_points_on_the_pixel_grid_is_create_remap(motion):
    """Translation and affine maps use the same expression for points and for the dense map (c_image_transform.cc:150-170,
    926-946): the point form at integer positions must reproduce the oracle's create_remap, which the GPU parity tests pin."""
    from serstacker_b200 import api
    rng = np.random.default_rng(400 + motion)
    p, _ = _random_case(motion, rng)
    o = otf.create_image_transform(motion)
    g = api.create_image_transform(motion)
    o.set_parameters(p)
    g.set_parameters(p)
    w, h = 61, 47
    yy, xx = np.mgrid[0:h, 0:w].astype(f32)
    pts = np.stack([xx.ravel(), yy.ravel()], axis=1)
    assert np.array_equal(g.remap_points(pts).reshape(h, w, 2), o.create_remap((w, h)))
