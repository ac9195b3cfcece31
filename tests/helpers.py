"""Shared helpers of the parity tests."""
import numpy as np

from oracle import transforms as otf


def oracle_transform(motion_type, params):
    t = otf.create_image_transform(motion_type)
    t.set_parameters(np.asarray(params, np.float32))
    return t


def map_diff_px(motion_type, p_a, p_b, size):
    """max |W_a(x) - W_b(x)| over the frame corners and centre, in pixels: the '1e-3 px' parity metric."""
    w, h = size
    ta, tb = oracle_transform(motion_type, p_a), oracle_transform(motion_type, p_b)
    ma, mb = ta.create_remap((w, h)), tb.create_remap((w, h))
    pts = [(0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (h // 2, w // 2)]
    return max(float(np.abs(ma[y, x] - mb[y, x]).max()) for y, x in pts)


def rel_l2(a, b, mask=None):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    if mask is not None:
        a, b = a[mask], b[mask]
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30))


# ---------------------------------------------------------------------------------------------------------
# Sensitivity envelope of the reference algorithm.
#
# The ECC normal equations of the multi-parameter transforms are badly conditioned in pixel coordinates
# (cond(H) ~ 4e6 for affine at 320x240, measured in tests/debug_trace.py), H and v are stored as float, and the
# over-relaxed / LM iterations stop on coarse thresholds.  One-ulp differences in the sums therefore move the
# final parameters by up to 1e-2..1e-1 px for the least stable solvers (forward-additive, euclidean whose eps()
# never falls below the threshold).  The reference's own cv::Mat::dot accumulates float blocks of 8192
# elements (relative error ~1e-6 against exact summation), so its answer is only defined up to that envelope.
# Where the GPU path cannot be bit-identical to the oracle (closed-form 3x3 solve, sin/cos, filter tail columns)
# parity is asserted against this envelope instead of a fixed 1e-3 px.
# ---------------------------------------------------------------------------------------------------------
class dot_noise:
    """Context manager: multiplies every oracle dot product / squared norm by (1 + rel * U(-1, 1))."""

    def __init__(self, rel=1e-6, seed=1234):
        self.rel, self.seed = rel, seed

    def __enter__(self):
        from oracle import ecc as oecc
        self.m = oecc
        self.d, self.n = oecc._dot, oecc._norm_l2sqr
        rng = np.random.default_rng(self.seed)
        rel = self.rel
        oecc._dot = lambda a, b, _f=self.d: _f(a, b) * (1.0 + rel * rng.uniform(-1, 1))
        oecc._norm_l2sqr = lambda a, _f=self.n: _f(a) * (1.0 + rel * rng.uniform(-1, 1))
        return self

    def __exit__(self, *a):
        self.m._dot, self.m._norm_l2sqr = self.d, self.n


def strict_case(motion, method):
    """Configurations whose GPU trajectory is bit-faithful to the oracle: translation (any solver) and
    affine with the inverse-compositional / LM solvers."""
    return motion == 0 or (motion == 3 and method in (1, 2, 3))
