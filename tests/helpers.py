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
