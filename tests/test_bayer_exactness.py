"""The exactness arguments behind k_fused_bayer's conversion-free forms (csrc/ssk_bayer.cu: tap_weights, sample_d), checked
in numpy: the device forms the reference's float expressions (c_frame_accumulation.cc:1040-1075: `src_x + 1 - p[0]`,
`p[0] - src_x` in float, then widened; samples convertTo(CV_32F, 1 / (1 << bpp)), then widened) from double operands, which
is only legitimate where the float operations are exact."""
import numpy as np

f32, f64 = np.float32, np.float64


def _coords(rng, n):
    # coordinates as a remap produces them: anywhere in a 4096-wide frame, plus values hugging integers and powers of two
    u = rng.uniform(1.0, 4096.0, n).astype(f32)
    near = (rng.integers(1, 4096, n).astype(f32) + rng.choice(np.array([0.0, 2.0 ** -20, -2.0 ** -20, 0.5, 2.0 ** -12], f32), n)).astype(f32)
    pw = (2.0 ** rng.integers(0, 12, n)).astype(f32)
    pw = np.nextafter(pw, rng.choice(np.array([0.0, 1e9], f32), n)).astype(f32)
    return np.concatenate([u, near, pw, np.array([1.0, 1.5, 2.0, 4095.99], f32)])


def test_tap_weights_from_one_widening_equal_the_float_differences():
    rng = np.random.default_rng(1)
    u = _coords(rng, 200000)
    s = np.trunc(u).astype(np.int32)
    sel = s >= 1                                   # the device keeps the literal float form for s == 0
    u, s = u[sel], s[sel]
    a_ref = (f32(1) * (s + 1).astype(f32) - u).astype(f32).astype(f64)      # (double)((float)(s + 1) - u)
    b_ref = (u - s.astype(f32)).astype(f32).astype(f64)                     # (double)(u - (float)s)
    du, ds = u.astype(f64), s.astype(f64)
    assert np.array_equal((ds + 1.0) - du, a_ref)
    assert np.array_equal(du - ds, b_ref)


def test_literal_form_is_needed_below_one():
    # why s == 0 is excluded: 1 - u rounds in float for tiny or negative u
    u = np.array([1e-9, -1e-9, 3e-8], f32)
    a_ref = (f32(1) - u).astype(f32).astype(f64)
    assert not np.array_equal(1.0 - u.astype(f64), a_ref)


def test_integer_samples_times_power_of_two_scale_are_exact():
    for bits, bpp in ((8, 8), (16, 16), (16, 12), (16, 14)):
        v = np.arange(0, 1 << bits, dtype=np.int64)
        scale = f32(1.0) / f32(1 << bpp)
        ref = (v.astype(f32) * scale).astype(f32).astype(f64)                # (double)((float)v * scale)
        assert np.array_equal(v.astype(f64) * f64(scale), ref)


def test_two_pow_52_trick_widens_non_negative_ints_exactly():
    v = np.array([0, 1, 2, 255, 65535, 4095, 2 ** 31 - 1], dtype=np.int64)
    bits = (np.uint64(0x43300000) << np.uint64(32)) | v.astype(np.uint64)
    d = bits.view(f64) - f64(4503599627370496.0)
    assert np.array_equal(d, v.astype(f64))


def test_by_channel_walk_keeps_the_reference_tap_order_per_channel():
    """The reference updates the channel of taps 00, 01, 10, 11 in that order; the device updates G (top row first), R, B.
    For every pattern and footprint parity the per-channel tap sequences must coincide."""
    patterns = {"RGGB": [[2, 1], [1, 0]], "GRBG": [[1, 2], [0, 1]], "GBRG": [[1, 0], [2, 1]], "BGGR": [[0, 1], [1, 2]]}   # B=0 G=1 R=2
    for name, pat in patterns.items():
        pos = {c: [(r, q) for r in range(2) for q in range(2) if pat[r][q] == c] for c in (0, 1, 2)}
        (rR, cR), (rB, cB) = pos[2][0], pos[0][0]
        cg0 = [q for q in range(2) if pat[0][q] == 1][0]
        for sy in range(2):
            for sx in range(2):
                ref = {0: [], 1: [], 2: []}
                for k in range(4):
                    dy, dx = k >> 1, k & 1
                    ref[pat[(sy + dy) & 1][(sx + dx) & 1]].append((dy, dx))
                gdx = (sx ^ cg0 ^ sy) & 1
                dev = {1: [(0, gdx), (1, gdx ^ 1)], 2: [((sy ^ rR) & 1, (sx ^ cR) & 1)], 0: [((sy ^ rB) & 1, (sx ^ cB) & 1)]}
                assert dev == ref, (name, sy, sx, dev, ref)
