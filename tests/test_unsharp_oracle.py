"""CPU checks of oracle/unsharp.py: the scalar models the CUDA kernels are written from, against cv2."""
import math
import numpy as np
import cv2

from oracle import unsharp as ou

f32 = np.float32


def test_gaussian_taps_formula_matches_cv2():
    for sigma in np.arange(0.2, 2.01, 0.05):
        sigma = float(sigma)
        taps = 2 * max(1, int(sigma * 5)) + 1
        cf = [math.exp(-0.5 / (sigma * sigma) * (i - (taps - 1) * 0.5) ** 2) for i in range(taps)]
        sm = 0.0
        for c in cf:
            sm += c
        k = np.array([c / sm for c in cf]).astype(f32)
        assert np.array_equal(k, cv2.getGaussianKernel(taps, sigma, cv2.CV_32F).reshape(-1))


def test_lowpass_fma_model_matches_cv2_on_vector_columns():
    rng = np.random.default_rng(0)
    H, W = 97, 136
    src = rng.random((H, W), dtype=f32)
    sigma = 1.0
    k = 2 * max(1, int(sigma * 5)) + 1
    G = cv2.getGaussianKernel(k, sigma, cv2.CV_32F)
    g, r = G.reshape(-1), k // 2
    want = cv2.sepFilter2D(src, -1, G, G, borderType=cv2.BORDER_REFLECT)

    def fma(a, b, c):
        return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(f32)

    P = np.pad(src, ((0, 0), (r, r)), mode="symmetric")
    acc = (P[:, 0:W] * g[0]).astype(f32)                       # RowVec_32f: taps in order, fma chain
    for i in range(1, k):
        acc = fma(P[:, i:i + W], g[i], acc)
    Q = np.pad(acc, ((r, r), (0, 0)), mode="symmetric")
    c = (Q[r:r + H] * g[r]).astype(f32)                         # SymmColumnVec_32f: centre tap, then fma over the pairs
    for i in range(1, r + 1):
        c = fma((Q[r + i:r + i + H] + Q[r - i:r - i + H]).astype(f32), g[r + i], c)
    assert np.array_equal(c, want)


def test_unsharp_identity_cases_and_pyramid_levels():
    rng = np.random.default_rng(1)
    src = rng.random((40, 50), dtype=f32)
    assert np.array_equal(ou.unsharp_mask(src, 0, 0.8), src)
    assert np.array_equal(ou.unsharp_mask(src, 1, 0), src)
    assert ou.lpass_pyramid_level(1080, 1920, 2.0) == (0, 0)
    assert ou.lpass_pyramid_level(1080, 1920, 2.5) == (1, 1)
    assert ou.lpass_pyramid_level(1080, 1920, 4.0) == (2, 5)
    assert ou.unsharp_mask(src, 3.0, 0.5).shape == src.shape


def test_small_symmetric_row_filter_model_matches_cv2():
    """cv2's SymmRowSmallVec_32f for general symmetric kernels: 3 taps fma(x0, k0, (x-1 + x1) k1); 5 taps
    fma(x-2 + x2, k2, fma(x0, k0, (x-1 + x1) k1)) - the form k_sepfilter uses for kernels up to 5 taps."""
    rng = np.random.default_rng(0)
    H, W = 64, 96
    src = rng.random((H, W), dtype=f32)
    one = np.array([[1.0]], dtype=f32)
    D = np.float64

    def fma(a, b, c):
        return (a.astype(D) * D(b) + c.astype(D)).astype(f32)

    for sigma in (0.3, 0.5):
        k = 2 * max(1, int(sigma * 5)) + 1
        G = cv2.getGaussianKernel(k, sigma, cv2.CV_32F)
        g, r = G.reshape(-1), k // 2
        want = cv2.sepFilter2D(src, -1, G, one, borderType=cv2.BORDER_REFLECT)
        P = np.pad(src, ((0, 0), (r, r)), mode="symmetric")
        x0 = P[:, r:r + W]
        a1 = (P[:, r - 1:r - 1 + W] + P[:, r + 1:r + 1 + W]).astype(f32)
        acc = fma(x0, g[r], (a1 * g[r + 1]).astype(f32))
        if r == 2:
            a2 = (P[:, 0:W] + P[:, 4:4 + W]).astype(f32)
            acc = fma(a2, g[4], acc)
        assert np.array_equal(acc, want), sigma
