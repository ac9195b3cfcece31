"""
Oracle restatement of debayer_nn2 (core/io/debayer.cc:827-1195): the bilinear Bayer demosaic read_input_frame applies to
raw Bayer frames (c_image_stacking_pipeline_base.cc:125-279) before registration.  Output is 3-channel BGR of the
source depth (CV_8U / CV_16U with integer rounding (c + sum) / n, CV_32F with the reference's left-to-right sums).

  debayer_nn2(raw, colorid)          vectorised restatement for the four RGGB-family patterns
  debayer_nn2_rggb_literal(raw)      the RGGB case transcribed loop by loop (first pair / interior / last pair of each row:
                                     debayer.cc:856-922) - pure Python, small images only; pins the vectorised form

Neighbour indexing at the borders is the reference's: row -1 -> 1, row H -> H-2, column -1 -> 1, column W -> W-2.
Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

COLORID_BAYER_RGGB, COLORID_BAYER_GRBG, COLORID_BAYER_GBRG, COLORID_BAYER_BGGR = 8, 9, 10, 11
_RED_AT = {COLORID_BAYER_RGGB: (0, 0), COLORID_BAYER_GRBG: (0, 1), COLORID_BAYER_GBRG: (1, 0), COLORID_BAYER_BGGR: (1, 1)}


def debayer_nn2(raw, colorid):
    raw = np.ascontiguousarray(raw)
    H, W = raw.shape
    if (H & 1) or (W & 1) or raw.ndim != 2:
        raise ValueError("Can not make debayer for uneven image size")
    integral = raw.dtype.kind in "ui"
    wide = np.int64 if integral else np.float32
    s = raw.astype(wide)
    ym = np.r_[1, np.arange(0, H - 1)]          # y - 1 (row -1 -> 1)
    yp = np.r_[np.arange(1, H), H - 2]          # y + 1 (row H -> H - 2)
    xm = np.r_[1, np.arange(0, W - 1)]
    xp = np.r_[np.arange(1, W), W - 2]
    T, B_ = s[ym], s[yp]
    L, R_ = s[:, xm], s[:, xp]
    TL, TR, BL, BR = T[:, xm], T[:, xp], B_[:, xm], B_[:, xp]
    if integral:
        diag = (2 + TL + TR + BL + BR) // 4
        cross = (2 + T + L + R_ + B_) // 4
        vert = (1 + T + B_) // 2
        horz = (1 + L + R_) // 2
    else:
        f4, f2 = np.float32(4), np.float32(2)
        diag = (((TL + TR) + BL) + BR) / f4       # (c2 + s0[x-1] + s0[x+1] + s2[x-1] + s2[x+1]) / 4
        cross = (((T + L) + R_) + B_) / f4         # (c2 + s0[x] + s1[x-1] + s1[x+1] + s2[x]) / 4
        vert = (T + B_) / f2
        horz = (L + R_) / f2
    ry, rx = _RED_AT[colorid]
    yy, xx = np.mgrid[0:H, 0:W]
    py, px = yy & 1, xx & 1
    is_r = (py == ry) & (px == rx)
    is_b = (py == 1 - ry) & (px == 1 - rx)
    g_on_r_row = ~is_r & ~is_b & (py == ry)
    g_on_b_row = ~is_r & ~is_b & (py != ry)
    out = np.empty((H, W, 3), dtype=wide)
    out[..., 2] = np.where(is_r, s, np.where(is_b, diag, np.where(g_on_r_row, horz, vert)))      # R
    out[..., 1] = np.where(is_r | is_b, cross, s)                                                 # G
    out[..., 0] = np.where(is_b, s, np.where(is_r, diag, np.where(g_on_r_row, vert, horz)))      # B
    if integral:
        info = np.iinfo(raw.dtype)
        out = np.clip(out, info.min, info.max)
    return out.astype(raw.dtype)


def debayer_nn2_rggb_literal(raw):
    """debayer.cc:856-922 transcribed (same-depth source and destination)."""
    raw = np.ascontiguousarray(raw)
    H, W = raw.shape
    integral = raw.dtype.kind in "ui"
    c1, c2 = (1, 2) if integral else (0, 0)
    conv = (lambda v: int(v)) if integral else (lambda v: np.float32(v))
    S = [[conv(v) for v in row] for row in raw]
    if integral:
        d4 = lambda a, b, c, d: (c2 + a + b + c + d) // 4
        d2 = lambda a, b: (c1 + a + b) // 2
    else:
        f = np.float32
        d4 = lambda a, b, c, d: f(f(f(f(a + b) + c) + d) / f(4))
        d2 = lambda a, b: f(f(a + b) / f(2))
    dst = np.zeros((H, W, 3), dtype=raw.dtype)
    for y1 in range(H):
        y0 = 1 if y1 == 0 else y1 - 1
        y2 = H - 2 if y1 == H - 1 else y1 + 1
        s0, s1, s2 = S[y0], S[y1], S[y2]
        o = []
        e2 = 2 if 2 < W else 0
        if not (y1 & 1):   # R G
            o += [d4(s0[1], s0[1], s2[1], s2[1]), d4(s0[0], s1[1], s1[1], s2[0]), s1[0],
                  d2(s0[1], s2[1]), s1[1], d2(s1[0], s1[e2])]
            for x in range(2, W - 2, 2):
                o += [d4(s0[x - 1], s0[x + 1], s2[x - 1], s2[x + 1]), d4(s0[x], s1[x - 1], s1[x + 1], s2[x]), s1[x],
                      d2(s0[x + 1], s2[x + 1]), s1[x + 1], d2(s1[x], s1[x + 2])]
            if W > 2:
                x = W - 2
                o += [d4(s0[x - 1], s0[x + 1], s2[x - 1], s2[x + 1]), d4(s0[x], s1[x - 1], s1[x + 1], s2[x]), s1[x],
                      d2(s0[x + 1], s2[x + 1]), s1[x + 1], d2(s1[x], s1[x])]
        else:              # G B
            o += [d2(s1[1], s1[1]), s1[0], d2(s0[0], s2[0]),
                  s1[1], d4(s0[1], s1[0], s1[e2], s2[1]), d4(s0[0], s0[e2], s2[0], s2[e2])]
            for x in range(2, W - 2, 2):
                o += [d2(s1[x - 1], s1[x + 1]), s1[x], d2(s0[x], s2[x]),
                      s1[x + 1], d4(s0[x + 1], s1[x], s1[x + 2], s2[x + 1]), d4(s0[x], s0[x + 2], s2[x], s2[x + 2])]
            if W > 2:
                x = W - 2
                o += [d2(s1[x - 1], s1[x + 1]), s1[x], d2(s0[x], s2[x]),
                      s1[x + 1], d4(s0[x + 1], s1[x], s1[x], s2[x + 1]), d4(s0[x], s0[x], s2[x], s2[x])]
        dst[y1] = np.array(o, dtype=raw.dtype).reshape(W, 3)
    return dst


def average_bayer_planes(raw):
    """average_bayer_planes, raw single-channel form (core/io/debayer.cc:277-328): (c1 + s00 + s01 + s10 + s11) / 4 per 2x2 cell,
    c1 = 2 with integer division for integer samples, float arithmetic in that order otherwise."""
    a, b, c, d = raw[0::2, 0::2], raw[0::2, 1::2], raw[1::2, 0::2], raw[1::2, 1::2]
    if raw.dtype == np.float32:
        f = np.float32
        return ((((f(0) + a) + b) + c) + d) / f(4)
    s = (2 + a.astype(np.int64) + b + c + d) // 4
    return s.astype(raw.dtype)
