"""
Oracle restatement of the frame accumulators.

Follows /root/reference/core/average/c_frame_accumulation.{h,cc}:
  _weighted_average_update     c_frame_accumulation.cc:20-129
  c_weigthed_average           c_frame_accumulation.cc:143-260
  _bayer_accumulate            c_frame_accumulation.cc:988-1126
  c_bayer_average              c_frame_accumulation.cc:1138-1334, channel ids c_frame_accumulation.h:230-234
COLORID values: core/io/debayer.h:19-35.

Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

f32 = np.float32

COLORID_MONO = 0
COLORID_BAYER_RGGB = 8
COLORID_BAYER_GRBG = 9
COLORID_BAYER_GBRG = 10
COLORID_BAYER_BGGR = 11

BAYER_B, BAYER_G, BAYER_R = 0, 1, 2


class WeightedAverage:
    """c_weigthed_average: running weighted mean A += (I - A) * w / (W + w), W += w."""

    def __init__(self):
        self.clear()

    def clear(self):
        self.accumulator = None
        self.weights = None
        self.accumulated_frames = 0

    def reinitialize(self, src, accw):
        # c_frame_accumulation.cc:171-180
        self.clear()
        self.accumulator = np.array(src, dtype=f32, copy=True)
        self.weights = np.array(accw, dtype=f32, copy=True)
        self.accumulated_frames = 1
        return True

    def add(self, src, weights=None):
        # c_frame_accumulation.cc:182-221 + :20-129
        cn = 1 if src.ndim == 2 else src.shape[2]
        if self.accumulated_frames < 1:
            self.accumulator = np.zeros(src.shape, dtype=f32)
            self.weights = np.zeros(src.shape[:2], dtype=f32)
            self.accumulated_frames = 0
        if src.shape != self.accumulator.shape:
            return False
        if weights is not None and weights.shape[:2] != src.shape[:2]:
            return False
        A = self.accumulator.reshape(src.shape[0], src.shape[1], cn)
        I = src.reshape(src.shape[0], src.shape[1], cn).astype(f32)
        W = self.weights
        if weights is None:
            W_new = W + f32(1)
            factor = (f32(1) / W_new).astype(f32)
            W[...] = W_new
            A += (I - A) * factor[..., None]
        elif weights.dtype == np.uint8:
            m = weights != 0
            W_new = W + f32(1)
            factor = (f32(1) / W_new).astype(f32)
            upd = A + (I - A) * factor[..., None]
            A[m] = upd[m]
            W[m] = W_new[m]
        elif weights.dtype == np.float32:
            m = weights > 0
            W_new = (W + weights).astype(f32)
            with np.errstate(divide="ignore", invalid="ignore"):
                factor = (weights / W_new).astype(f32)
            upd = A + (I - A) * factor[..., None]
            A[m] = upd[m]
            W[m] = W_new[m]
        else:
            return False
        self.accumulated_frames += 1
        return True

    def compute(self):
        # c_frame_accumulation.cc:223-248 -> (avg, mask)
        if self.accumulated_frames < 1:
            return None, None
        return self.accumulator.copy(), ((self.weights > 0).astype(np.uint8) * 255)

    def get_acc_counters(self):
        return self.weights.copy()


def generate_bayer_pattern_mask(size, colorid):
    """c_frame_accumulation.cc:1262-1334; size = (h, w)."""
    h, w = size
    pat = np.zeros((h, w), dtype=np.uint8)
    tbl = {COLORID_BAYER_RGGB: (BAYER_R, BAYER_G, BAYER_G, BAYER_B),
           COLORID_BAYER_GRBG: (BAYER_G, BAYER_R, BAYER_B, BAYER_G),
           COLORID_BAYER_GBRG: (BAYER_G, BAYER_B, BAYER_R, BAYER_G),
           COLORID_BAYER_BGGR: (BAYER_B, BAYER_G, BAYER_G, BAYER_R)}[colorid]
    h2, w2 = (h // 2) * 2, (w // 2) * 2
    pat[0:h2:2, 0:w2:2] = tbl[0]
    pat[0:h2:2, 1:w2:2] = tbl[1]
    pat[1:h2:2, 0:w2:2] = tbl[2]
    pat[1:h2:2, 1:w2:2] = tbl[3]
    return pat


class BayerAverage:
    """c_bayer_average: per-colour sum/count accumulation of raw Bayer samples."""

    def __init__(self):
        self.clear()
        self.colorid = COLORID_BAYER_RGGB

    def clear(self):
        self.accumulator = None
        self.counter = None
        self.rmap = None
        self.pattern = None
        self.accumulated_frames = 0

    def set_bayer_pattern(self, colorid):
        self.colorid = colorid
        if self.accumulator is not None:
            self.pattern = generate_bayer_pattern_mask(self.accumulator.shape[:2], colorid)

    def set_remap(self, rmap):
        self.rmap = rmap

    def add(self, src, weights=None):
        # c_frame_accumulation.cc:1175-1203 + :988-1126
        h, w = src.shape[:2]
        if self.accumulated_frames < 1:
            self.accumulator = np.zeros((h, w, 3), dtype=f32)
            self.counter = np.zeros((h, w, 3), dtype=f32)
            self.accumulated_frames = 0
            self.pattern = generate_bayer_pattern_mask((h, w), self.colorid)
        acc, cntr, pat = self.accumulator, self.counter, self.pattern
        yy, xx = np.mgrid[0:h, 0:w]
        if self.rmap is None:
            if weights is None:
                sel = np.ones((h, w), bool); wv = None
            elif weights.dtype == np.uint8:
                sel = weights != 0; wv = None
            else:
                sel = np.ones((h, w), bool); wv = weights
            cc = pat[sel]
            s = src[sel].astype(f32)
            if wv is None:
                acc[yy[sel], xx[sel], cc] += s
                cntr[yy[sel], xx[sel], cc] += f32(1)
            else:
                acc[yy[sel], xx[sel], cc] += (s * wv[sel]).astype(f32)
                cntr[yy[sel], xx[sel], cc] += wv[sel]
        else:
            p0 = self.rmap[..., 0]
            p1 = self.rmap[..., 1]
            sx = np.trunc(p0).astype(np.int64)     # (int) cast truncates toward zero
            sy = np.trunc(p1).astype(np.int64)
            ok = (sx >= 0) & (sx < w - 1) & (sy >= 0) & (sy < h - 1)
            if weights is None:
                wgt = np.ones((h, w), dtype=np.float64)
            elif weights.dtype == np.uint8:
                ok &= weights != 0
                wgt = np.ones((h, w), dtype=np.float64)
            else:
                wgt = weights.astype(np.float64)
            sxc = np.clip(sx, 0, w - 2)
            syc = np.clip(sy, 0, h - 2)
            # (src_x + 1 - p[0]) is evaluated in float, then widened to double (c_frame_accumulation.cc:1054-1057)
            ax = ((sx + 1).astype(f32) - p0).astype(np.float64)
            ay = ((sy + 1).astype(f32) - p1).astype(np.float64)
            bx = (p0 - sx.astype(f32)).astype(np.float64)
            by = (p1 - sy.astype(f32)).astype(np.float64)
            for dy, dx, s in ((0, 0, ax * ay * wgt), (0, 1, bx * ay * wgt), (1, 0, ax * by * wgt), (1, 1, bx * by * wgt)):
                c = pat[syc + dy, sxc + dx]
                v = src[syc + dy, sxc + dx].astype(np.float64) * s
                # acc (float) += double: the sum is formed in double, then narrowed
                a_old = acc[yy, xx, c].astype(np.float64)
                c_old = cntr[yy, xx, c].astype(np.float64)
                a_new = (a_old + v).astype(f32)
                c_new = (c_old + s).astype(f32)
                acc[yy[ok], xx[ok], c[ok]] = a_new[ok]
                cntr[yy[ok], xx[ok], c[ok]] = c_new[ok]
        self.accumulated_frames += 1
        return True

    def compute(self):
        # c_frame_accumulation.cc:1205-1238 -> (avg HxWx3, mask)
        if self.accumulated_frames < 1:
            return None, None
        img = np.zeros_like(self.accumulator)
        m = self.counter > 0
        img[m] = self.accumulator[m] / self.counter[m]
        mask = (m.any(axis=2).astype(np.uint8)) * 255
        return img, mask

    def get_acc_counters(self):
        # c_frame_accumulation.cc:1240-1250
        return (self.counter * np.array([1, 0.5, 1], dtype=f32)).astype(f32)
