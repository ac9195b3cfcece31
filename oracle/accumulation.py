"""
Oracle restatement of the frame accumulators.

Follows /root/reference/core/average/c_frame_accumulation.{h,cc}:
  _weighted_average_update     c_frame_accumulation.cc:20-129
  c_weigthed_average           c_frame_accumulation.cc:143-260
  _bayer_accumulate            c_frame_accumulation.cc:988-1126
  c_bayer_average              c_frame_accumulation.cc:1138-1334, channel ids c_frame_accumulation.h:230-234
  c_canvas_average             c_frame_accumulation.h:65-137, c_frame_accumulation.cc:264-445
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


def _weighted_average_update(I, weights, A, W):
    """_weighted_average_update (c_frame_accumulation.cc:20-129) on views: A (h, w, cn) and W (h, w) are updated in place.
    Returns False on the size / type mismatches the reference rejects."""
    if A.shape[:2] != I.shape[:2] or W.shape != I.shape[:2]:
        return False
    cn = A.shape[2]
    I = I.reshape(I.shape[0], I.shape[1], cn).astype(f32)
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
    else:
        m = weights > 0
        W_new = (W + weights).astype(f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            factor = (weights / W_new).astype(f32)
        upd = A + (I - A) * factor[..., None]
        A[m] = upd[m]
        W[m] = W_new[m]
    return True


class CanvasAverage:
    """c_canvas_average (c_frame_accumulation.h:65-137, c_frame_accumulation.cc:264-445): a weighted average on a canvas
    larger than the frames; every frame is remapped into a bounding box of the canvas, and the canvas content is shifted by
    64 px when a box comes within 32 px of an edge."""

    def __init__(self, interpolation=None, canvas_size=(0, 0)):
        import cv2
        self.interpolation = cv2.INTER_LINEAR if interpolation is None else interpolation   # c_frame_accumulation.h:71-73
        self.canvas_size = tuple(canvas_size)    # (w, h), setCanvasSize
        self.clear()

    def clear(self):
        self.accumulator = None
        self.weights = None
        self.accumulated_frames = 0
        self.last_bbox = (0, 0, 0, 0)            # x, y, w, h

    @staticmethod
    def compute_canvas_size(frame_size):
        w, h = frame_size
        return (3 * w // 2, 3 * h // 2)

    @staticmethod
    def _intersect(a, b):
        x0, y0 = max(a[0], b[0]), max(a[1], b[1])
        x1, y1 = min(a[0] + a[2], b[0] + b[2]), min(a[1] + a[3], b[1] + b[3])
        return (x0, y0, x1 - x0, y1 - y0) if (x1 > x0 and y1 > y0) else (0, 0, 0, 0)

    def _maintain_canvas_boundaries(self, bbox):
        # c_frame_accumulation.cc:272-322
        x, y, w, h = bbox
        if self.accumulator is None or w <= 0 or h <= 0:
            return bbox
        margin = 32
        H, Wd = self.accumulator.shape[:2]
        sx = 2 * margin if x < margin else (-2 * margin if x + w >= Wd - margin else 0)
        sy = 2 * margin if y < margin else (-2 * margin if y + h >= H - margin else 0)
        if sx or sy:
            cw, chh = Wd - abs(sx), H - abs(sy)
            if cw > 0 and chh > 0:
                src_x, src_y = (0 if sx > 0 else -sx), (0 if sy > 0 else -sy)
                dst_x, dst_y = (sx if sx > 0 else 0), (sy if sy > 0 else 0)
                na, nw = np.zeros_like(self.accumulator), np.zeros_like(self.weights)
                na[dst_y:dst_y + chh, dst_x:dst_x + cw] = self.accumulator[src_y:src_y + chh, src_x:src_x + cw]
                nw[dst_y:dst_y + chh, dst_x:dst_x + cw] = self.weights[src_y:src_y + chh, src_x:src_x + cw]
                self.accumulator, self.weights = na, nw
                x, y = x + sx, y + sy
        return (x, y, w, h)

    def add(self, image, weights_or_mask=None, rmap=None, new_canvas_bbox=None):
        import cv2
        # c_frame_accumulation.cc:325-400
        if rmap is not None and self.accumulator is not None and (rmap.shape[1] > self.accumulator.shape[1] or rmap.shape[0] > self.accumulator.shape[0]):
            return False
        img = image if image.ndim == 3 else image[..., None]
        h, w, cn = img.shape
        if self.accumulator is None:
            cs = self.compute_canvas_size((w, h))
            cw, chh = max(self.canvas_size[0], cs[0]), max(self.canvas_size[1], cs[1])
            tx, ty = cw // 2 - w // 2, chh // 2 - h // 2
            self.accumulator = np.zeros((chh, cw, cn), f32)
            self.weights = np.zeros((chh, cw), f32)
            _weighted_average_update(img, weights_or_mask, self.accumulator[ty:ty + h, tx:tx + w], self.weights[ty:ty + h, tx:tx + w])
            self.last_bbox = (tx, ty, w, h)
            self.accumulated_frames += 1
            return True
        H, Wd = self.accumulator.shape[:2]
        if (new_canvas_bbox is None or new_canvas_bbox[2] <= 0 or new_canvas_bbox[3] <= 0) and rmap is None:
            rimg, rw = img, weights_or_mask
            roi = self._intersect((self.last_bbox[0], self.last_bbox[1], w, h), (0, 0, Wd, H))
        else:
            roi = self._intersect(tuple(new_canvas_bbox), (0, 0, Wd, H))
            if roi[2] <= 0:
                return False
            roi = self._maintain_canvas_boundaries(roi)
            rimg = cv2.remap(image, rmap, None, self.interpolation, borderMode=cv2.BORDER_REPLICATE)
            rimg = rimg if rimg.ndim == 3 else rimg[..., None]
            rw = None
            if weights_or_mask is not None:
                mi = cv2.INTER_NEAREST if weights_or_mask.dtype == np.uint8 else cv2.INTER_LINEAR
                rw = cv2.remap(weights_or_mask, rmap, None, mi, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        x, y, rwid, rh = roi
        if x < 0 or y < 0 or x + rwid > Wd or y + rh > H:
            # cv::Mat::operator()(Rect) asserts the box lies inside the matrix (the shift can push it out on a small canvas)
            raise ValueError("c_canvas_average: ROI outside the canvas after maintainCanvasBoundaries")
        # the reference ignores the update's result (a size mismatch skips the frame but still counts it)
        _weighted_average_update(rimg, rw, self.accumulator[y:y + rh, x:x + rwid], self.weights[y:y + rh, x:x + rwid])
        self.last_bbox = roi
        self.accumulated_frames += 1
        return True

    def compute(self, rbbox=None):
        # c_frame_accumulation.cc:405-437 -> (avg, mask) of the requested box (whole canvas when rbbox is None)
        if self.accumulated_frames < 1:
            return None, None
        H, Wd = self.accumulator.shape[:2]
        box = (0, 0, Wd, H) if rbbox is None or rbbox[2] <= 0 or rbbox[3] <= 0 else self._intersect(tuple(rbbox), (0, 0, Wd, H))
        if box[2] <= 0:
            return None, None
        x, y, w, h = box
        avg = self.accumulator[y:y + h, x:x + w].copy()
        if avg.shape[2] == 1:
            avg = avg[..., 0]
        return avg, ((self.weights[y:y + h, x:x + w] > 0).astype(np.uint8) * 255)
