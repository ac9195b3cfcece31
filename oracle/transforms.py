"""
Oracle restatement of the parametric image transforms.

Follows /root/reference/core/proc/image_registration/c_image_transform.{h,cc}:
  translation  c_image_transform.cc:76-269,  c_image_transform.h:159-167
  euclidean    c_image_transform.cc:273-833
  affine       c_image_transform.cc:837-1070, c_image_transform.h:305-318
  homography   c_image_transform.cc:1075-1366, c_image_transform.h:372-384
and the factory image_transform.cc:34-63.

Parameters are kept as float32 vectors exactly like the reference's `cv::Mat1f _parameters` (Mx1).
All map arithmetic is done in float32 in the reference's operand order.

Test infrastructure only (see oracle/__init__.py).
"""
import math
import numpy as np
import cv2

f32 = np.float32

# image_transform.h:14-25
IMAGE_MOTION_TRANSLATION = 0
IMAGE_MOTION_EUCLIDEAN = 1
IMAGE_MOTION_SCALED_EUCLIDEAN = 2
IMAGE_MOTION_AFFINE = 3
IMAGE_MOTION_HOMOGRAPHY = 4


def _grid(size):
    w, h = size
    x = np.arange(w, dtype=f32)[None, :]
    y = np.arange(h, dtype=f32)[:, None]
    return x, y


class ImageTransform:
    """c_image_transform (c_image_transform.h:42-117)."""

    def parameters(self):
        return self._p

    def clone_parameters(self):
        return self._p.copy()

    def invertible(self):
        return False

    def create_remap(self, size, p=None):
        """size = (width, height) -> HxWx2 float32 map."""
        raise NotImplementedError


class TranslationTransform(ImageTransform):
    """c_translation_image_transform."""
    motion = IMAGE_MOTION_TRANSLATION

    def __init__(self, tx=0.0, ty=0.0):
        self._p = np.array([tx, ty], dtype=f32)

    def reset(self):
        self._p = np.zeros(2, dtype=f32)

    def set_translation(self, T):
        self._p = np.array([T[0], T[1]], dtype=f32)

    def translation(self):
        return (f32(self._p[0]), f32(self._p[1]))

    def set_parameters(self, p):
        p = np.asarray(p, dtype=f32).reshape(-1)
        assert p.size == 2
        self._p = p.copy()
        return True

    def scale_transfrom(self, factor):
        # c_image_transform.cc:130-134  (float *= double)
        self._p[0] = f32(float(self._p[0]) * factor)
        self._p[1] = f32(float(self._p[1]) * factor)

    def eps(self, dp, size):
        # c_image_transform.cc:136-139: dp(0,1) of a continuous 2x1 Mat1f aliases dp(1,0).
        dp = np.asarray(dp, dtype=f32).reshape(-1)
        return math.sqrt(float(f32(dp[0] * dp[0] + dp[1] * dp[1])))

    def create_remap(self, size, p=None):
        # c_image_transform.cc:150-170
        p = self._p if p is None else np.asarray(p, dtype=f32).reshape(-1)
        x, y = _grid(size)
        w, h = size
        m = np.empty((h, w, 2), dtype=f32)
        m[..., 0] = x + p[0]
        m[..., 1] = y + p[1]
        return m

    def create_steepest_descent_images(self, gx, gy, p=None):
        # c_image_transform.cc:262-269
        return [gx, gy]

    def invertible(self):
        return True

    def invert_and_compose(self, p, dp):
        # c_image_transform.h:164-167
        return (np.asarray(p, f32).reshape(-1) - np.asarray(dp, f32).reshape(-1)).astype(f32)


class EuclideanTransform(ImageTransform):
    """c_euclidean_image_transform (c_image_transform.cc:273-833)."""

    def __init__(self, fix_scale=False, fix_rotation=False, fix_translation=False):
        self._T = np.zeros(2, dtype=f32)
        self._C = np.zeros(2, dtype=f32)
        self._angle = f32(0)
        self._scale = f32(1)
        self._fix_translation = fix_translation
        self._fix_rotation = fix_rotation
        self._fix_scale = fix_scale
        self.motion = IMAGE_MOTION_EUCLIDEAN if fix_scale else IMAGE_MOTION_SCALED_EUCLIDEAN
        self._update_parameters()

    def num_adjustable_parameters(self):
        return (0 if self._fix_translation else 2) + (0 if self._fix_rotation else 1) + (0 if self._fix_scale else 1)

    def _update_parameters(self):
        # c_image_transform.cc:312-341
        p = []
        if not self._fix_translation:
            p += [self._T[0], self._T[1]]
        if not self._fix_rotation:
            p += [self._angle]
        if not self._fix_scale:
            p += [self._scale]
        self._p = np.array(p, dtype=f32)

    def reset(self):
        self._T[:] = 0
        self._C[:] = 0
        self._angle = f32(0)
        self._scale = f32(1)
        self._update_parameters()

    def set_center(self, C):
        self._C = np.array(C, dtype=f32)

    def set_translation(self, T):
        self._T = np.array([T[0], T[1]], dtype=f32)
        self._update_parameters()

    def translation(self):
        return (f32(self._T[0]), f32(self._T[1]))

    def set_parameters(self, p):
        # c_image_transform.cc:356-382
        p = np.asarray(p, dtype=f32).reshape(-1)
        assert p.size == self.num_adjustable_parameters()
        i = 0
        if not self._fix_translation:
            self._T = np.array([p[0], p[1]], dtype=f32)
            i = 2
        if not self._fix_rotation:
            self._angle = f32(p[i]); i += 1
        if not self._fix_scale:
            self._scale = f32(p[i]); i += 1
        self._update_parameters()
        return True

    def get_parameters(self, p):
        # c_image_transform.cc:384-423 -> (Tx, Ty, angle, scale, Cx, Cy)
        p = np.asarray(p, dtype=f32).reshape(-1)
        i = 0
        if self._fix_translation:
            Tx, Ty = self._T
        else:
            Tx, Ty = p[0], p[1]; i = 2
        if self._fix_rotation:
            angle = self._angle
        else:
            angle = p[i]; i += 1
        if self._fix_scale:
            scale = self._scale
        else:
            scale = p[i]; i += 1
        return f32(Tx), f32(Ty), f32(angle), f32(scale), f32(self._C[0]), f32(self._C[1])

    def scale_transfrom(self, factor):
        # c_image_transform.cc:502-507 (Vec2f *= double)
        self._T = (self._T.astype(np.float64) * factor).astype(f32)
        self._C = (self._C.astype(np.float64) * factor).astype(f32)
        self._update_parameters()

    def eps(self, dp, size):
        # c_image_transform.cc:509-523
        dTx, dTy, da, ds, _, _ = self.get_parameters(dp)
        w, h = size
        sa = f32(math.sin(float(da)))
        sq = lambda v: f32(v) * f32(v)
        return float(np.sqrt(f32(sq(dTx) + sq(dTy) + sq(f32(w) * sa) + sq(f32(h) * sa) + sq(f32(max(w, h)) * ds))))

    def create_remap(self, size, p=None):
        # c_image_transform.cc:525-555
        p = self._p if p is None else p
        Tx, Ty, angle, scale, Cx, Cy = self.get_parameters(p)
        sa = f32(math.sin(float(angle)))
        ca = f32(math.cos(float(angle)))
        x, y = _grid(size)
        w, h = size
        xx = x - Cx
        yy = y - Cy
        m = np.empty((h, w, 2), dtype=f32)
        m[..., 0] = scale * (ca * xx - sa * yy) + Tx
        m[..., 1] = scale * (sa * xx + ca * yy) + Ty
        return m

    def create_steepest_descent_images(self, gx, gy, p=None):
        # c_image_transform.cc:587-668
        p = self._p if p is None else p
        Tx, Ty, angle, scale, Cx, Cy = self.get_parameters(p)
        sa = f32(math.sin(float(angle)))
        ca = f32(math.cos(float(angle)))
        h, w = gx.shape
        x, y = _grid((w, h))
        xx = x - Cx
        yy = y - Cy
        J = []
        if not self._fix_translation:
            J += [gx, gy]
        if not self._fix_rotation:
            J.append((scale * (-gx * (sa * xx + ca * yy) + gy * (ca * xx - sa * yy))).astype(f32))
        if not self._fix_scale:
            J.append((gx * (ca * xx - sa * yy) + gy * (sa * xx + ca * yy)).astype(f32))
        return J

    def invertible(self):
        return True

    @staticmethod
    def _matrix3x3(tX, tY, ang, scl, cX, cY):
        # lambda at c_image_transform.cc:768-782
        sa = f32(math.sin(float(ang)))
        ca = f32(math.cos(float(ang)))
        M = np.eye(3, dtype=f32)
        M[0, 0] = scl * ca
        M[0, 1] = -scl * sa
        M[0, 2] = tX - scl * ca * cX + scl * sa * cY
        M[1, 0] = scl * sa
        M[1, 1] = scl * ca
        M[1, 2] = tY - scl * sa * cX - scl * ca * cY
        return M

    def invert_and_compose(self, p, dp):
        # c_image_transform.cc:736-833
        Tx, Ty, angle, scale, Cx, Cy = self.get_parameters(p)
        dTx, dTy, dAngle, dScale, _, _ = self.get_parameters(dp)
        if self._fix_translation:
            dTx = dTy = f32(0)
        if self._fix_rotation:
            dAngle = f32(0)
        scale_dp = f32(1) if self._fix_scale else f32(f32(1) + dScale)
        Mp = self._matrix3x3(Tx, Ty, angle, scale, Cx, Cy)
        Mdp = self._matrix3x3(dTx, dTy, dAngle, scale_dp, Cx, Cy)
        Mdp_inv = np.eye(3, dtype=f32)
        Mdp_inv[:2, :] = cv2.invertAffineTransform(np.ascontiguousarray(Mdp[:2, :]))
        # Matx33f product (s += a(i,k)*b(k,j), float, in order)
        M_res = np.zeros((3, 3), dtype=f32)
        for i in range(3):
            for j in range(3):
                acc = f32(0)
                for k in range(3):
                    acc = f32(acc + f32(Mp[i, k] * Mdp_inv[k, j]))
                M_res[i, j] = acc
        m00, m10 = M_res[0, 0], M_res[1, 0]
        res_scale = scale if self._fix_scale else f32(np.sqrt(f32(m00 * m00 + m10 * m10)))
        res_angle = angle if self._fix_rotation else f32(math.atan2(float(m10), float(m00)))
        rca = f32(math.cos(float(res_angle)))
        rsa = f32(math.sin(float(res_angle)))
        res_Tx = Tx if self._fix_translation else f32(M_res[0, 2] + res_scale * rca * Cx - res_scale * rsa * Cy)
        res_Ty = Ty if self._fix_translation else f32(M_res[1, 2] + res_scale * rsa * Cx + res_scale * rca * Cy)
        out = []
        if not self._fix_translation:
            out += [res_Tx, res_Ty]
        if not self._fix_rotation:
            out.append(res_angle)
        if not self._fix_scale:
            out.append(res_scale)
        return np.array(out, dtype=f32)


class AffineTransform(ImageTransform):
    """c_affine_image_transform (c_image_transform.cc:837-1070)."""
    motion = IMAGE_MOTION_AFFINE

    def __init__(self):
        self.reset()

    def reset(self):
        self._p = np.array([1, 0, 0, 0, 1, 0], dtype=f32)

    def set_translation(self, T):
        # c_image_transform.cc:864-868
        self._p[2] = f32(T[0])
        self._p[5] = f32(T[1])

    def translation(self):
        return (f32(self._p[2]), f32(self._p[5]))

    def set_parameters(self, p):
        p = np.asarray(p, dtype=f32).reshape(-1)
        assert p.size == 6
        self._p = p.copy()
        return True

    def scale_transfrom(self, factor):
        # c_image_transform.cc:911-915
        self._p[2] = f32(float(self._p[2]) * factor)
        self._p[5] = f32(float(self._p[5]) * factor)

    def eps(self, dp, size):
        # c_image_transform.cc:917-924
        dp = np.asarray(dp, dtype=f32).reshape(-1)
        w, h = f32(size[0]), f32(size[1])
        sq = lambda v: f32(v) * f32(v)
        return float(np.sqrt(f32(sq(w * dp[0]) + sq(h * dp[1]) + sq(dp[2]) + sq(w * dp[3]) + sq(h * dp[4]) + sq(dp[5]))))

    def create_remap(self, size, p=None):
        # c_image_transform.cc:926-946
        a = self._p if p is None else np.asarray(p, dtype=f32).reshape(-1)
        x, y = _grid(size)
        w, h = size
        m = np.empty((h, w, 2), dtype=f32)
        m[..., 0] = a[0] * x + a[1] * y + a[2]
        m[..., 1] = a[3] * x + a[4] * y + a[5]
        return m

    def create_steepest_descent_images(self, gx, gy, p=None):
        # c_image_transform.cc:1045-1070
        h, w = gx.shape
        x, y = _grid((w, h))
        return [(gx * x).astype(f32), (gx * y).astype(f32), gx, (gy * x).astype(f32), (gy * y).astype(f32), gy]

    def invertible(self):
        return True

    def invert_and_compose(self, p, dp):
        # c_image_transform.h:312-318
        p = np.asarray(p, dtype=f32).reshape(2, 3)
        dp = np.asarray(dp, dtype=f32).reshape(2, 3)
        a = cv2.invertAffineTransform(np.ascontiguousarray(p))
        a = cv2.invertAffineTransform(np.ascontiguousarray((a + dp).astype(f32)))
        return a.astype(f32).reshape(-1)


class HomographyTransform(ImageTransform):
    """c_homography_image_transform (c_image_transform.cc:1075-1366)."""
    motion = IMAGE_MOTION_HOMOGRAPHY

    def __init__(self):
        self.reset()

    def _update_parameters(self):
        self._p = self._m.reshape(-1)[:8].astype(f32).copy()

    def reset(self):
        self._m = np.eye(3, dtype=f32)
        self._update_parameters()

    def matrix(self, p=None):
        # c_image_transform.cc:1154-1164: a22 comes from the object, not from p
        if p is None:
            return self._m
        p = np.asarray(p, dtype=f32).reshape(-1)
        m = np.empty(9, dtype=f32)
        m[:8] = p
        m[8] = self._m[2, 2]
        return m.reshape(3, 3)

    def set_translation(self, T):
        # c_image_transform.cc:1124-1133
        self._m[0, 2] = f32(T[0]) * self._m[2, 2]
        self._m[1, 2] = f32(T[1]) * self._m[2, 2]
        self._update_parameters()

    def translation(self):
        return (f32(self._m[0, 2] / self._m[2, 2]), f32(self._m[1, 2] / self._m[2, 2]))

    def set_parameters(self, p):
        self._m = self.matrix(p).copy()
        self._update_parameters()
        return True

    def scale_transfrom(self, factor):
        # c_image_transform.cc:1186-1194
        self._m[0, 2] = f32(float(self._m[0, 2]) * factor)
        self._m[1, 2] = f32(float(self._m[1, 2]) * factor)
        self._m[2, 0] = f32(float(self._m[2, 0]) / factor)
        self._m[2, 1] = f32(float(self._m[2, 1]) / factor)
        self._update_parameters()

    def eps(self, dp, size):
        # c_image_transform.cc:1196-1205
        dp = np.asarray(dp, dtype=f32).reshape(-1)
        w, h = f32(size[0]), f32(size[1])
        sq = lambda v: f32(v) * f32(v)
        return float(np.sqrt(f32(sq(dp[2]) + sq(dp[5]) + sq(w * dp[0]) + sq(h * dp[1]) + sq(w * dp[3]) + sq(h * dp[4]))))

    def create_remap(self, size, p=None):
        # c_image_transform.cc:1207-1223
        a = self.matrix(p)
        x, y = _grid(size)
        w_, h_ = size
        w = a[2, 0] * x + a[2, 1] * y + a[2, 2]
        m = np.empty((h_, w_, 2), dtype=f32)
        m[..., 0] = (a[0, 0] * x + a[0, 1] * y + a[0, 2]) / w
        m[..., 1] = (a[1, 0] * x + a[1, 1] * y + a[1, 2]) / w
        return m

    def create_steepest_descent_images(self, gx, gy, p=None):
        # c_image_transform.cc:1319-1366 (note the literal 1.f in the denominator)
        a = self.matrix(p)
        h, w = gx.shape
        x, y = _grid((w, h))
        den = (f32(1) / (x * a[2, 0] + y * a[2, 1] + f32(1))).astype(f32)
        hatX = -(x * a[0, 0] + y * a[0, 1] + a[0, 2]) * den
        hatY = -(x * a[1, 0] + y * a[1, 1] + a[1, 2]) * den
        ggx = (gx * den).astype(f32)
        ggy = (gy * den).astype(f32)
        gg = (hatX * ggx + hatY * ggy).astype(f32)
        return [(ggx * x).astype(f32), (ggx * y).astype(f32), ggx,
                (ggy * x).astype(f32), (ggy * y).astype(f32), ggy,
                (gg * x).astype(f32), (gg * y).astype(f32)]

    def invertible(self):
        return True

    def invert_and_compose(self, p, dp):
        # c_image_transform.h:379-384
        dp = np.asarray(dp, dtype=f32).reshape(-1)
        dm = np.zeros(9, dtype=f32)
        dm[:8] = dp
        dm = dm.reshape(3, 3)
        _, inv1 = cv2.invert(np.ascontiguousarray(self.matrix(p)))
        _, aii = cv2.invert(np.ascontiguousarray((inv1 + dm).astype(f32)))
        aii = (aii * (f32(1) / aii[2, 2])).astype(f32)
        return aii.reshape(-1)[:8].copy()


def remap_points(tr, rpts, p=None):
    """c_image_transform::remap(params, rpts, cpts) (c_image_transform.cc:232-249 translation, 557-585 euclidean, 1019-1033
    affine, 1294-1306 homography): float arithmetic in the reference's operand order.  The homography multiplies by 1/w where
    create_remap divides by w; the euclidean form calls the float overloads of sin / cos."""
    rpts = np.asarray(rpts, dtype=f32).reshape(-1, 2)
    x, y = rpts[:, 0], rpts[:, 1]
    out = np.empty_like(rpts)
    if isinstance(tr, TranslationTransform):
        q = tr.parameters() if p is None else np.asarray(p, dtype=f32).reshape(-1)
        out[:, 0] = x + q[0]
        out[:, 1] = y + q[1]
    elif isinstance(tr, EuclideanTransform):
        Tx, Ty, angle, scale, Cx, Cy = tr.get_parameters(tr.parameters() if p is None else p)
        sa, ca = np.sin(f32(angle)), np.cos(f32(angle))          # float32 in, float32 out (sinf / cosf)
        xx, yy = x - Cx, y - Cy
        out[:, 0] = scale * (ca * xx - sa * yy) + Tx
        out[:, 1] = scale * (sa * xx + ca * yy) + Ty
    elif isinstance(tr, AffineTransform):
        a = tr.parameters() if p is None else np.asarray(p, dtype=f32).reshape(-1)
        out[:, 0] = a[0] * x + a[1] * y + a[2]
        out[:, 1] = a[3] * x + a[4] * y + a[5]
    elif isinstance(tr, HomographyTransform):
        a = tr.matrix(p)
        w = f32(1) / (a[2, 0] * x + a[2, 1] * y + a[2, 2])
        out[:, 0] = (a[0, 0] * x + a[0, 1] * y + a[0, 2]) * w
        out[:, 1] = (a[1, 0] * x + a[1, 1] * y + a[1, 2]) * w
    else:
        raise ValueError("unsupported transform")
    return out


def create_image_transform(motion_type):
    """image_transform.cc:34-63."""
    if motion_type == IMAGE_MOTION_TRANSLATION:
        return TranslationTransform()
    if motion_type == IMAGE_MOTION_EUCLIDEAN:
        return EuclideanTransform(fix_scale=True)
    if motion_type == IMAGE_MOTION_SCALED_EUCLIDEAN:
        return EuclideanTransform()
    if motion_type == IMAGE_MOTION_AFFINE:
        return AffineTransform()
    if motion_type == IMAGE_MOTION_HOMOGRAPHY:
        return HomographyTransform()
    raise ValueError("unsupported motion type %r" % (motion_type,))
