"""
Oracle restatement of the jovian derotation map.

Follows /root/reference/core/proc/feature2d/ellipsoid.cc and ellipsoid.h:
  build_ellipsoid_rotation            ellipsoid.h:47-63 (+ build_rotation, core/proc/pose.h:18-58): R = Rz * Rx * Ry
  ellipsoid_bbox                      ellipsoid.cc:16-84
  ellipse_bounding_box / crop_box     ellipsoid.cc:279-328
  ellipsoid_from_cart2d / to_cart2d   ellipsoid.h:84-158
  compute_ellipsoid_zrotation_remap   ellipsoid.cc:206-277
and c_jovian_derotation_remap::compute_derotation_for_angle (c_jovian_derotation_remap.cc:47-60).

Double precision with numpy's single-rounded elementwise operations in the reference's evaluation order.
Test infrastructure only (see oracle/__init__.py).
"""
import math
import numpy as np
import cv2


def build_ellipsoid_rotation(longitude_rotation, tilt_to_earth, position_angle):
    ax, ay, az = tilt_to_earth, longitude_rotation, position_angle
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], np.float64)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], np.float64)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], np.float64)
    return Rz @ Rx @ Ry


def ellipsoid_bbox(center, A, B, C, R):
    """-> ((cx, cy), (width, height), angle_deg) like cv::RotatedRect (float members)."""
    RR = np.eye(4)
    RR[:3, :3] = R
    Q = RR @ np.diag([1 / (A * A), 1 / (B * B), 1 / (C * C), -1.0]) @ RR.T
    Qi = np.linalg.inv(Q)
    P = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1]], np.float64)
    conic = np.linalg.inv(P @ Qi @ P.T)
    q = conic[:2, :2]
    ok, vals, vecs = cv2.eigen(q)            # descending eigenvalues, eigenvectors as rows
    ax_x = 2 / math.sqrt(vals[1, 0])
    ax_y = 2 / math.sqrt(vals[0, 0])
    t0 = math.atan2(vecs[1, 1], vecs[1, 0])
    if t0 > math.pi:
        t0 -= 2 * math.pi
    if t0 < -math.pi:
        t0 += math.pi
    f = np.float32
    return (f(center[0]), f(center[1])), (f(ax_x), f(ax_y)), f(t0 * 180 / math.pi)


def ellipse_bounding_box(rc):
    (cx, cy), (w, h), angle = rc
    f = np.float32
    a = f(float(angle) * math.pi / 180)
    ca, sa = f(math.cos(a)), f(math.sin(a))
    ux, uy = f(w * ca / f(2)), f(-w * sa / f(2))
    vx, vy = f(h * sa / f(2)), f(h * ca / f(2))
    hw, hh = f(math.sqrt(ux * ux + vx * vx)), f(math.sqrt(uy * uy + vy * vy))
    left, top = f(cx - hw), f(cy - hh)
    return [int(left), int(top), int(f(2) * hw), int(f(2) * hh)]


def ellipse_crop_box(rc, image_size, margin=1):
    x, y, w, h = ellipse_bounding_box(rc)
    if margin < 0:
        margin = max(16, int(rc[1][0] / 5))
    x -= margin
    y -= margin
    w += 2 * margin
    h += 2 * margin
    x = max(x, 0)
    y = max(y, 0)
    if x + w >= image_size[0]:
        w = image_size[0] - x
    if y + h >= image_size[1]:
        h = image_size[1] - y
    return [x, y, w, h]


def compute_ellipsoid_zrotation_remap(size, center, axes, R1, R2, wscale=1.0):
    """size = (w, h) -> (rmap HxWx2 float32, wmap HxW float32, rmask HxW uint8, ebox, cbox)."""
    W, H = size
    A, B, C = [float(v) for v in axes]
    ebox = ellipsoid_bbox(center, A, B, C, R2)
    bx, by, bw, bh = ellipse_crop_box(ebox, size)
    angle = float(ebox[2]) * math.pi / 180
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    rmap = np.dstack([xx, yy]).astype(np.float32)
    rmask = np.zeros((H, W), np.uint8)
    inbox = (yy >= by) & (yy <= by + bh) & (xx >= bx) & (xx <= bx + bw)
    xs, ys = xx - center[0], yy - center[1]
    R = np.asarray(R2, np.float64)
    x_stat = R[0, 0] * xs + R[1, 0] * ys
    y_stat = R[0, 1] * xs + R[1, 1] * ys
    z_stat = R[0, 2] * xs + R[1, 2] * ys
    rzx, rzy, rzz = R[2, 0], R[2, 1], R[2, 2]
    iA, iB, iC = 1.0 / (A * A), 1.0 / (B * B), 1.0 / (C * C)
    K2 = (rzx * rzx) * iA + (rzy * rzy) * iB + (rzz * rzz) * iC
    K1 = (x_stat * rzx) * iA + (y_stat * rzy) * iB + (z_stat * rzz) * iC
    K0 = (x_stat * x_stat) * iA + (y_stat * y_stat) * iB + (z_stat * z_stat) * iC - 1.0
    disc = K1 * K1 - K2 * K0
    hit = inbox & ~(disc < 0.0)
    sq = np.sqrt(np.where(hit, disc, 0.0))
    zs = np.minimum((-K1 - sq) / K2, (-K1 + sq) / K2)
    vx = R[0, 0] * xs + R[1, 0] * ys + R[2, 0] * zs
    vy = R[0, 1] * xs + R[1, 1] * ys + R[2, 1] * zs
    vz = R[0, 2] * xs + R[1, 2] * ys + R[2, 2] * zs
    Q = np.asarray(R1, np.float64)
    px = Q[0, 0] * vx + Q[0, 1] * vy + Q[0, 2] * vz
    py = Q[1, 0] * vx + Q[1, 1] * vy + Q[1, 2] * vz
    pz = Q[2, 0] * vx + Q[2, 1] * vy + Q[2, 2] * vz
    vis = hit & (pz <= 0.0)
    rmask[hit] = 255
    rmap[..., 0] = np.where(vis, (px + center[0]).astype(np.float32), np.where(hit, np.float32(-1), rmap[..., 0]))
    rmap[..., 1] = np.where(vis, (py + center[1]).astype(np.float32), np.where(hit, np.float32(-1), rmap[..., 1]))
    wmap = np.zeros((H, W), np.float32)
    a, b, sa, ca = 1.0 / A, 1.0 / B, math.sin(angle), math.cos(angle)
    inw = hit & (yy >= by) & (yy < by + bh) & (xx >= bx) & (xx < bx + bw)
    ex = (xs * ca + ys * sa) * a
    ey = (-xs * sa + ys * ca) * b
    rr = ex * ex + ey * ey
    ok = inw & (rr <= 1.0)
    wmap[ok] = (wscale * np.sqrt(np.maximum(0.0, 1.0 - rr[ok]))).astype(np.float32)
    wmap = cv2.remap(wmap, rmap, None, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    return rmap, wmap, rmask, ebox, [bx, by, bw, bh]


def compute_derotation_for_angle(size, center, axes, target_pose, longitude_rotation_radians, wscale=1.0):
    """c_jovian_derotation_remap::compute_derotation_for_angle."""
    Rt = build_ellipsoid_rotation(*target_pose)
    Rc = build_ellipsoid_rotation(target_pose[0] + longitude_rotation_radians, target_pose[1], target_pose[2])
    return compute_ellipsoid_zrotation_remap(size, center, axes, Rc, Rt, wscale)


def jdr_derotate_and_add(acc, frame, mask, size, center, axes, target_pose, longitude_rotation_radians, wscale, is_master,
                         enable_weighted_average=True, lpg_opts=None):
    """One iteration of c_jdr_pipeline::derotate_and_average_frames after preproc_align_and_remap
    (c_jdr_pipeline.cc:1184-1236): returns (derotated frame, weights) and adds them to `acc`."""
    from .weights import lpg
    rmap, wmap, rmask, ebox, cbox = compute_derotation_for_angle(size, center, axes, target_pose,
                                                                 longitude_rotation_radians, wscale)
    w = wmap.copy()
    w[w < 1e-5] = 0
    if enable_weighted_average:
        l = lpg(frame, **(lpg_opts or {}))
        ld = l.copy()                                           # cv::remap in place: the destination starts as the source
        cv2.remap(l, rmap, None, cv2.INTER_LINEAR, dst=ld, borderMode=cv2.BORDER_TRANSPARENT)
        w = cv2.multiply(w, ld)
    if is_master:
        w[rmask == 0] = 1
    if mask is not None:
        w[mask == 0] = 0
    w = cv2.GaussianBlur(w, (0, 0), 1, None, 1, cv2.BORDER_REPLICATE)
    f = frame.copy()
    cv2.remap(frame, rmap, None, cv2.INTER_LINEAR, dst=f, borderMode=cv2.BORDER_TRANSPARENT)
    acc.add(f, w)
    return f, w
