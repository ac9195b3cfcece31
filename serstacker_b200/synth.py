"""
Deterministic synthetic inputs for the five BASELINE.json configs.

Pure data generation (numpy + cv2 blur); shared by tests, the oracle-side golden script and bench.py.
The scene is evaluated analytically at transformed coordinates (never through the code under test).
"""
import math
import numpy as np
import cv2

f32 = np.float32


class PlanetScene:
    """Limb-darkened disk + Gaussian-band belts + random Gaussian spots on a dim background."""

    def __init__(self, width, height, radius, seed, nspots=40, nbelts=6):
        rng = np.random.default_rng(seed)
        self.w, self.h, self.R = width, height, float(radius)
        self.cx, self.cy = (width - 1) / 2.0, (height - 1) / 2.0
        self.belts = [(rng.uniform(-0.8, 0.8) * radius, rng.uniform(0.03, 0.08) * radius, rng.uniform(-0.25, 0.25))
                      for _ in range(nbelts)]
        self.spots = [(rng.uniform(-0.7, 0.7) * radius, rng.uniform(-0.7, 0.7) * radius,
                       rng.uniform(2.0, 6.0) * max(1.0, radius / 150.0), rng.uniform(-0.3, 0.3))
                      for _ in range(nspots)]

    def render(self, A, background=0.02):
        """Render I(A @ [x, y, 1]) on the pixel grid; A is a 2x3 matrix mapping output pixel -> scene coords."""
        y, x = np.mgrid[0:self.h, 0:self.w].astype(np.float64)
        sx = A[0][0] * x + A[0][1] * y + A[0][2] - self.cx
        sy = A[1][0] * x + A[1][1] * y + A[1][2] - self.cy
        r2 = (sx * sx + sy * sy) / (self.R * self.R)
        inside = r2 < 1.0
        mu = np.sqrt(np.clip(1.0 - r2, 0.0, 1.0))
        img = 0.8 * (1.0 - 0.6 * (1.0 - mu))
        tex = np.zeros_like(img)
        for (by, bs, ba) in self.belts:
            tex += ba * np.exp(-0.5 * ((sy - by) / bs) ** 2)
        for (px, py, ps, pa) in self.spots:
            tex += pa * np.exp(-0.5 * (((sx - px) / ps) ** 2 + ((sy - py) / ps) ** 2))
        img = img * (1.0 + tex)
        # soft limb (1 px) so the edge is not aliased
        edge = np.clip((1.0 - np.sqrt(r2)) * self.R + 0.5, 0.0, 1.0)
        out = background + (img - background) * edge * inside.astype(np.float64) * (edge > 0)
        return out


def jitter_matrix(rng, sigma_t, sigma_rot_deg=0.0, sigma_scale=0.0, clip_t=10.0, center=(0.0, 0.0)):
    """Random similarity (about `center`) close to identity."""
    tx, ty = np.clip(rng.normal(0.0, sigma_t, 2), -clip_t, clip_t)
    a = math.radians(rng.normal(0.0, sigma_rot_deg)) if sigma_rot_deg > 0 else 0.0
    s = rng.normal(1.0, sigma_scale) if sigma_scale > 0 else 1.0
    ca, sa = s * math.cos(a), s * math.sin(a)
    cx, cy = center
    return np.array([[ca, -sa, cx - ca * cx + sa * cy + tx],
                     [sa, ca, cy - sa * cx - ca * cy + ty]], dtype=np.float64)


def make_planet_sequence(width, height, nframes, seed, radius=None, sigma_t=3.0, sigma_rot_deg=0.0,
                         sigma_scale=0.0, blur_sigma=1.2, blur_range=None, noise=0.01, dtype="u16",
                         first_is_reference=True):
    """Returns (frames[list of HxW arrays], matrices[list of 2x3], bpp).

    dtype 'u16' -> uint16 scaled by 65535 (config #1), 'f32' -> float32 in [0,1] (config #2).
    blur_range=(lo,hi): per-frame defocus sigma ~ U(lo,hi) (config #2) instead of the fixed blur_sigma.
    Frame 0 is unjittered when first_is_reference (master = frame 0)."""
    rng = np.random.default_rng(seed)
    radius = radius if radius is not None else min(width, height) * 0.3125
    scene = PlanetScene(width, height, radius, seed)
    frames, mats = [], []
    c = ((width - 1) / 2.0, (height - 1) / 2.0)
    for i in range(nframes):
        if i == 0 and first_is_reference:
            A = np.array([[1.0, 0, 0], [0, 1.0, 0]])
        else:
            A = jitter_matrix(rng, sigma_t, sigma_rot_deg, sigma_scale, center=c)
        img = scene.render(A)
        s = blur_sigma if blur_range is None or (i == 0 and first_is_reference) else rng.uniform(*blur_range)
        if blur_range is not None and i == 0 and first_is_reference:
            s = blur_range[0]
        img = cv2.GaussianBlur(img, (0, 0), s)
        img = img + rng.normal(0.0, noise, img.shape)
        if dtype == "u16":
            frames.append(np.clip(np.rint(img * 65535.0), 0, 65535).astype(np.uint16))
        else:
            frames.append(np.clip(img, 0.0, 1.0).astype(f32))
        mats.append(A)
    return frames, mats, (16 if dtype == "u16" else 32)


def make_bayer_sequence(width, height, nframes, seed, sigma_t=2.0, noise=0.003):
    """Config #3: colour star field + nebula gradient sampled through an RGGB mosaic, uint16."""
    rng = np.random.default_rng(seed)
    nstars = max(50, (width * height) // 20000)
    stars = [(rng.uniform(0, width), rng.uniform(0, height), rng.uniform(1.2, 2.5), rng.uniform(0.05, 0.8),
              rng.uniform(0.5, 1.0), rng.uniform(0.5, 1.0), rng.uniform(0.5, 1.0)) for _ in range(nstars)]
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    frames, shifts = [], []
    for i in range(nframes):
        tx, ty = (0.0, 0.0) if i == 0 else np.clip(rng.normal(0.0, sigma_t, 2), -8, 8)
        sx, sy = x + tx, y + ty
        rgb = [0.03 + 0.05 * (sx / width), 0.03 + 0.04 * (sy / height), 0.04 + 0.03 * ((sx + sy) / (width + height))]
        rgb = [np.array(c) for c in rgb]
        for (px, py, ps, pa, cr, cg, cb) in stars:
            x0, x1 = int(max(0, px - tx - 6 * ps)), int(min(width, px - tx + 6 * ps + 1))
            y0, y1 = int(max(0, py - ty - 6 * ps)), int(min(height, py - ty + 6 * ps + 1))
            if x1 <= x0 or y1 <= y0:
                continue
            g = pa * np.exp(-0.5 * (((sx[y0:y1, x0:x1] - px) / ps) ** 2 + ((sy[y0:y1, x0:x1] - py) / ps) ** 2))
            rgb[0][y0:y1, x0:x1] += cr * g
            rgb[1][y0:y1, x0:x1] += cg * g
            rgb[2][y0:y1, x0:x1] += cb * g
        mosaic = np.empty((height, width), dtype=np.float64)
        mosaic[0::2, 0::2] = rgb[0][0::2, 0::2]   # R
        mosaic[0::2, 1::2] = rgb[1][0::2, 1::2]   # G
        mosaic[1::2, 0::2] = rgb[1][1::2, 0::2]   # G
        mosaic[1::2, 1::2] = rgb[2][1::2, 1::2]   # B
        mosaic += rng.normal(0.0, noise, mosaic.shape)
        frames.append(np.clip(np.rint(mosaic * 65535.0), 0, 65535).astype(np.uint16))
        shifts.append((tx, ty))
    return frames, shifts, 16
