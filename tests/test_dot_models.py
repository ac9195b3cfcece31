"""CPU: how far the reference's own answer moves with the arithmetic of cv::Mat::dot.

oracle/ecc.py accumulates the normal-equation sums in double (what IPP's ippsDotProd_32f64f does in the cv2 wheel of this
image); a distribution build of OpenCV without IPP accumulates fp32 FMA lanes in 8192-element blocks (oracle/cvmodel.py
::dot_f32_simd, restated from the published source, not pinnable to a cv2 call).  The spread of the final parameters between
the two models is the envelope inside which "the reference's result" is defined; the GPU path reproduces the double form."""
import numpy as np
import pytest

from oracle import cvmodel
from oracle import ecc as oecc
from oracle import transforms as otf
from serstacker_b200 import synth
from helpers import map_diff_px


def test_simd_dot_model_basics():
    rng = np.random.default_rng(0)
    for n in (1, 7, 63, 64, 100, 8192, 8192 * 3 + 77, 76800):
        a = rng.integers(-8, 9, n).astype(np.float32)
        b = rng.integers(-8, 9, n).astype(np.float32)
        for lanes in (4, 8, 16):
            assert cvmodel.dot_f32_simd(a, b, lanes) == float(np.dot(a.astype(np.float64), b.astype(np.float64)))   # exact on small integers
    a = rng.random(76800).astype(np.float32)
    b = rng.random(76800).astype(np.float32)
    exact = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    for lanes in (4, 8, 16):
        rel = abs(cvmodel.dot_f32_simd(a, b, lanes) - exact) / exact
        assert 0 < rel < 5e-6          # float-lane accumulation: visibly not the double sum, but close


class simd_dot:
    def __init__(self, lanes):
        self.lanes = lanes

    def __enter__(self):
        self.d, self.n = oecc._dot, oecc._norm_l2sqr
        oecc._dot = lambda a, b: cvmodel.dot_f32_simd(a, b, self.lanes)
        return self

    def __exit__(self, *a):
        oecc._dot, oecc._norm_l2sqr = self.d, self.n


CASES = [(0, oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM), (3, oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM), (3, oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL),
         (3, oecc.ECC_ALIGN_LM), (0, oecc.ECC_ALIGN_FORWARD_ADDITIVE), (3, oecc.ECC_ALIGN_FORWARD_ADDITIVE)]


@pytest.mark.parametrize("motion,method", CASES)
def test_parameter_spread_between_dot_models(motion, method, capsys):
    frames, _, _ = synth.make_planet_sequence(320, 240, 4, 11 + motion, sigma_t=2.0, sigma_rot_deg=0.2 if motion else 0.0,
                                              sigma_scale=0.002 if motion else 0.0, dtype="f32")
    kw = dict(maxlevel=-1, minimum_image_size=16, epsx=0.05, max_iterations=30, update_step_scale=1.0)

    def run():
        t = otf.create_image_transform(motion)
        e = oecc.EccH(t, method=method, **kw)
        e.set_reference_image(frames[0], None)
        out = []
        for f in frames[1:]:
            t.reset()
            e.align(f, None)
            out.append((t.parameters().copy(), e.num_iterations))
        return out

    base = run()
    spread, its = 0.0, 0
    for lanes in (8, 16):
        with simd_dot(lanes):
            alt = run()
        for (p0, n0), (p1, n1) in zip(base, alt):
            spread = max(spread, map_diff_px(motion, p1, p0, (320, 240)))
            its += int(n0 != n1)
    with capsys.disabled():
        print("\n  dot-model spread  motion=%d method=%d: max |d map| = %.3g px, iteration counts differing in %d of %d runs" % (
            motion, method, spread, its, 2 * len(base)))
    # the models must stay within the parity budget where the solver is stable, and within the known envelope elsewhere
    assert spread <= (1e-3 if method == oecc.ECC_ALIGN_INVERSE_COMPOSITIONAL_LM else 0.1)
