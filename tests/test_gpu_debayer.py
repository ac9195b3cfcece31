"""GPU parity: debayer_nn2 (core/io/debayer.cc:827-1195) against oracle/debayer.py.  Bit-exact for every depth."""
import numpy as np
import pytest

from oracle import debayer as od

pytestmark = pytest.mark.gpu


def _raw(rng, shape, dtype):
    if dtype == np.float32:
        return rng.random(shape, dtype=np.float32)
    return rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("colorid", [8, 9, 10, 11])
@pytest.mark.parametrize("shape", [(2, 2), (2, 6), (6, 2), (10, 14), (250, 334), (480, 640)])
def test_debayer_nn2_matches_oracle(gpu, dtype, colorid, shape):
    from serstacker_b200 import api
    rng = np.random.default_rng(shape[0] + colorid)
    raw = _raw(rng, shape, dtype)
    assert np.array_equal(api.debayer_nn2(raw, colorid), od.debayer_nn2(raw, colorid))


def test_debayer_nn2_config3_size_and_saturated_values(gpu):
    """Config #3's frame: 4096x3000 RGGB16, including full-scale samples (the integer averages must not wrap)."""
    from serstacker_b200 import api
    rng = np.random.default_rng(0)
    raw = _raw(rng, (3000, 4096), np.uint16)
    raw[100:200, 100:300] = 65535
    assert np.array_equal(api.debayer_nn2(raw, 8), od.debayer_nn2(raw, 8))


def test_debayer_nn2_rejects_uneven_sizes_and_unknown_patterns(gpu):
    from serstacker_b200 import api
    with pytest.raises(Exception):
        api.debayer_nn2(np.zeros((5, 6), np.uint16), 8)
    with pytest.raises(Exception):
        api.debayer_nn2(np.zeros((6, 6), np.uint16), 3)
