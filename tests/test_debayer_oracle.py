"""CPU: the vectorised debayer_nn2 restatement (oracle/debayer.py) against the loop-by-loop transcription of the
reference's RGGB case (core/io/debayer.cc:856-922), and pattern symmetries that tie the other three patterns to it."""
import numpy as np
import pytest

from oracle import debayer as od


def _raw(rng, shape, dtype):
    if dtype == np.float32:
        return rng.random(shape, dtype=np.float32)
    return rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32])
@pytest.mark.parametrize("shape", [(2, 2), (2, 6), (6, 2), (4, 4), (10, 14), (16, 8)])
def test_vectorised_rggb_equals_literal_transcription(dtype, shape):
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    raw = _raw(rng, shape, dtype)
    assert np.array_equal(od.debayer_nn2(raw, od.COLORID_BAYER_RGGB), od.debayer_nn2_rggb_literal(raw))


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_own_colour_samples_pass_through_and_patterns_shift(dtype):
    rng = np.random.default_rng(3)
    raw = _raw(rng, (12, 16), dtype)
    for cid, (ry, rx) in od._RED_AT.items():
        out = od.debayer_nn2(raw, cid)
        assert np.array_equal(out[ry::2, rx::2, 2], raw[ry::2, rx::2])                   # R samples
        assert np.array_equal(out[1 - ry::2, 1 - rx::2, 0], raw[1 - ry::2, 1 - rx::2])   # B samples
        assert np.array_equal(out[ry::2, 1 - rx::2, 1], raw[ry::2, 1 - rx::2])           # G samples
    # BGGR is RGGB with red and blue exchanged
    a = od.debayer_nn2(raw, od.COLORID_BAYER_RGGB)
    b = od.debayer_nn2(raw, od.COLORID_BAYER_BGGR)
    assert np.array_equal(a[..., ::-1], b)


def test_uneven_size_is_rejected():
    with pytest.raises(ValueError):
        od.debayer_nn2(np.zeros((5, 6), np.uint16), od.COLORID_BAYER_RGGB)
