"""CPU: the SER reader of the C ABI (ssk_ser_*, host side, no GPU needed) against files written per the format the reference
reads (core/io/c_ser_file.h:42-56, c_ser_file.cc:272-531): 178-byte header, inverted endianness flag, frames back to back,
optional uint64 time-stamp trailer, the author's bits_per_plane = -32 float extension."""
import struct

import numpy as np
import pytest


def write_ser(path, frames, color_id, bits_per_plane, little_endian=True, timestamps=None):
    h, w = frames[0].shape[:2]
    stored_flag = 0 if little_endian else 1          # c_ser_file.cc:305, 580: the flag is stored inverted
    hdr = b"LUCAM-RECORDER" + struct.pack("<iiiiiii", 0, color_id, stored_flag, w, h, bits_per_plane, len(frames))
    hdr += b"o" * 40 + b"i" * 40 + b"t" * 40 + struct.pack("<QQ", 0, 0)
    assert len(hdr) == 178
    with open(path, "wb") as f:
        f.write(hdr)
        for fr in frames:
            f.write(fr.tobytes() if little_endian else fr.byteswap().tobytes())
        if timestamps is not None:
            ts = np.asarray(timestamps, dtype=np.uint64)
            f.write(ts.tobytes() if little_endian else ts.byteswap().tobytes())


@pytest.mark.parametrize("dtype,bits,color_id", [(np.uint8, 8, 0), (np.uint16, 16, 0), (np.uint16, 12, 8), (np.float32, -32, 0),
                                                 (np.uint8, 8, 100)])
@pytest.mark.parametrize("little", [True, False])
def test_ser_round_trip(tmp_path, dtype, bits, color_id, little):
    from serstacker_b200 import api
    rng = np.random.default_rng(1)
    shape = (24, 40, 3) if color_id == 100 else (24, 40)
    frames = [(rng.random(shape) * (255 if dtype == np.uint8 else 4000 if dtype == np.uint16 else 1)).astype(dtype) for _ in range(5)]
    ts = [1000 + 7 * i for i in range(5)]
    p = str(tmp_path / "a.ser")
    write_ser(p, frames, color_id, bits, little, ts)
    r = api.c_ser_reader(p)
    assert (r.cols, r.rows, r.num_frames, r.color_id, r.bits_per_plane) == (40, 24, 5, color_id, bits)
    assert r.has_timestamps
    for i in (3, 0, 4):                                # seek order does not matter
        img, t = r.read(i)
        assert img.dtype == dtype and np.array_equal(img, frames[i]) and t == ts[i]
    with pytest.raises(Exception):
        r.read(5)


def test_ser_without_timestamps_and_rejects_other_files(tmp_path):
    from serstacker_b200 import api
    frames = [np.full((8, 8), i, np.uint16) for i in range(3)]
    p = str(tmp_path / "b.ser")
    write_ser(p, frames, 0, 16)
    r = api.c_ser_reader(p)
    assert not r.has_timestamps and r.bpp() == 16
    img, t = r.read(2)
    assert t == 0 and np.array_equal(img, frames[2])
    q = str(tmp_path / "c.ser")
    open(q, "wb").write(b"x" * 400)
    with pytest.raises(Exception):
        api.c_ser_reader(q)
