"""PFM interchange: Bitmap::writePFM / readPFM semantics (bitmap.cpp:3745-3850) and the
file set MultiFilm::develop writes for a gpt render (multifilm.cpp:423-516)."""
import os
import struct

import numpy as np
import pytest

from gdb200 import pfm


def test_header_and_bottom_up_scanlines(tmp_path):
    img = np.arange(2 * 3 * 3, dtype=np.float32).reshape(2, 3, 3)
    p = tmp_path / "a.pfm"
    pfm.write_pfm(p, img)
    raw = p.read_bytes()
    assert raw.startswith(b"PF\n3 2\n-1\n")
    body = np.frombuffer(raw[len(b"PF\n3 2\n-1\n"):], dtype="<f4").reshape(2, 3, 3)
    assert np.array_equal(body[0], img[1]) and np.array_equal(body[1], img[0])     # first scanline in the file = last image row
    assert np.array_equal(pfm.read_pfm(p), img)


def test_reader_honours_byte_order_and_scale(tmp_path):
    img = np.array([[[1.0, 2.0, 3.0]], [[4.0, 5.0, 6.0]]], dtype=np.float32)        # 2 rows x 1 column
    p = tmp_path / "be.pfm"
    with open(p, "wb") as f:
        f.write(b"PF\n1 2\n2.0\n")                                                  # positive: big endian, scale 2
        for row in img[::-1]:
            f.write(struct.pack(">3f", *row[0]))
    assert np.array_equal(pfm.read_pfm(p), img * 2)
    g = tmp_path / "g.pfm"
    pfm.write_pfm(g, img[:, :, 0])
    assert g.read_bytes().startswith(b"Pf\n1 2\n-1\n")
    assert np.array_equal(pfm.read_pfm(g), img[:, :, 0])


def test_errors(tmp_path):
    p = tmp_path / "bad.pfm"
    p.write_bytes(b"P6\n1 1\n255\n\0\0\0")
    with pytest.raises(ValueError, match="Invalid PFM header"):
        pfm.read_pfm(p)
    p.write_bytes(b"PF\nx 1\n-1\n")
    with pytest.raises(ValueError, match="dimensions"):
        pfm.read_pfm(p)
    p.write_bytes(b"PF\n2 2\n-1\n\0\0\0\0")
    with pytest.raises(ValueError, match="truncated"):
        pfm.read_pfm(p)
    with pytest.raises(ValueError, match="pixel format"):
        pfm.write_pfm(p, np.zeros((2, 2, 4)))


def test_multifilm_file_set(tmp_path):
    rng = np.random.default_rng(0)
    bufs = {n: rng.standard_normal((5, 7, 3)) for n in pfm.BUFFER_NAMES}
    paths = pfm.save_multifilm(str(tmp_path / "render.exr"), bufs)                   # the extension is replaced
    assert [os.path.basename(x) for x in paths] == ["render-final.pfm", "render-throughput.pfm", "render-dx.pfm",
                                                    "render-dy.pfm", "render-direct.pfm"]
    back = pfm.load_multifilm(str(tmp_path / "render"))
    for n in pfm.BUFFER_NAMES:
        assert np.array_equal(back[n], bufs[n].astype(np.float32))                   # Float -> float32 like multifilm's "pfm" format
