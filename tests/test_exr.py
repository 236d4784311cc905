"""gdb200.exr: the OpenEXR layout a MultiFilm writes by default (multifilm.cpp:110-120, bitmap.cpp:3170-3345), checked in
both directions against the OpenEXR library bundled with OpenCV (an independent implementation of the format)."""
import os

import numpy as np
import pytest

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"

import gdb200
from gdb200 import exr, pfm


def _image(h=37, w=53, seed=0):
    rng = np.random.default_rng(seed)
    img = (rng.random((h, w, 3)) ** 3 * 10).astype(np.float32)
    img[min(5, h - 1), min(7, w - 1)] = (0.0, 1e-8, 60000.0)
    img[0, :8, 0] = 0.25                                     # a flat run (compresses)
    return img


def _cv2():
    cv2 = pytest.importorskip("cv2")
    if "OpenEXR" not in cv2.getBuildInformation():
        pytest.skip("OpenCV without OpenEXR")
    return cv2


@pytest.mark.parametrize("component", ["float16", "float32"])
@pytest.mark.parametrize("compression", ["zip", "zips", "none"])
def test_written_files_are_read_by_openexr(tmp_path, component, compression):
    cv2 = _cv2()
    img = _image()
    path = str(tmp_path / "a.exr")
    exr.write_exr(path, img, component, compression)
    got = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert got is not None and got.shape == img.shape
    expect = img.astype(np.float16).astype(np.float32) if component == "float16" else img
    assert np.array_equal(got[..., ::-1], expect)
    assert np.array_equal(exr.read_exr(path), expect)


@pytest.mark.parametrize("name", ["ZIP", "ZIPS", "NO", "RLE"])
@pytest.mark.parametrize("half", [False, True])
def test_files_written_by_openexr_are_read(tmp_path, name, half):
    cv2 = _cv2()
    img = _image(41, 33, seed=2)
    path = str(tmp_path / "b.exr")
    flags = [cv2.IMWRITE_EXR_COMPRESSION, getattr(cv2, "IMWRITE_EXR_COMPRESSION_" + name),
             cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF if half else cv2.IMWRITE_EXR_TYPE_FLOAT]
    assert cv2.imwrite(path, np.ascontiguousarray(img[..., ::-1]), flags)
    expect = img.astype(np.float16).astype(np.float32) if half else img
    assert np.array_equal(exr.read_exr(path), expect)
    planes, attrs = exr.read_exr_channels(path)
    assert sorted(planes) == ["B", "G", "R"] and attrs["compression"][0] == "compression"


def test_unsupported_files_fail_loudly(tmp_path):
    cv2 = _cv2()
    path = str(tmp_path / "p.exr")
    cv2.imwrite(path, _image()[..., ::-1].copy(), [cv2.IMWRITE_EXR_COMPRESSION, cv2.IMWRITE_EXR_COMPRESSION_PIZ])
    with pytest.raises(exr.ExrError, match="not supported"):
        exr.read_exr(path)
    (tmp_path / "n.exr").write_bytes(b"PF\n1 1\n-1\n0000")
    with pytest.raises(exr.ExrError, match="not an OpenEXR"):
        exr.read_exr(str(tmp_path / "n.exr"))


def test_header_is_what_the_reference_writes(tmp_path):
    """Imf::Header(w, h) defaults + bitmap.cpp:3192-3238: ZIP, increasing y, R/G/B HALF, chromaticities, generatedBy."""
    path = str(tmp_path / "h.exr")
    exr.write_exr(path, _image(20, 10), metadata={"spp": 64, "alpha": 0.2})
    planes, attrs = exr.read_exr_channels(path)
    assert attrs["compression"][1] == bytes([exr.ZIP_COMPRESSION]) and attrs["lineOrder"][1] == b"\0"
    assert all(planes[c].dtype == np.float16 and planes[c].shape == (20, 10) for c in "RGB")
    assert attrs["chromaticities"][0] == "chromaticities" and len(attrs["chromaticities"][1]) == 32
    assert attrs["generatedBy"] == ("string", b"gdb200") and attrs["spp"][0] == "int" and attrs["alpha"][0] == "float"
    assert np.frombuffer(attrs["dataWindow"][1], "<i4").tolist() == [0, 0, 9, 19]
    one = str(tmp_path / "y.exr")
    exr.write_exr(one, _image(4, 4)[..., 0], "float32", "none")
    assert sorted(exr.read_exr_channels(one)[0]) == ["Y"] and exr.read_exr(one).shape == (4, 4, 3)


def test_multifilm_on_disk_in_both_formats(tmp_path):
    bufs = {n: _image(12, 16, seed=i) for i, n in enumerate(pfm.BUFFER_NAMES)}
    integ = gdb200.GPTIntegrator()
    paths = integ.save(str(tmp_path / "render.png"), bufs, "openexr", "float32")
    assert [os.path.basename(p) for p in paths] == ["render" + n + ".exr" for n in pfm.BUFFER_NAMES]
    back = pfm.load_multifilm(str(tmp_path / "render"))
    assert all(np.array_equal(back[n], bufs[n]) for n in bufs)
    paths = integ.save(str(tmp_path / "r2"), bufs)                               # default stays PFM
    assert all(p.endswith(".pfm") for p in paths)
    half = integ.save(str(tmp_path / "r3"), bufs, "openexr")
    assert np.array_equal(exr.read_exr(half[0]), bufs["-final"].astype(np.float16).astype(np.float32))
    with pytest.raises(ValueError):
        integ.save(str(tmp_path / "r4"), bufs, "rgbe")
