"""PFM files as the reference reads and writes them, and the on-disk layout of a MultiFilm render.

* Bitmap::writePFM / readPFM (src/libcore/bitmap.cpp:3745-3812, 3814-3850): header "PF\\n<w> <h>\\n-1\\n" (little-endian
  host; 'Pf' for one channel), float32 scanlines stored BOTTOM-UP; a reader honours the sign of the third header token
  as byte order and its magnitude as a scale factor, and flips the image back to top-down.
* MultiFilm::develop (src/films/multifilm.cpp:423-516): one file per buffer named "<dest><buffer name>.pfm" with the
  buffer names of gpt.cpp:1380 ("-final", "-throughput", "-dx", "-dy", "-direct"); fileFormat "pfm" forces RGB float32
  (multifilm.cpp:222-235).  These are the files README.txt:68-74 suggests feeding to an offline reconstruction.
"""
import os

import numpy as np

BUFFER_NAMES = ("-final", "-throughput", "-dx", "-dy", "-direct")


def write_pfm(path, image):
    """image: [h, w, 3] or [h, w] array (top-left origin, like every buffer of this package)."""
    a = np.asarray(image, dtype=np.float32)
    if a.ndim == 3 and a.shape[2] == 1:
        a = a[:, :, 0]
    if a.ndim not in (2, 3) or (a.ndim == 3 and a.shape[2] != 3):
        raise ValueError("writePFM(): pixel format must be RGB or luminance")
    h, w = a.shape[:2]
    header = f"P{'F' if a.ndim == 3 else 'f'}\n{w} {h}\n-1\n".encode("ascii")
    with open(path, "wb") as f:
        f.write(header)
        f.write(np.ascontiguousarray(a[::-1]).astype("<f4").tobytes())      # scanline y of the file = row h-1-y


def _token(f):
    out = b""
    while True:
        c = f.read(1)
        if not c:
            raise ValueError("Unexpected end of PFM header")
        if c.isspace():
            if out:
                return out.decode("ascii")
            continue
        out += c


def read_pfm(path):
    """Returns a float32 array [h, w, 3] (or [h, w]) with top-left origin, scale applied."""
    with open(path, "rb") as f:
        magic = f.read(2)
        if magic not in (b"PF", b"Pf"):
            raise ValueError("Invalid PFM header!")
        color = magic == b"PF"
        try:
            w, h = int(_token(f)), int(_token(f))
        except ValueError:
            raise ValueError("Could not parse image dimensions!")
        try:
            scale_and_order = float(_token(f))
        except ValueError:
            raise ValueError("Could not parse scale/order information!")
        n = w * h * (3 if color else 1)
        data = np.frombuffer(f.read(4 * n), dtype="<f4" if scale_and_order <= 0 else ">f4")
        if data.size != n:
            raise ValueError("PFM file is truncated")
    data = data.astype(np.float32)
    scale = abs(scale_and_order)
    if scale != 1:
        data = data * np.float32(scale)
    img = data.reshape((h, w, 3) if color else (h, w))
    return np.ascontiguousarray(img[::-1])


def save_multifilm(dest, buffers):
    """Writes the five G-PT buffers like MultiFilm::develop does for fileFormat 'pfm'.  `dest` may carry any
    extension (it is replaced, multifilm.cpp:466-468).  Returns the paths written, in buffer order."""
    root = os.path.splitext(dest)[0] if os.path.splitext(dest)[1].lower() in (".pfm", ".exr", ".rgbe", ".png") else dest
    paths = []
    for name in BUFFER_NAMES:
        if name in buffers and buffers[name] is not None:
            paths.append(root + name + ".pfm")
            write_pfm(paths[-1], buffers[name])
    return paths


def load_multifilm(dest):
    """The inverse of save_multifilm: {buffer name: float32 [h, w, 3]} for the files that exist -- ".pfm", or the ".exr"
    files a MultiFilm writes by default (gdb200.exr)."""
    out = {}
    for name in BUFFER_NAMES:
        p = dest + name + ".pfm"
        if os.path.exists(p):
            out[name] = read_pfm(p)
        elif os.path.exists(dest + name + ".exr"):
            from .exr import read_exr
            out[name] = read_exr(dest + name + ".exr")
    return out
