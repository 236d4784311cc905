"""OpenEXR files of a MultiFilm render (the film's default fileFormat, multifilm.cpp:110-120), without the OpenEXR library.

What the reference writes (Bitmap::writeOpenEXR, src/libcore/bitmap.cpp:3170-3345, through `Imf::Header(w, h)`): a
single-part scan-line file, channels R, G, B of type HALF (componentFormat "float16", the default) or FLOAT ("float32"),
ZIP compression (16 scan lines per chunk -- the `Imf::Header` default), increasing-y line order, a `chromaticities`
attribute with the Rec. 709 primaries and a `generatedBy` string.  `write_exr` produces that layout; `read_exr` reads
scan-line files with NO / RLE / ZIPS / ZIP compression and HALF / FLOAT / UINT channels, which covers the reference's own
output, so `tools/reconstruct.py` can re-run the reconstruction on images rendered by the reference.

The container (magic 20000630, version 2, attribute list, chunk offset table, per-chunk "y, size, data"; inside a chunk
every scan line stores its channels one after the other in alphabetical channel order) and the ZIP pre-processing (bytes
de-interleaved into two halves, then delta-encoded with bias 128) are those of the published OpenEXR file layout; the
tests check both directions against the OpenEXR library that OpenCV bundles.
"""
import os
import struct
import zlib

import numpy as np

MAGIC = 20000630
NO_COMPRESSION, RLE_COMPRESSION, ZIPS_COMPRESSION, ZIP_COMPRESSION, PIZ_COMPRESSION = 0, 1, 2, 3, 4
_LINES = {NO_COMPRESSION: 1, RLE_COMPRESSION: 1, ZIPS_COMPRESSION: 1, ZIP_COMPRESSION: 16}
_COMPRESSION_NAMES = {"none": NO_COMPRESSION, "zips": ZIPS_COMPRESSION, "zip": ZIP_COMPRESSION}
UINT, HALF, FLOAT = 0, 1, 2
_DTYPES = {UINT: np.dtype("<u4"), HALF: np.dtype("<f2"), FLOAT: np.dtype("<f4")}
BUFFER_NAMES = ("-final", "-throughput", "-dx", "-dy", "-direct")


class ExrError(RuntimeError):
    pass


def _attr(name, typ, value):
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(value)) + value


def _zip_pack(raw):
    a = np.frombuffer(raw, dtype=np.uint8)
    t = np.concatenate([a[0::2], a[1::2]])                                          # even bytes, then odd bytes
    d = t.astype(np.int16)
    d[1:] = d[1:] - t[:-1].astype(np.int16) + (128 + 256)
    packed = zlib.compress(d.astype(np.uint8).tobytes())
    return packed if len(packed) < len(raw) else raw                                 # stored raw when compression does not help


def _zip_unpack(data, raw_size):
    if len(data) == raw_size:
        return data
    t = np.frombuffer(zlib.decompress(data), dtype=np.uint8)
    if t.size != raw_size:
        raise ExrError("EXR chunk inflates to an unexpected size")
    t = (np.cumsum(t.astype(np.int64) - 128) + 128).astype(np.uint8)                 # undo t[i] = t[i] - t[i-1] + 128 (mod 256)
    half = (raw_size + 1) // 2
    out = np.empty(raw_size, dtype=np.uint8)
    out[0::2] = t[:half]
    out[1::2] = t[half:]
    return out.tobytes()


def _rle_unpack(data, raw_size):
    if len(data) == raw_size:
        return data
    out, i = bytearray(), 0
    while i < len(data):
        n = struct.unpack_from("b", data, i)[0]
        i += 1
        if n < 0:
            out += data[i:i - n]
            i += -n
        else:
            out += data[i:i + 1] * (n + 1)
            i += 1
    t = np.frombuffer(bytes(out), dtype=np.uint8)
    if t.size != raw_size:
        raise ExrError("EXR RLE chunk expands to an unexpected size")
    t = (np.cumsum(t.astype(np.int64) - 128) + 128).astype(np.uint8)
    half = (raw_size + 1) // 2
    res = np.empty(raw_size, dtype=np.uint8)
    res[0::2] = t[:half]
    res[1::2] = t[half:]
    return res.tobytes()


def write_exr(path, image, component_format="float16", compression="zip", channel_names=None, metadata=None):
    """image: [h, w, C] (or [h, w]) array, top-left origin.  Channels default to R, G, B (Y for one channel).  metadata:
    {name: str | int | float} written as string / int / float attributes like bitmap.cpp:3198-3228."""
    a = np.asarray(image)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    if component_format not in ("float16", "float32", "uint32"):
        raise ExrError("writeOpenEXR(): Invalid component type (must be float16, float32, or uint32)")
    ptype = {"float16": HALF, "float32": FLOAT, "uint32": UINT}[component_format]
    if channel_names is None:
        channel_names = {1: ("Y",), 3: ("R", "G", "B"), 4: ("R", "G", "B", "A")}.get(c)
    if channel_names is None or len(channel_names) != c:
        raise ExrError("writeOpenEXR(): channel names do not match the channel count")
    if compression not in _COMPRESSION_NAMES:
        raise ExrError(f"write_exr: unsupported compression \"{compression}\"")
    comp = _COMPRESSION_NAMES[compression]
    with np.errstate(over="ignore"):
        planes = {n: np.ascontiguousarray(a[:, :, i]).astype(_DTYPES[ptype]) for i, n in enumerate(channel_names)}
    order = sorted(channel_names)                                                    # the channel list is kept sorted by name

    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iBBBBii", ptype, 0, 0, 0, 0, 1, 1) for n in order) + b"\0"
    head = struct.pack("<iI", MAGIC, 2)
    attrs = {"channels": ("chlist", chlist),
             "compression": ("compression", struct.pack("<B", comp)),
             "dataWindow": ("box2i", struct.pack("<4i", 0, 0, w - 1, h - 1)),
             "displayWindow": ("box2i", struct.pack("<4i", 0, 0, w - 1, h - 1)),
             "lineOrder": ("lineOrder", struct.pack("<B", 0)),
             "pixelAspectRatio": ("float", struct.pack("<f", 1.0)),
             "screenWindowCenter": ("v2f", struct.pack("<2f", 0.0, 0.0)),
             "screenWindowWidth": ("float", struct.pack("<f", 1.0))}
    if set(channel_names) >= {"R", "G", "B"}:                                         # Imf::addChromaticities(header, Chromaticities()), bitmap.cpp:3236-3238
        attrs["chromaticities"] = ("chromaticities", struct.pack("<8f", 0.64, 0.33, 0.30, 0.60, 0.15, 0.06, 0.3127, 0.3290))
    meta = {"generatedBy": "gdb200"}
    meta.update(metadata or {})
    for k, v in meta.items():
        if isinstance(v, bool) or isinstance(v, int):
            attrs[k] = ("int", struct.pack("<i", int(v)))
        elif isinstance(v, float):
            attrs[k] = ("float", struct.pack("<f", v))
        else:
            attrs[k] = ("string", str(v).encode())
    head += b"".join(_attr(k, *attrs[k]) for k in sorted(attrs)) + b"\0"

    lines = _LINES[comp]
    chunks = []
    for y0 in range(0, h, lines):
        raw = b"".join(planes[n][y].tobytes() for y in range(y0, min(y0 + lines, h)) for n in order)
        data = raw if comp == NO_COMPRESSION else _zip_pack(raw)
        chunks.append(struct.pack("<ii", y0, len(data)) + data)
    table_at = len(head)
    offsets, pos = [], table_at + 8 * len(chunks)
    for ch in chunks:
        offsets.append(pos)
        pos += len(ch)
    with open(path, "wb") as f:
        f.write(head + struct.pack("<%dQ" % len(offsets), *offsets) + b"".join(chunks))


def read_exr_channels(path):
    """Returns ({channel name: [h, w] array in the file's own type}, {attribute name: (type, raw bytes)})."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 8 or struct.unpack_from("<i", data, 0)[0] != MAGIC:
        raise ExrError(f"\"{path}\": not an OpenEXR file")
    version = struct.unpack_from("<I", data, 4)[0]
    if (version & 0xFF) != 2 or (version & 0x1A00):                                   # tiled (0x200), non-image (0x800), multi-part (0x1000)
        raise ExrError(f"\"{path}\": only single-part scan-line OpenEXR files are supported")
    pos, attrs = 8, {}
    while data[pos] != 0:
        e = data.index(b"\0", pos)
        name = data[pos:e].decode()
        e2 = data.index(b"\0", e + 1)
        typ = data[e + 1:e2].decode()
        size = struct.unpack_from("<i", data, e2 + 1)[0]
        attrs[name] = (typ, data[e2 + 5:e2 + 5 + size])
        pos = e2 + 5 + size
    pos += 1
    for need in ("channels", "compression", "dataWindow"):
        if need not in attrs:
            raise ExrError(f"\"{path}\": OpenEXR header without '{need}'")
    channels, cl, p = [], attrs["channels"][1], 0
    while cl[p] != 0:
        e = cl.index(b"\0", p)
        ptype, _, xs, ys = struct.unpack_from("<iI2i", cl, e + 1)
        if xs != 1 or ys != 1:
            raise ExrError(f"\"{path}\": sub-sampled channels are not supported")
        channels.append((cl[p:e].decode(), ptype))
        p = e + 17
    comp = attrs["compression"][1][0]
    if comp not in _LINES:
        raise ExrError(f"\"{path}\": compression type {comp} (PIZ/PXR24/B44/DWA) is not supported; re-save as ZIP")
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    lines = _LINES[comp]
    n_chunks = (h + lines - 1) // lines
    offsets = struct.unpack_from("<%dQ" % n_chunks, data, pos)
    line_bytes = sum(_DTYPES[t].itemsize for _, t in channels) * w
    planes = {n: np.empty((h, w), dtype=_DTYPES[t]) for n, t in channels}
    for off in offsets:
        cy, size = struct.unpack_from("<ii", data, off)
        rows = min(lines, y1 - cy + 1)
        raw_size = rows * line_bytes
        blob = data[off + 8:off + 8 + size]
        if comp in (ZIPS_COMPRESSION, ZIP_COMPRESSION):
            blob = _zip_unpack(blob, raw_size)
        elif comp == RLE_COMPRESSION:
            blob = _rle_unpack(blob, raw_size)
        if len(blob) != raw_size:
            raise ExrError(f"\"{path}\": truncated OpenEXR chunk")
        q = 0
        for r in range(rows):
            for n, t in channels:
                nb = _DTYPES[t].itemsize * w
                planes[n][cy - y0 + r] = np.frombuffer(blob, dtype=_DTYPES[t], count=w, offset=q)
                q += nb
    return planes, attrs


def read_exr(path):
    """Returns a float32 [h, w, 3] RGB image (top-left origin); a luminance-only file is replicated into three channels."""
    planes, _ = read_exr_channels(path)
    if all(k in planes for k in "RGB"):
        return np.stack([planes[k].astype(np.float32) for k in "RGB"], -1)
    if "Y" in planes:
        return np.repeat(planes["Y"].astype(np.float32)[:, :, None], 3, axis=2)
    raise ExrError(f"\"{path}\": no R,G,B or Y channels (has {sorted(planes)})")


def save_multifilm(dest, buffers, component_format="float16"):
    """MultiFilm::develop for the default fileFormat "openexr" (multifilm.cpp:423-481): "<dest><buffer name>.exr" for the
    five G-PT buffers, RGB, float16 unless componentFormat says float32."""
    root = os.path.splitext(dest)[0] if os.path.splitext(dest)[1].lower() in (".pfm", ".exr", ".rgbe", ".png") else dest
    paths = []
    for name in BUFFER_NAMES:
        if name in buffers and buffers[name] is not None:
            paths.append(root + name + ".exr")
            write_exr(paths[-1], buffers[name], component_format=component_format)
    return paths
