"""Triangle-mesh files of the reference's shape plugins, read into the arrays `SceneBuilder.mesh` takes.

* `serialized` (src/shapes/serialized.cpp:78-146 format description; TriMesh::loadCompressed, trimesh.cpp:175-252;
  the end-of-file dictionary, trimesh.cpp:272-293) -- Mitsuba's own zlib-compressed mesh container, versions 3 and 4,
  single or double precision, several meshes per file (`shapeIndex`).  `save_serialized` writes version-4 files.
* `ply` (src/shapes/ply.cpp): ascii / binary_little_endian / binary_big_endian, `vertex` element with x y z [nx ny nz],
  `face` element with a `vertex_indices` / `vertex_index` list of 3 or 4 entries (quads split as (0,1,2),(3,0,2),
  ply.cpp:299-312).  Texture coordinates and colours are parsed past (nothing on the hot path reads them).

What happens to the normals afterwards is TriMesh::computeNormals (trimesh.cpp:608-681): `finish_mesh`.
"""
import math
import struct
import zlib

import numpy as np

FILEFORMAT_HEADER, VERSION_V3, VERSION_V4 = 0x041C, 0x0003, 0x0004                  # trimesh.cpp:34-36
HAS_NORMALS, HAS_TEXCOORDS, HAS_COLORS, FACE_NORMALS = 0x0001, 0x0002, 0x0008, 0x0010   # trimesh.cpp:89-97
SINGLE_PRECISION, DOUBLE_PRECISION = 0x1000, 0x2000


class MeshError(RuntimeError):
    pass


def unit_angle(u, v):
    """unitAngle (util.h:305-310) on rows of unit vectors."""
    d = np.einsum("ij,ij->i", u, v)
    neg = math.pi - 2 * np.arcsin(np.minimum(1.0, 0.5 * np.linalg.norm(v + u, axis=1)))
    pos = 2 * np.arcsin(np.minimum(1.0, 0.5 * np.linalg.norm(v - u, axis=1)))
    return np.where(d < 0, neg, pos)


def compute_normals(verts, tris, flip=False):
    """The generated-normals branch of TriMesh::computeNormals (trimesh.cpp:631-672): angle-weighted face normals summed
    per vertex in triangle order; a degenerate triangle contributes nothing; untouched vertices get (1, 0, 0)."""
    P = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    T = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
    N = np.zeros_like(P)
    if len(T):
        fn = np.cross(P[T[:, 1]] - P[T[:, 0]], P[T[:, 2]] - P[T[:, 0]])
        length = np.linalg.norm(fn, axis=1)
        ok = length != 0
        fn = fn / np.where(ok, length, 1.0)[:, None]
        contrib = np.zeros((len(T), 3, 3))
        with np.errstate(invalid="ignore", divide="ignore"):
            for i in range(3):
                a = P[T[:, (i + 1) % 3]] - P[T[:, i]]
                b = P[T[:, (i + 2) % 3]] - P[T[:, i]]
                ang = unit_angle(a / np.linalg.norm(a, axis=1, keepdims=True), b / np.linalg.norm(b, axis=1, keepdims=True))
                contrib[:, i] = fn * ang[:, None]
        contrib[~ok] = 0
        np.add.at(N, T.reshape(-1), contrib.reshape(-1, 3))                         # sequential: triangle order, corner order
    length = np.linalg.norm(N, axis=1)
    if flip:
        length = -length
    bad = length == 0
    N = N / np.where(bad, 1.0, length)[:, None]
    N[bad] = (1.0, 0.0, 0.0)
    return N


def finish_mesh(verts, tris, normals, face_normals=False, flip_normals=False):
    """TriMesh::computeNormals as configure() calls it (trimesh.cpp:608-681): faceNormals drops stored normals (flipNormals
    then swaps the winding), stored normals are kept (negated by flipNormals), otherwise smooth normals are generated.
    Returns (verts [V,3], tris [T,3] int, normals [V,3] or None)."""
    P = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    T = np.asarray(tris, dtype=np.int64).reshape(-1, 3).copy()
    if face_normals:
        if flip_normals:
            T[:, [0, 1]] = T[:, [1, 0]]
        return P, T, None
    if normals is not None:
        N = np.asarray(normals, dtype=np.float64).reshape(-1, 3)
        return P, T, (-N if flip_normals else N)
    return P, T, compute_normals(P, T, flip=flip_normals)


def apply_to_world(verts, normals, tris, to_world, flip_winding_on_mirror):
    """objectToWorld on points and normals (serialized.cpp:184-202, ply.cpp:227-233): normals go through the inverse
    transpose and are renormalised; `serialized` (not `ply`) also swaps the winding when the transform mirrors."""
    P = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    T = np.asarray(tris, dtype=np.int64).reshape(-1, 3).copy()
    N = None if normals is None else np.asarray(normals, dtype=np.float64).reshape(-1, 3)
    if to_world is None or np.array_equal(np.asarray(to_world), np.eye(4)):
        return P, N, T
    m = np.asarray(to_world, dtype=np.float64).reshape(4, 4)
    P = P @ m[:3, :3].T + m[:3, 3]
    if N is not None:
        N = N @ np.linalg.inv(m[:3, :3])                                             # rows times inv(M) = (inv(M)^T n)^T
        N = N / np.linalg.norm(N, axis=1, keepdims=True)
    if flip_winding_on_mirror and np.linalg.det(m[:3, :3]) < 0:
        T[:, [0, 1]] = T[:, [1, 0]]
    return P, N, T


# ------------------------------------------------------------------ serialized
def _mesh_offset(data, version, index):
    """TriMesh::readOffset (trimesh.cpp:272-293): the dictionary at the end of the file."""
    count = struct.unpack_from("<I", data, len(data) - 4)[0]
    if index < 0 or index > count:                                                   # the reference's own bound (idx > count)
        raise MeshError(f"Unable to unserialize mesh, shape index is out of range! (requested {index} out of 0..{count - 1})")
    if version == VERSION_V4:
        return struct.unpack_from("<Q", data, len(data) - 8 * (count - index) - 4)[0]
    return struct.unpack_from("<I", data, len(data) - 4 * (count - index + 1))[0]


def load_serialized(path, shape_index=0, to_world=None, face_normals=False, flip_normals=False):
    """SerializedMesh (serialized.cpp:148-209).  The plugin's faceNormals parameter replaces the file's flag
    (serialized.cpp:179).  Returns (verts, tris, normals-or-None) ready for SceneBuilder.mesh."""
    if shape_index < 0:
        raise MeshError("Shape index must be nonnegative!")
    with open(path, "rb") as f:
        data = f.read()
    fmt, version = struct.unpack_from("<HH", data, 0)
    if fmt == 0x1C04:
        raise MeshError("Encountered a geometry file generated by an old version of Mitsuba. Please re-import the scene to "
                        "update this file to the current format.")
    if fmt != FILEFORMAT_HEADER:
        raise MeshError("Encountered an invalid file format!")
    if version not in (VERSION_V3, VERSION_V4):
        raise MeshError("Encountered an incompatible file version!")
    start = 4
    if shape_index != 0:
        start = _mesh_offset(data, version, shape_index) + 4                          # seek(offset), skip the 2-short header
    raw = zlib.decompressobj().decompress(data[start:])
    flags = struct.unpack_from("<I", raw, 0)[0]
    pos = 4
    if version == VERSION_V4:
        end = raw.index(b"\0", pos)
        pos = end + 1
    n_vert, n_tri = struct.unpack_from("<QQ", raw, pos)
    pos += 16
    dt = np.dtype("<f8") if flags & DOUBLE_PRECISION else np.dtype("<f4")

    def take(count, width):
        nonlocal pos
        arr = np.frombuffer(raw, dtype=dt, count=count * width, offset=pos).astype(np.float64).reshape(count, width)
        pos += count * width * dt.itemsize
        return arr
    verts = take(n_vert, 3)
    normals = take(n_vert, 3) if flags & HAS_NORMALS else None
    if flags & HAS_TEXCOORDS:
        take(n_vert, 2)
    if flags & HAS_COLORS:
        take(n_vert, 3)
    tris = np.frombuffer(raw, dtype="<u4", count=n_tri * 3, offset=pos).astype(np.int64).reshape(n_tri, 3)
    if len(tris) and tris.max() >= n_vert:
        raise MeshError("serialized: triangle index out of range")
    verts, normals, tris = apply_to_world(verts, normals, tris, to_world, flip_winding_on_mirror=True)
    return finish_mesh(verts, tris, normals, face_normals, flip_normals)


def save_serialized(path, meshes, double_precision=False):
    """Writes a version-4 `.serialized` file (TriMesh::serialize, trimesh.cpp:1131-1175, zlib framing zstream.cpp:29-32, + the
    end-of-file offset dictionary): meshes = [(name, verts, tris, normals-or-None), ...]."""
    out, offsets = bytearray(), []
    dt = "<f8" if double_precision else "<f4"
    for name, verts, tris, normals in meshes:
        offsets.append(len(out))
        flags = (DOUBLE_PRECISION if double_precision else SINGLE_PRECISION) | (HAS_NORMALS if normals is not None else 0)
        V = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
        T = np.asarray(tris, dtype=np.int64).reshape(-1, 3)
        body = struct.pack("<I", flags) + name.encode("utf-8") + b"\0" + struct.pack("<QQ", len(V), len(T))
        body += V.astype(dt).tobytes()
        if normals is not None:
            body += np.asarray(normals, dtype=np.float64).reshape(-1, 3).astype(dt).tobytes()
        body += T.astype("<u4").tobytes()
        out += struct.pack("<HH", FILEFORMAT_HEADER, VERSION_V4) + zlib.compress(body)
    for o in offsets:
        out += struct.pack("<Q", o)
    out += struct.pack("<I", len(offsets))
    with open(path, "wb") as f:
        f.write(bytes(out))


# ------------------------------------------------------------------ ply
_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def load_ply(path, to_world=None, face_normals=False, flip_normals=False):
    """PLYLoader (ply.cpp:88-136, callbacks :203-312).  Positions and normals pass through float32 like the reference's
    `ply::float32` callbacks; no winding swap for mirroring transforms (ply.cpp has none)."""
    with open(path, "rb") as f:
        data = f.read()
    if not data.startswith(b"ply"):
        raise MeshError(f"\"{path}\": not a PLY file")
    end = data.find(b"end_header")
    if end < 0:
        raise MeshError(f"\"{path}\": PLY header without end_header")
    body_at = data.index(b"\n", end) + 1
    fmt, elements = None, []
    for line in data[:end].decode("ascii", "replace").splitlines()[1:]:
        t = line.split()
        if not t or t[0] in ("comment", "obj_info"):
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            elements.append({"name": t[1], "count": int(t[2]), "props": []})
        elif t[0] == "property":
            if not elements:
                raise MeshError(f"\"{path}\": PLY property outside an element")
            if t[1] == "list":
                elements[-1]["props"].append(("list", t[4], _PLY_TYPES[t[2]], _PLY_TYPES[t[3]]))
            else:
                elements[-1]["props"].append(("scalar", t[2], _PLY_TYPES[t[1]]))
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise MeshError(f"\"{path}\": unknown PLY format {fmt}")
    order = ">" if fmt == "binary_big_endian" else "<"
    verts = normals = None
    tris = []
    tokens = data[body_at:].split() if fmt == "ascii" else None
    tpos, bpos = 0, body_at
    for el in elements:
        scalars_only = all(p[0] == "scalar" for p in el["props"])
        if fmt != "ascii" and scalars_only:
            dt = np.dtype([(p[1], order + p[2]) for p in el["props"]])
            rows = np.frombuffer(data, dtype=dt, count=el["count"], offset=bpos)
            bpos += dt.itemsize * el["count"]
            cols = {p[1]: rows[p[1]] for p in el["props"]}
            lists = {}
        else:
            cols = {p[1]: np.empty(el["count"], dtype=p[2]) for p in el["props"] if p[0] == "scalar"}
            lists = {p[1]: [] for p in el["props"] if p[0] == "list"}
            for r in range(el["count"]):
                for p in el["props"]:
                    if p[0] == "scalar":
                        if fmt == "ascii":
                            cols[p[1]][r] = float(tokens[tpos]) if p[2][0] == "f" else int(tokens[tpos])
                            tpos += 1
                        else:
                            cols[p[1]][r] = np.frombuffer(data, dtype=order + p[2], count=1, offset=bpos)[0]
                            bpos += np.dtype(p[2]).itemsize
                    else:
                        if fmt == "ascii":
                            n = int(tokens[tpos])
                            lists[p[1]].append([int(x) for x in tokens[tpos + 1:tpos + 1 + n]])
                            tpos += 1 + n
                        else:
                            n = int(np.frombuffer(data, dtype=order + p[2], count=1, offset=bpos)[0])
                            bpos += np.dtype(p[2]).itemsize
                            lists[p[1]].append(np.frombuffer(data, dtype=order + p[3], count=n, offset=bpos).astype(np.int64).tolist())
                            bpos += np.dtype(p[3]).itemsize * n
        if el["name"] == "vertex":
            if not all(k in cols for k in "xyz"):
                raise MeshError(f"\"{path}\": PLY vertex element without x/y/z")
            f32 = lambda k: np.asarray(cols[k]).astype(np.float32).astype(np.float64)
            verts = np.stack([f32("x"), f32("y"), f32("z")], 1)
            if "nx" in cols:                                                           # ply.cpp:338-344
                normals = np.stack([f32("nx"), f32("ny") if "ny" in cols else np.zeros(len(verts)),
                                    f32("nz") if "nz" in cols else np.zeros(len(verts))], 1)
        elif el["name"] == "face":
            faces = lists.get("vertex_indices", lists.get("vertex_index"))
            for face in faces or []:
                if len(face) not in (3, 4):
                    raise MeshError(f"Encountered a face with {len(face)} vertices! Only triangle and quad-based PLY meshes "
                                    "are supported for now.")
                tris.append((face[0], face[1], face[2]))
                if len(face) == 4:
                    tris.append((face[3], face[0], face[2]))                          # ply.cpp:306-309
    if verts is None or not len(verts) or not tris:
        raise MeshError(f"Unable to load \"{path}\" (no triangles or vertices found)!")
    tris = np.asarray(tris, dtype=np.int64)
    if tris.min() < 0 or tris.max() >= len(verts):
        raise MeshError(f"\"{path}\": PLY face index out of range")
    verts, normals, tris = apply_to_world(verts, normals, tris, to_world, flip_winding_on_mirror=False)
    return finish_mesh(verts, tris, normals, face_normals, flip_normals)
