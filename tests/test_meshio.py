"""Mesh files of the reference's shape plugins (gdb200.meshio): `serialized` (serialized.cpp:78-146, trimesh.cpp:175-293)
and `ply` (ply.cpp).  The byte layouts are built here by hand from the reference's format description, so the reader is
checked against the specification and not against its own writer alone."""
import math
import struct
import zlib

import numpy as np
import pytest

import gdb200
from gdb200 import meshio, scenes, xmlscene

OCTA_V = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]
OCTA_T = [(0, 2, 4), (2, 1, 4), (1, 3, 4), (3, 0, 4), (2, 0, 5), (1, 2, 5), (3, 1, 5), (0, 3, 5)]


def _raw_serialized(version, flags, name, verts, tris, normals=None, uvs=None, colors=None):
    """One mesh record as the format description lays it out (serialized.cpp:86-128)."""
    dt = "<f8" if flags & 0x2000 else "<f4"
    body = struct.pack("<I", flags)
    if version == 4:
        body += name.encode() + b"\0"
    body += struct.pack("<QQ", len(verts), len(tris)) + np.asarray(verts, dt).tobytes()
    for arr in (normals, uvs, colors):
        if arr is not None:
            body += np.asarray(arr, dt).tobytes()
    body += np.asarray(tris, "<u4").tobytes()
    return struct.pack("<HH", 0x041C, version) + zlib.compress(body)


def test_serialized_v4_multi_mesh_file(tmp_path):
    quad_v, quad_t = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0)], [(0, 1, 2), (0, 2, 3)]
    nrm = [(0, 0, 1)] * 4
    recs = [_raw_serialized(4, 0x1000, "octa", OCTA_V, OCTA_T),
            _raw_serialized(4, 0x2000 | 0x0001 | 0x0002 | 0x0008, "quad", quad_v, quad_t, normals=nrm,
                            uvs=[(0, 0), (1, 0), (1, 1), (0, 1)], colors=[(1, 0, 0)] * 4)]
    offs, blob = [], b""
    for r in recs:
        offs.append(len(blob))
        blob += r
    blob += b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<I", len(offs))     # serialized.cpp:131-143
    path = tmp_path / "two.serialized"
    path.write_bytes(blob)

    v, t, n = meshio.load_serialized(str(path), 0)
    assert np.array_equal(v, np.array(OCTA_V, float)) and np.array_equal(t, np.array(OCTA_T))
    assert np.allclose(n, np.array(OCTA_V, float))                                # generated smooth normals of an octahedron are radial
    v, t, n = meshio.load_serialized(str(path), 1)
    assert np.array_equal(v, np.array(quad_v, float)) and np.array_equal(t, np.array(quad_t)) and np.array_equal(n, np.array(nrm, float))
    v, t, n = meshio.load_serialized(str(path), 1, face_normals=True, flip_normals=True)
    assert n is None and np.array_equal(t, np.array([(1, 0, 2), (2, 0, 3)]))       # faceNormals + flipNormals: winding swapped
    v, t, n = meshio.load_serialized(str(path), 1, flip_normals=True)
    assert np.array_equal(n, -np.array(nrm, float))
    with pytest.raises(meshio.MeshError, match="out of range"):
        meshio.load_serialized(str(path), 3)


def test_serialized_v3_and_transform(tmp_path):
    """Version 3 has no name field and a uint32 dictionary; a mirroring toWorld swaps the winding (serialized.cpp:197-202)
    and normals go through the inverse transpose."""
    nrm = [(0, 0, 1)] * 3
    rec = _raw_serialized(3, 0x1000 | 0x0001, "", [(0, 0, 0), (1, 0, 0), (0, 1, 0)], [(0, 1, 2)], normals=nrm)
    rec2 = _raw_serialized(3, 0x1000, "", [(0, 0, 0), (2, 0, 0), (0, 2, 0)], [(0, 1, 2)])
    blob = rec + rec2 + struct.pack("<II", 0, len(rec)) + struct.pack("<I", 2)
    path = tmp_path / "v3.serialized"
    path.write_bytes(blob)
    m = scenes.translate((0, 0, 5)) @ scenes.scale((2, 1, -3))
    v, t, n = meshio.load_serialized(str(path), 0, to_world=m)
    assert np.allclose(v, [(0, 0, 5), (2, 0, 5), (0, 1, 5)]) and np.array_equal(t, [(1, 0, 2)]) and np.allclose(n, [(0, 0, -1)] * 3)
    v, t, n = meshio.load_serialized(str(path), 1)
    assert np.allclose(v[1], (2, 0, 0)) and np.allclose(n, [(0, 0, 1)] * 3)


@pytest.mark.parametrize("blob,msg", [(struct.pack("<HH", 0x1C04, 4), "old version"), (struct.pack("<HH", 0x1234, 4), "invalid file format"),
                                      (struct.pack("<HH", 0x041C, 5), "incompatible file version")])
def test_serialized_header_errors(tmp_path, blob, msg):
    path = tmp_path / "bad.serialized"
    path.write_bytes(blob + b"\0" * 16)
    with pytest.raises(meshio.MeshError, match=msg):
        meshio.load_serialized(str(path))


def test_save_serialized_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    v = rng.normal(size=(50, 3))
    t = rng.integers(0, 50, size=(80, 3))
    n = rng.normal(size=(50, 3))
    path = tmp_path / "rt.serialized"
    meshio.save_serialized(str(path), [("a", v, t, n), ("b", v[:10], t[:5] % 10, None)], double_precision=True)
    v2, t2, n2 = meshio.load_serialized(str(path), 0)
    assert np.array_equal(v2, v) and np.array_equal(t2, t) and np.array_equal(n2, n)
    v3, t3, n3 = meshio.load_serialized(str(path), 1, face_normals=True)
    assert np.array_equal(v3, v[:10]) and np.array_equal(t3, t[:5] % 10) and n3 is None
    meshio.save_serialized(str(path), [("a", v, t, None)])                              # single precision
    v4, _, _ = meshio.load_serialized(str(path), 0, face_normals=True)
    assert np.array_equal(v4, v.astype(np.float32).astype(np.float64))


PLY_HEADER = """ply
format {fmt} 1.0
comment made by hand
element vertex 5
property float x
property float y
property float z
property float nx
property float ny
property float nz
property uchar red
property uchar green
property uchar blue
element face 3
property list uchar int vertex_indices
end_header
"""
PLY_V = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0.5, 0.5, 1)]
PLY_F = [(0, 1, 2, 3), (0, 1, 4), (1, 2, 4)]


def _ply_bytes(fmt):
    head = PLY_HEADER.format(fmt=fmt).encode()
    if fmt == "ascii":
        body = "".join("%g %g %g 0 0 1 255 0 0\n" % v for v in PLY_V) + "".join("%d %s\n" % (len(f), " ".join(map(str, f))) for f in PLY_F)
        return head + body.encode()
    e = "<" if fmt == "binary_little_endian" else ">"
    body = b"".join(struct.pack(e + "6f3B", *v, 0, 0, 1, 255, 0, 0) for v in PLY_V)
    body += b"".join(struct.pack(e + "B%di" % len(f), len(f), *f) for f in PLY_F)
    return head + body


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_ply_formats(tmp_path, fmt):
    path = tmp_path / "m.ply"
    path.write_bytes(_ply_bytes(fmt))
    v, t, n = meshio.load_ply(str(path))
    assert np.array_equal(v, np.array(PLY_V, float))
    assert np.array_equal(t, [(0, 1, 2), (3, 0, 2), (0, 1, 4), (1, 2, 4)])          # quad split of ply.cpp:299-312
    assert np.array_equal(n, [(0, 0, 1)] * 5)
    v, t, n = meshio.load_ply(str(path), to_world=scenes.scale((-1, 1, 1)), face_normals=True)
    assert n is None and np.array_equal(t[0], (0, 1, 2)) and v[1][0] == -1           # no winding swap in the ply plugin


def test_ply_errors(tmp_path):
    path = tmp_path / "m.ply"
    path.write_bytes(PLY_HEADER.format(fmt="ascii").replace("element face 3", "element face 1").encode()
                     + ("".join("%g %g %g 0 0 1 255 0 0\n" % v for v in PLY_V) + "5 0 1 2 3 4\n").encode())
    with pytest.raises(meshio.MeshError, match="face with 5 vertices"):
        meshio.load_ply(str(path))
    path.write_bytes(b"solid not a ply")
    with pytest.raises(meshio.MeshError, match="not a PLY"):
        meshio.load_ply(str(path))


def test_vectorised_normals_match_the_loop_restatement():
    """meshio.compute_normals (numpy) against xmlscene.compute_normals (the per-triangle loop of trimesh.cpp:631-672),
    including a degenerate triangle and an unreferenced vertex."""
    rng = np.random.default_rng(11)
    v = rng.normal(size=(40, 3))
    t = rng.integers(0, 38, size=(90, 3))                                           # vertices 38, 39 unreferenced
    t[7] = (3, 3, 5)                                                                # zero-area triangle
    t = t[(t[:, 0] != t[:, 1]) | (np.arange(90) == 7)]
    t = np.array([r for r in t if len(set(r)) == 3 or tuple(r) == (3, 3, 5)])
    a = meshio.compute_normals(v, t)
    b = np.array(xmlscene.compute_normals(list(v), [tuple(r) for r in t]))
    assert np.allclose(a, b, rtol=0, atol=1e-15) and np.array_equal(a[39], (1, 0, 0))
    assert np.allclose(meshio.compute_normals(v, t, flip=True)[:38], -a[:38])


def test_scene_with_serialized_and_ply_shapes(tmp_path, oracle):
    meshio.save_serialized(str(tmp_path / "octa.serialized"), [("octa", OCTA_V, OCTA_T, None)])
    (tmp_path / "m.ply").write_bytes(_ply_bytes("binary_little_endian"))
    xml = """<scene version="0.5.0"><integrator type="gpt"/>
      <sensor type="perspective"><float name="fov" value="45"/>
        <transform name="toWorld"><lookat origin="0,1.5,5" target="0,0.3,0" up="0,1,0"/></transform>
        <sampler type="independent"><integer name="sampleCount" value="2"/></sampler>
        <film type="multifilm"><integer name="width" value="16"/><integer name="height" value="12"/><rfilter type="box"/></film></sensor>
      <shape type="rectangle"><transform name="toWorld"><rotate x="1" angle="-90"/><scale value="4"/></transform><bsdf type="diffuse"/></shape>
      <shape type="serialized"><string name="filename" value="octa.serialized"/><integer name="shapeIndex" value="0"/>
        <transform name="toWorld"><scale value="0.5"/><translate x="-0.8" y="0.5"/></transform><bsdf type="diffuse"/></shape>
      <shape type="ply"><string name="filename" value="m.ply"/><boolean name="faceNormals" value="true"/>
        <transform name="toWorld"><translate x="0.3" y="0.01"/></transform><bsdf type="diffuse"/></shape>
      <emitter type="point"><point name="position" x="0" y="3" z="1"/><rgb name="intensity" value="20,20,20"/></emitter></scene>"""
    (tmp_path / "s.xml").write_text(xml)
    parsed = gdb200.load_scene(str(tmp_path / "s.xml"))
    d = parsed.desc
    assert d.n_shapes == 3 and d.n_triangles == 8 + 4 and d.n_vertices == 6 + 5
    assert d.shapes[1].has_vertex_normals == 1 and d.shapes[2].has_vertex_normals == 0
    out, _, cnt = oracle.gpt(d, parsed.integrator().params(parsed.spp, parsed.seed))
    assert cnt[0] == 16 * 12 * 2 and np.isfinite(out["-throughput"]).all() and out["-throughput"].mean() > 0
    (tmp_path / "s2.xml").write_text(xml.replace('<integer name="shapeIndex" value="0"/>', '<float name="maxSmoothAngle" value="30"/>'))
    with pytest.raises(Exception, match="maxSmoothAngle"):
        gdb200.load_scene(str(tmp_path / "s2.xml"))


# ---------------------------------------------------------------- against the reference's own loaders (oracle/_ref)
def _ref_mesh(kind, path, index=0, face_normals=False, flip_normals=False):
    import ctypes
    from conftest import RefMitsuba
    if not RefMitsuba.available():
        pytest.skip("needs the compiled reference")
    lib = RefMitsuba().lib
    lib.gdbref_last_error.restype = ctypes.c_char_p
    counts = (ctypes.c_int * 3)()
    assert lib.gdbref_mesh_load(kind, str(path).encode(), index, int(face_normals), int(flip_normals), counts) == 0, lib.gdbref_last_error()
    v, n, t = np.zeros((counts[0], 3)), np.zeros((counts[0], 3)), np.zeros((counts[1], 3), dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    assert lib.gdbref_mesh_copy(p(v), p(n), p(t)) == 0
    return v, t, (n if counts[2] else None)


def _bumpy_sphere(seed=0):
    v, t, _ = scenes.uv_sphere_mesh((0.1, -0.2, 0.3), 0.8, segments=10, rings=6)
    rng = np.random.default_rng(seed)
    return np.asarray(v) * (1 + 0.1 * rng.random((len(v), 1))), np.asarray(t)


@pytest.mark.parametrize("double_precision", [False, True])
def test_serialized_files_load_like_the_reference(tmp_path, double_precision):
    """Files written by save_serialized through TriMesh(Stream *, index) + TriMesh::configure (trimesh.cpp:80-86,175-252,
    608-681): same vertices and triangles, and the generated smooth normals of meshio.compute_normals agree with the
    reference's to rounding."""
    v, t = _bumpy_sphere()
    n = np.random.default_rng(1).normal(size=v.shape)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    path = tmp_path / "m.serialized"
    meshio.save_serialized(str(path), [("plain", v, t, None), ("with normals", v, t, n)], double_precision=double_precision)
    for index in (0, 1):
        rv, rt, rn = _ref_mesh(0, path, index)
        ov, ot, on = meshio.load_serialized(str(path), index)
        assert np.array_equal(rv, ov) and np.array_equal(rt, ot)
        np.testing.assert_allclose(on, rn, rtol=0, atol=1e-14 if index == 0 else 0)


def test_obj_files_load_like_the_reference(tmp_path):
    """An OBJ without normals through the reference's obj plugin: the same triangles over the same positions and the same
    generated normals per triangle corner (the two loaders may number the vertices differently)."""
    v, t = _bumpy_sphere(3)
    path = tmp_path / "m.obj"
    path.write_text("".join("v %.9g %.9g %.9g\n" % tuple(p) for p in v) + "".join("f %d %d %d\n" % tuple(i + 1 for i in tri) for tri in t))
    rv, rt, rn = _ref_mesh(2, path)
    ov, ot, on = xmlscene.load_obj(str(path))
    ov, on = np.asarray(ov), np.asarray(on)
    assert len(rt) == len(ot) and rn is not None
    np.testing.assert_allclose(ov[np.asarray(ot)], rv[rt], rtol=0, atol=0)          # corner positions, triangle by triangle
    np.testing.assert_allclose(on[np.asarray(ot)], rn[rt], rtol=0, atol=1e-13)      # corner normals
