"""Scene-file front end (gdb200.xmlscene): a Mitsuba 0.5 XML that selects the reference's `gpt` path
flattens to the same C-ABI scene as the programmatic builder — checked by rendering both with the CPU oracle — and the
host-side semantics of scenehandler.cpp / sensor.cpp (transform order, fov axes, $parameters, references, errors)."""
import math

import numpy as np
import pytest

import gdb200
from gdb200 import scenes, xmlscene


def _mat(m):
    return " ".join(repr(float(x)) for x in np.asarray(m).reshape(-1))


def _rect(center, s_axis, t_axis):
    s_axis, t_axis = np.asarray(s_axis, float), np.asarray(t_axis, float)
    n = np.cross(s_axis, t_axis)
    n /= np.linalg.norm(n)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = s_axis, t_axis, n, center
    return _mat(m)


CBOX_XML = f"""
<scene version="0.5.0">
  <default name="spp" value="6"/>
  <integrator type="gpt">
    <integer name="maxDepth" value="-1"/> <integer name="rrDepth" value="5"/>
    <boolean name="strictNormals" value="false"/> <float name="shiftThreshold" value="0.001"/>
    <boolean name="reconstructL1" value="false"/> <boolean name="reconstructL2" value="true"/>
    <float name="reconstructAlpha" value="0.2"/>
  </integrator>
  <sensor type="perspective">
    <float name="fov" value="39.3077"/>
    <transform name="toWorld"><lookat origin="0, 0, 3.9" target="0, 0, 0" up="0, 1, 0"/></transform>
    <sampler type="independent"><integer name="sampleCount" value="$spp"/></sampler>
    <film type="multifilm">
      <integer name="width" value="40"/> <integer name="height" value="32"/>
      <string name="fileFormat" value="pfm"/> <rfilter type="box"/>
    </film>
  </sensor>
  <bsdf type="diffuse" id="white"><rgb name="reflectance" value="0.725, 0.71, 0.68"/></bsdf>
  <bsdf type="diffuse" id="red"><rgb name="reflectance" value="0.63 0.065 0.05"/></bsdf>
  <bsdf type="diffuse" id="green"><rgb name="reflectance" value="0.14, 0.45, 0.091"/></bsdf>
  <bsdf type="diffuse" id="black"><spectrum name="reflectance" value="0"/></bsdf>
  <shape type="rectangle"><transform name="toWorld"><matrix value="{_rect((0, -1, 0), (1, 0, 0), (0, 0, -1))}"/></transform><ref id="white"/></shape>
  <shape type="rectangle"><transform name="toWorld"><matrix value="{_rect((0, 1, 0), (1, 0, 0), (0, 0, 1))}"/></transform><ref id="white"/></shape>
  <shape type="rectangle"><transform name="toWorld"><matrix value="{_rect((0, 0, -1), (1, 0, 0), (0, 1, 0))}"/></transform><ref id="white"/></shape>
  <shape type="rectangle"><transform name="toWorld"><matrix value="{_rect((-1, 0, 0), (0, 0, -1), (0, 1, 0))}"/></transform><ref id="red"/></shape>
  <shape type="rectangle"><transform name="toWorld"><matrix value="{_rect((1, 0, 0), (0, 0, 1), (0, 1, 0))}"/></transform><ref id="green"/></shape>
  <shape type="rectangle"><transform name="toWorld"><matrix value="{_rect((0, 0.99, 0), (0.25, 0, 0), (0, 0, 0.25))}"/></transform>
    <ref id="black"/> <emitter type="area"><rgb name="radiance" value="17, 12, 4"/></emitter></shape>
  <shape type="sphere"><point name="center" x="-0.55" y="-0.7" z="-0.2"/><float name="radius" value="0.3"/>
    <bsdf type="roughconductor"><string name="distribution" value="beckmann"/><float name="alpha" value="0.15"/>
      <rgb name="eta" value="0.2004, 0.9240, 1.1022"/><rgb name="k" value="3.9129, 2.4528, 2.1421"/><float name="extEta" value="1"/></bsdf></shape>
  <shape type="sphere"><point name="center" x="0.1" y="-0.65" z="0.45"/><float name="radius" value="0.35"/>
    <bsdf type="conductor"><rgb name="eta" value="1.6574, 0.8803, 0.5212"/><rgb name="k" value="9.2238, 6.2695, 4.8370"/><float name="extEta" value="1"/></bsdf></shape>
  <shape type="sphere"><point name="center" x="0.6" y="-0.75" z="-0.3"/><float name="radius" value="0.25"/>
    <bsdf type="dielectric"><float name="intIOR" value="1.5"/><float name="extIOR" value="1"/></bsdf></shape>
</scene>
"""


def _builder_equivalent(w, h):
    b = scenes._cornell(w, h, boxes=False)
    beck = b.material(type=scenes.BSDF_ROUGHCONDUCTOR, alpha=0.15, eta=scenes.CU_ETA, k=scenes.CU_K, distribution=scenes.MICROFACET_BECKMANN)
    mirror = b.material(type=scenes.BSDF_CONDUCTOR, eta=scenes.AL_ETA, k=scenes.AL_K)
    glass = b.material(type=scenes.BSDF_DIELECTRIC, ior_ratio=1.5)
    b.sphere((-0.55, -0.7, -0.2), 0.3, beck)
    b.sphere((0.1, -0.65, 0.45), 0.35, mirror)
    b.sphere((0.6, -0.75, -0.3), 0.25, glass)
    return b.build()


def test_xml_scene_renders_like_the_builder_scene(oracle):
    parsed = gdb200.load_scene(CBOX_XML)
    assert parsed.spp == 6 and parsed.seed == 0
    assert parsed.integrator_kwargs == dict(maxDepth=-1, rrDepth=5, strictNormals=False, shiftThreshold=0.001, reconstructL1=False,
                                            reconstructL2=True, reconstructAlpha=0.2)
    integ = parsed.integrator()
    ref_desc = _builder_equivalent(40, 32)
    a, _, _ = oracle.gpt(parsed.desc, integ.params(parsed.spp, parsed.seed))
    b, _, _ = oracle.gpt(ref_desc, integ.params(6, 0))
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert gdb200.load_scene(CBOX_XML, {"spp": "9"}).spp == 9                    # -D spp=9


def test_transform_operations_compose_in_document_order():
    import xml.etree.ElementTree as ET
    el = ET.fromstring('<transform name="toWorld"><scale x="2" y="3" z="4"/><rotate y="1" angle="90"/><translate x="1" y="2" z="3"/></transform>')
    m = xmlscene.parse_transform(el)
    p = m @ np.array([1.0, 1.0, 1.0, 1.0])                                        # scale first, then rotate about y, then translate
    assert np.allclose(p[:3], [4 + 1, 3 + 2, -2 + 3])
    el = ET.fromstring('<transform name="toWorld"><lookat origin="1,2,3" target="1,2,0"/></transform>')   # no 'up': an arbitrary axis is chosen
    m = xmlscene.parse_transform(el)
    assert np.allclose(m[:3, 2], [0, 0, -1]) and np.allclose(m[:3, 3], [1, 2, 3]) and abs(np.linalg.det(m[:3, :3])) > 0.99
    assert np.allclose(xmlscene.rotate((0, 0, 2), 90) @ np.array([1, 0, 0, 1.0]), [0, 1, 0, 1])


@pytest.mark.parametrize("props,expect_xfov", [
    ('<float name="fov" value="40"/>', 40.0),
    ('<float name="fov" value="40"/><string name="fovAxis" value="y"/>', math.degrees(2 * math.atan(math.tan(math.radians(20)) * 1.5))),
    ('<float name="fov" value="40"/><string name="fovAxis" value="smaller"/>', math.degrees(2 * math.atan(math.tan(math.radians(20)) * 1.5))),
    ('<float name="fov" value="40"/><string name="fovAxis" value="larger"/>', 40.0),
    ('<string name="focalLength" value="50mm"/>', math.degrees(2 * math.atan((2 * math.tan(0.5 * 2 * math.atan(math.sqrt(36 * 36 + 24 * 24) / 100)) / math.sqrt(1 + 1 / 2.25)) * 0.5))),
])
def test_field_of_view_axes(props, expect_xfov):
    xml = f"""<scene version="0.5.0"><integrator type="gpt"/><sensor type="perspective">{props}
      <film type="multifilm"><integer name="width" value="30"/><integer name="height" value="20"/></film></sensor>
      <emitter type="point"><point name="position" x="0" y="1" z="0"/><rgb name="intensity" value="1,1,1"/></emitter></scene>"""
    parsed = gdb200.load_scene(xml)
    s2c = np.array(parsed.desc.camera.sample_to_camera).reshape(4, 4)
    near = s2c @ np.array([0.0, 0.5, 0.0, 1.0])                                     # left edge of the film, mid height
    near = near[:3] / near[3]
    assert math.isclose(math.degrees(2 * math.atan(abs(near[0]) / near[2])), expect_xfov, rel_tol=1e-9)
    assert parsed.desc.rfilter_radius == 2.0                                        # film default: gaussian (film.cpp:89-95)
    assert parsed.spp == 4 and parsed.integrator_kwargs == {}


def test_shapes_and_emitters(tmp_path, oracle):
    obj = tmp_path / "quad.obj"
    obj.write_text("v -1 0 -1\nv 1 0 -1\nv 1 0 1\nv -1 0 1\nvn 0 1 0\nf 1//1 4//1 3//1 2//1\n")
    env = tmp_path / "sky.pfm"
    gdb200.pfm.write_pfm(env, scenes.sky_envmap(16, 8))
    xml = f"""<scene version="0.5.0"><integrator type="gpt"><integer name="streamsPerPixel" value="2"/><integer name="seed" value="7"/></integrator>
      <sensor type="thinlens"><float name="apertureRadius" value="0.05"/><float name="focusDistance" value="4"/><float name="fov" value="45"/>
        <transform name="toWorld"><lookat origin="0,1.5,5" target="0,0.5,0" up="0,1,0"/></transform>
        <sampler type="independent"><integer name="sampleCount" value="3"/></sampler>
        <film type="multifilm"><integer name="width" value="20"/><integer name="height" value="16"/><rfilter type="tent"/></film></sensor>
      <bsdf type="twosided" id="sheet"><bsdf type="plastic"><rgb name="diffuseReflectance" value=".3,.4,.5"/></bsdf></bsdf>
      <shape type="obj"><string name="filename" value="quad.obj"/><ref id="sheet"/></shape>
      <shape type="obj"><string name="filename" value="quad.obj"/><boolean name="faceNormals" value="true"/>
        <transform name="toWorld"><scale value="0.2"/><rotate x="1" angle="180"/><translate y="2"/></transform>
        <emitter type="area"><rgb name="radiance" value="10,10,10"/></emitter></shape>
      <shape type="cube"><transform name="toWorld"><scale value="0.3"/><translate x="0.5" y="0.3"/></transform><bsdf type="diffuse"/></shape>
      <shape type="sphere"><transform name="toWorld"><scale value="0.25"/><translate x="-0.6" y="0.25"/></transform><bsdf type="dielectric"/></shape>
      <emitter type="point"><transform name="toWorld"><translate x="1" y="2" z="1"/></transform><spectrum name="intensity" value="3"/></emitter>
      <emitter type="envmap"><string name="filename" value="sky.pfm"/><float name="scale" value="0.5"/></emitter></scene>"""
    path = tmp_path / "scene.xml"
    path.write_text(xml)
    parsed = gdb200.load_scene(str(path))
    d = parsed.desc
    assert (parsed.spp, parsed.seed, parsed.streams) == (3, 7, 2)
    assert d.camera.aperture_radius == 0.05 and d.camera.focus_distance == 4 and d.rfilter_radius == 1.0
    assert d.n_shapes == 4 and d.n_emitters == 3 and d.n_triangles == 2 + 2 + 12 and d.n_vertices == 4 + 4 + 24
    # Scene::m_emitters order: scene-level emitters in document order (scene.cpp:496-516), then the shapes' area lights (scene.cpp:570-571)
    assert [d.emitters[i].type for i in range(3)] == [scenes.EMITTER_POINT, scenes.EMITTER_ENVMAP, scenes.EMITTER_AREA]
    assert d.shapes[0].has_vertex_normals == 1 and d.shapes[1].has_vertex_normals == 0 and d.shapes[2].has_vertex_normals == 1
    assert d.materials[d.shapes[0].material].twosided == 1 and d.materials[d.shapes[0].material].type == scenes.BSDF_PLASTIC
    assert math.isclose(d.shapes[3].radius, 0.25) and np.allclose(list(d.shapes[3].center), [-0.6, 0.25, 0])
    assert d.materials[d.shapes[3].material].ior_ratio == float(np.float32(1.5046)) / float(np.float32(1.000277))   # ior.h: float literals
    assert d.envmap.contents.width == 16 and d.envmap.contents.scale == 0.5
    out, _, cnt = oracle.gpt(d, parsed.integrator().params(parsed.spp, parsed.seed, streams=parsed.streams))   # and it renders
    assert cnt[0] == 20 * 16 * 3 and np.isfinite(out["-throughput"]).all() and out["-throughput"].mean() > 0


@pytest.mark.parametrize("fragment,message", [
    ('<film type="hdrfilm"/>', "without MultiFilm"),
    ('<film type="multifilm"/><bogus/>', None),
])
def test_film_must_be_multifilm(fragment, message):
    xml = f"""<scene version="0.5.0"><integrator type="gpt"/><sensor type="perspective"><float name="fov" value="40"/>{fragment}</sensor>
      <emitter type="point"><point name="position" x="0" y="1" z="0"/><rgb name="intensity" value="1,1,1"/></emitter></scene>"""
    if message:
        with pytest.raises(gdb200.Gdb200Error, match=message):
            gdb200.load_scene(xml)
    else:
        assert gdb200.load_scene(xml).desc.camera.width == 1          # multifilm's default size (film.cpp:31-32)


@pytest.mark.parametrize("body,message", [
    ('<integrator type="path"/>', 'only "gpt"'),
    ('<integrator type="gpt"><float name="bogus" value="1"/></integrator>', "Unqueried property"),
    ('<integrator type="gpt"/><shape type="hair"/>', 'shape plugin "hair"'),
    ('<integrator type="gpt"/><bsdf type="roughplastic"/>', 'BSDF plugin "roughplastic"'),
    ('<integrator type="gpt"/><bsdf type="conductor"/>', "spectral data files"),
    ('<integrator type="gpt"/><emitter type="directional"/>', 'emitter plugin "directional"'),
    ('<integrator type="gpt"/><shape type="sphere"><ref id="nope"/></shape>', "not found"),
    ('<integrator type="gpt"><integer name="maxDepth" value="$depth"/></integrator>', r'"\$depth" was not specified'),
])
def test_outside_the_subset_fails_loudly(body, message):
    xml = f"""<scene version="0.5.0">{body}<sensor type="perspective"><float name="fov" value="40"/><film type="multifilm"/></sensor>
      <emitter type="point"><point name="position" x="0" y="1" z="0"/><rgb name="intensity" value="1,1,1"/></emitter></scene>"""
    with pytest.raises(gdb200.Gdb200Error, match=message):
        gdb200.load_scene(xml)


def test_integrator_validation_applies_to_scene_files():
    xml = """<scene version="0.5.0"><integrator type="gpt"><boolean name="reconstructL1" value="true"/><boolean name="reconstructL2" value="true"/></integrator>
      <sensor type="perspective"><float name="fov" value="40"/><film type="multifilm"/></sensor>
      <emitter type="point"><point name="position" x="0" y="1" z="0"/><rgb name="intensity" value="1,1,1"/></emitter></scene>"""
    with pytest.raises(gdb200.Gdb200Error, match="Cannot display two reconstructions"):
        gdb200.load_scene(xml).integrator()


def test_obj_without_normals_gets_angle_weighted_vertex_normals(tmp_path):
    """TriMesh::computeNormals (trimesh.cpp:631-672): a closed octahedron's synthesised normals point radially."""
    obj = tmp_path / "octa.obj"
    obj.write_text("v 1 0 0\nv -1 0 0\nv 0 1 0\nv 0 -1 0\nv 0 0 1\nv 0 0 -1\n"
                   "f 1 3 5\nf 3 2 5\nf 2 4 5\nf 4 1 5\nf 3 1 6\nf 2 3 6\nf 4 2 6\nf 1 4 6\n")
    verts, tris, nrms = xmlscene.load_obj(str(obj))
    assert len(verts) == 6 and len(tris) == 8 and nrms is not None
    for v, n in zip(verts, nrms):
        assert np.allclose(n, np.asarray(v) / np.linalg.norm(v), atol=1e-12)
    _, _, flat = xmlscene.load_obj(str(obj), face_normals=True)
    assert flat is None


def test_spot_emitter_from_xml(tmp_path, oracle):
    """<emitter type="spot"> (spot.cpp:70-75): lookat toWorld, cutoffAngle / beamWidth in degrees with the 3/4 default,
    the inverse rotation the falloff uses; rendering the parsed scene equals rendering the builder's scene."""
    xml = """<scene version="0.5.0"><integrator type="gpt"/>
      <sensor type="perspective"><float name="fov" value="50"/>
        <transform name="toWorld"><lookat origin="0,2,4" target="0,0,0" up="0,1,0"/></transform>
        <sampler type="independent"><integer name="sampleCount" value="4"/></sampler>
        <film type="multifilm"><integer name="width" value="24"/><integer name="height" value="16"/><rfilter type="box"/></film></sensor>
      <shape type="rectangle"><transform name="toWorld"><rotate x="1" angle="-90"/><scale value="3"/></transform><bsdf type="diffuse"/></shape>
      <emitter type="spot"><transform name="toWorld"><lookat origin="1,2,0.5" target="0,0,0" up="0,1,0"/></transform>
        <rgb name="intensity" value="9,8,7"/><float name="cutoffAngle" value="32"/></emitter>
      <emitter type="spot"><transform name="toWorld"><lookat origin="-1,1.5,0" target="-0.5,0,0.5"/></transform>
        <rgb name="intensity" value="2,3,4"/><float name="cutoffAngle" value="40"/><float name="beamWidth" value="10"/>
        <float name="samplingWeight" value="2"/></emitter></scene>"""
    path = tmp_path / "spot.xml"
    path.write_text(xml)
    parsed = gdb200.load_scene(str(path))
    d = parsed.desc
    assert d.n_emitters == 2 and all(d.emitters[i].type == scenes.EMITTER_SPOT for i in range(2))
    e0, e1 = d.emitters[0], d.emitters[1]
    assert math.isclose(e0.cutoff_angle, math.radians(32)) and math.isclose(e0.beam_width, math.radians(24))
    assert math.isclose(e1.cutoff_angle, math.radians(40)) and math.isclose(e1.beam_width, math.radians(10)) and e1.sampling_weight == 2
    assert np.allclose(list(e0.position), [1, 2, 0.5])
    axis = np.array([-1, -2, -0.5]) / np.linalg.norm([1, 2, 0.5])
    assert np.allclose(np.array(list(e0.to_local)).reshape(3, 3) @ axis, [0, 0, 1], atol=1e-12)     # the cone axis is local +z
    out, _, cnt = oracle.gpt(d, parsed.integrator().params(parsed.spp, parsed.seed))
    assert cnt[0] == 24 * 16 * 4 and out["-throughput"].max() > 0.1 and (out["-throughput"].sum(-1) == 0).mean() > 0.1   # lit cone, dark outside

    bad = xml.replace('<float name="cutoffAngle" value="32"/>', '<float name="cutoffAngle" value="10"/><float name="beamWidth" value="20"/>')
    path.write_text(bad)
    with pytest.raises(Exception, match="cutoffAngle"):
        gdb200.load_scene(str(path))
    path.write_text(xml.replace('<float name="cutoffAngle" value="32"/>', '<texture name="texture" type="bitmap"/>'))
    with pytest.raises(Exception, match="projection textures"):
        gdb200.load_scene(str(path))


def test_obj_flip_normals(tmp_path):
    """flipNormals on a TriMesh (trimesh.cpp:610-628,662-664): vertex normals negated, or the winding swapped with faceNormals."""
    (tmp_path / "t.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    _, tris, nrms = xmlscene.load_obj(str(tmp_path / "t.obj"))
    assert np.allclose(nrms, [(0, 0, 1)] * 3) and tris == [(0, 1, 2)]
    _, tris, nrms = xmlscene.load_obj(str(tmp_path / "t.obj"), flip_normals=True)
    assert np.allclose(nrms, [(0, 0, -1)] * 3) and tris == [(0, 1, 2)]
    _, tris, nrms = xmlscene.load_obj(str(tmp_path / "t.obj"), face_normals=True, flip_normals=True)
    assert nrms is None and tris == [(1, 0, 2)]


def test_film_file_format(tmp_path):
    """multifilm.cpp:110-128,222-235: fileFormat defaults to openexr / float16; pfm forces float32; anything else raises."""
    xml = """<scene version="0.5.0"><integrator type="gpt"/>
      <sensor type="perspective"><film type="multifilm"><integer name="width" value="8"/><integer name="height" value="8"/>FMT</film></sensor>
      <shape type="rectangle"><bsdf type="diffuse"/></shape>
      <emitter type="point"><point name="position" x="0" y="0" z="1"/><rgb name="intensity" value="1,1,1"/></emitter></scene>"""
    path = tmp_path / "f.xml"
    for fmt, expect in (("", ("openexr", "float16")), ('<string name="fileFormat" value="PFM"/>', ("pfm", "float32")),
                        ('<string name="componentFormat" value="float32"/>', ("openexr", "float32"))):
        path.write_text(xml.replace("FMT", fmt))
        parsed = gdb200.load_scene(str(path))
        assert (parsed.file_format, parsed.component_format) == expect
    for fmt, msg in (('<string name="fileFormat" value="rgbe"/>', "fileFormat"), ('<string name="componentFormat" value="uint32"/>', "componentFormat"),
                     ('<string name="pixelFormat" value="luminance"/>', "pixelFormat")):
        path.write_text(xml.replace("FMT", fmt))
        with pytest.raises(Exception, match=msg):
            gdb200.load_scene(str(path))


def test_envmap_from_openexr(tmp_path):
    """envmap `filename` as .exr (the usual container for light probes) gives the same texels as the .pfm."""
    from gdb200 import exr, pfm
    sky = scenes.sky_envmap(16, 8)
    pfm.write_pfm(str(tmp_path / "sky.pfm"), sky)
    exr.write_exr(str(tmp_path / "sky.exr"), sky, "float32")
    xml = """<scene version="0.5.0"><integrator type="gpt"/>
      <sensor type="perspective"><film type="multifilm"><integer name="width" value="8"/><integer name="height" value="8"/></film></sensor>
      <shape type="sphere"><bsdf type="diffuse"/></shape>
      <emitter type="envmap"><string name="filename" value="sky.EXT"/></emitter></scene>"""
    texels = []
    for ext in ("pfm", "exr"):
        (tmp_path / "s.xml").write_text(xml.replace("EXT", ext))
        d = gdb200.load_scene(str(tmp_path / "s.xml")).desc
        env = d.envmap.contents
        texels.append(np.ctypeslib.as_array(env.rgb, shape=(env.height, env.width, 3)).copy())
    assert np.array_equal(texels[0], texels[1]) and texels[0].shape == (8, 16, 3)


def test_include_alias_and_colour_syntax(tmp_path):
    """<include> merges the included scene's objects in place and shares ids and parameters (scenehandler.cpp:658-681),
    <alias> gives an object a second id (:646-656), colours may be "#rrggbb" and <srgb> applies the transfer curve
    (:461-545, spectrum.cpp:400-419)."""
    sub = tmp_path / "parts"
    sub.mkdir()
    (sub / "tri.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    (sub / "mats.xml").write_text("""<scene version="0.5.0"><default name="spp" value="9"/><default name="tint" value="0.25"/>
      <bsdf type="diffuse" id="paint"><srgb name="reflectance" value="#ff8000"/></bsdf>
      <shape type="obj"><string name="filename" value="tri.obj"/><ref id="paint"/></shape></scene>""")
    (tmp_path / "main.xml").write_text("""<scene version="0.5.0"><default name="spp" value="5"/><integrator type="gpt"/>
      <sensor type="perspective"><sampler type="independent"><integer name="sampleCount" value="$spp"/></sampler>
        <film type="multifilm"><integer name="width" value="8"/><integer name="height" value="8"/></film></sensor>
      <include filename="parts/mats.xml"/>
      <alias id="paint" as="wallpaint"/>
      <shape type="rectangle"><ref id="wallpaint"/></shape>
      <shape type="sphere"><bsdf type="diffuse"><rgb name="reflectance" value="#336699"/></bsdf></shape>
      <shape type="sphere"><bsdf type="diffuse"><srgb name="reflectance" value="$tint"/></bsdf></shape>
      <emitter type="point"><point name="position" x="0" y="0" z="3"/><rgb name="intensity" value="5"/></emitter></scene>""")
    parsed = gdb200.load_scene(str(tmp_path / "main.xml"))
    d = parsed.desc
    assert parsed.spp == 5                                                     # the first <default> wins, the include's does not override
    assert d.n_shapes == 4 and d.shapes[0].type == scenes.SHAPE_MESH           # the included mesh comes first (document order)
    assert d.shapes[0].material == d.shapes[1].material                        # alias -> the same BSDF instance
    lin = lambda v: v / 12.92 if v <= 0.04045 else ((v + 0.055) / 1.055) ** 2.4
    assert np.allclose(list(d.materials[d.shapes[0].material].reflectance), [1.0, lin(128 / 255), 0.0])
    assert np.allclose(list(d.materials[d.shapes[2].material].reflectance), [0x33 / 255, 0x66 / 255, 0x99 / 255])
    assert np.allclose(list(d.materials[d.shapes[3].material].reflectance), [lin(0.25)] * 3)
    assert list(d.emitters[0].radiance) == [5.0, 5.0, 5.0]
    (tmp_path / "bad.xml").write_text((tmp_path / "main.xml").read_text().replace('value="#336699"', 'value="400:0.1, 700:0.9"').replace('<rgb name="reflectance"', '<spectrum name="reflectance"', 1))
    with pytest.raises(Exception, match="CIE"):
        gdb200.load_scene(str(tmp_path / "bad.xml"))
    (tmp_path / "dup.xml").write_text((tmp_path / "main.xml").read_text().replace('as="wallpaint"', 'as="paint"'))
    with pytest.raises(Exception, match="Duplicate ID"):
        gdb200.load_scene(str(tmp_path / "dup.xml"))
