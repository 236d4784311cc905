"""The integrator plugin's scene flattening (plugin/gpt_plugin.cpp: Mitsuba Scene -> gdb200_scene_desc) on REAL Mitsuba
objects: a gdb200_scene_desc is turned into the reference's own Scene / Shape / BSDF / Emitter / Sensor / Film instances by
oracle/_ref/libref_mitsuba.so, the plugin flattens that Scene again, and the result must describe the same scene -- and
render to the same image with the CPU restatement.  (Rendering through the plugin needs a GPU; the flattening does not.)"""
import ctypes
import os

import numpy as np
import pytest

import gdb200  # noqa: F401
from gdb200 import scenes
from conftest import ROOT, REFERENCE, _make

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_plugin_roundtrip.so")


@pytest.fixture(scope="module")
def flatten():
    if not os.path.exists(LIB):
        if not os.path.isdir(REFERENCE):
            pytest.skip("needs oracle/_ref/libref_plugin_roundtrip.so (a build of /root/reference)")
        _make("ref")
    lib = ctypes.CDLL(LIB)
    lib.gdbref_plugin_flatten.restype = ctypes.POINTER(scenes.SceneDesc)
    lib.gdbref_roundtrip_last_error.restype = ctypes.c_char_p

    def run(desc, prm):
        fov, rfilter = scenes.mitsuba_sensor_args(desc)
        out = lib.gdbref_plugin_flatten(ctypes.byref(desc), ctypes.byref(prm), ctypes.c_double(fov), rfilter.encode())
        if not out:
            raise RuntimeError(lib.gdbref_roundtrip_last_error().decode())
        return out.contents
    return run


def _arr(x, n=None):
    return np.array(list(x) if n is None else [x[i] for i in range(n)], dtype=float)


@pytest.mark.parametrize("name", ["cbox_diffuse", "cbox_glossy", "cbox_materials", "cbox_mesh_lights", "cbox_smooth", "cbox_point",
                                  "cbox_spot", "cbox_dof", "cbox_roughglass", "cbox_sphere_lights", "cbox_env"])
def test_flattening_a_mitsuba_scene_gives_back_the_description(flatten, oracle, name):
    desc = getattr(scenes, name)(20, 16)
    prm = scenes.default_params(spp=2, seed=4)
    got = flatten(desc, prm)
    cam, gcam = desc.camera, got.camera
    assert (gcam.width, gcam.height) == (cam.width, cam.height) and gcam.near_clip == cam.near_clip and gcam.far_clip == cam.far_clip
    np.testing.assert_allclose(_arr(gcam.sample_to_camera), _arr(cam.sample_to_camera), rtol=1e-12, atol=1e-15)   # numpy's inverse vs Mitsuba's
    np.testing.assert_array_equal(_arr(gcam.camera_to_world), _arr(cam.camera_to_world))
    assert gcam.aperture_radius == cam.aperture_radius and gcam.focus_distance == cam.focus_distance
    assert got.rfilter_radius == desc.rfilter_radius
    assert (got.n_shapes, got.n_emitters, got.n_triangles) == (desc.n_shapes, desc.n_emitters, desc.n_triangles)
    for i in range(desc.n_emitters):
        a, b = desc.emitters[i], got.emitters[i]
        assert (a.type, a.shape, a.sampling_weight) == (b.type, b.shape, b.sampling_weight), i
        np.testing.assert_array_equal(_arr(a.radiance), _arr(b.radiance))
        if a.type in (scenes.EMITTER_POINT, scenes.EMITTER_SPOT):
            np.testing.assert_allclose(_arr(a.position), _arr(b.position), rtol=0, atol=1e-15)
        if a.type == scenes.EMITTER_SPOT:
            np.testing.assert_allclose(_arr(a.to_local), _arr(b.to_local), rtol=0, atol=1e-14)
            assert abs(a.cutoff_angle - b.cutoff_angle) < 1e-15 and abs(a.beam_width - b.beam_width) < 1e-15
    for i in range(desc.n_shapes):
        a, b = desc.shapes[i], got.shapes[i]
        assert (a.type, a.emitter, a.tri_count, a.has_vertex_normals) == (b.type, b.emitter, b.tri_count, b.has_vertex_normals), i
        if a.type == scenes.SHAPE_RECTANGLE:
            np.testing.assert_array_equal(_arr(a.to_world), _arr(b.to_world))
            np.testing.assert_allclose(_arr(a.to_object), _arr(b.to_object), rtol=1e-12, atol=1e-15)
        elif a.type == scenes.SHAPE_SPHERE:
            np.testing.assert_array_equal(_arr(a.center), _arr(b.center))
            assert a.radius == b.radius and a.flip_normals == b.flip_normals
        ma, mb = desc.materials[a.material], got.materials[b.material]
        assert (ma.type, ma.twosided, ma.nonlinear) == (mb.type, mb.twosided, mb.nonlinear), (i, ma.type, mb.type)
        if ma.type in (scenes.BSDF_ROUGHCONDUCTOR, scenes.BSDF_ROUGHDIELECTRIC):
            assert ma.alpha == mb.alpha and ma.distribution == mb.distribution
        if ma.type in (scenes.BSDF_ROUGHCONDUCTOR, scenes.BSDF_CONDUCTOR):
            np.testing.assert_array_equal(_arr(ma.eta), _arr(mb.eta))
            np.testing.assert_array_equal(_arr(ma.k), _arr(mb.k))
        if ma.type in (scenes.BSDF_DIELECTRIC, scenes.BSDF_ROUGHDIELECTRIC, scenes.BSDF_PLASTIC):
            assert ma.ior_ratio == mb.ior_ratio
        if ma.type in (scenes.BSDF_DIFFUSE, scenes.BSDF_PLASTIC):
            np.testing.assert_array_equal(_arr(ma.reflectance), _arr(mb.reflectance))
    if desc.envmap:
        e0, e1 = desc.envmap.contents, got.envmap.contents
        assert (e0.width, e0.height, e0.scale, e0.bsphere_radius) == (e1.width, e1.height, e1.scale, e1.bsphere_radius)
        np.testing.assert_array_equal(_arr(e0.bsphere_center), _arr(e1.bsphere_center))          # the kd-tree's enlarged box + the sensor, scene.cpp:386-396
        np.testing.assert_array_equal(_arr(e0.to_world), _arr(e1.to_world))
        np.testing.assert_array_equal(np.ctypeslib.as_array(e0.rgb, shape=(e0.height, e0.width, 3)),
                                      np.ctypeslib.as_array(e1.rgb, shape=(e1.height, e1.width, 3)))      # half-precision texels, envmap.cpp:102-103
    # the same picture: both descriptions through the CPU restatement (vertex tables may be laid out differently)
    ref, _, c1 = oracle.gpt(desc, prm, threads=1)
    out, _, c2 = oracle.gpt(got, prm, threads=1)
    assert c1[0] == c2[0]
    for k in ref:
        scale = max(float(np.abs(ref[k]).mean()), 1e-12)
        assert np.abs(out[k] - ref[k]).max() <= 1e-9 * scale, (name, k)


def test_plugin_render_fails_loudly_without_a_gpu(tmp_path):
    """The same harness on a box without a GPU: scene, scheduler resources and render() are set up as on the GPU box, and the
    plugin turns the library's "no CUDA device" status into Mitsuba's Log(EError) exception -- no CPU fallback, no silent
    success."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    if not os.path.exists(LIB):
        pytest.skip("needs oracle/_ref/libref_plugin_roundtrip.so (a build of /root/reference)")
    lib = ctypes.CDLL(LIB)
    lib.gdbref_roundtrip_last_error.restype = ctypes.c_char_p
    desc = scenes.cbox_glossy(16, 12)
    prm = scenes.default_params(spp=2, seed=1)
    fov, rfilter = scenes.mitsuba_sensor_args(desc)
    rc = lib.gdbref_plugin_render(ctypes.byref(desc), ctypes.byref(prm), ctypes.c_double(fov), rfilter.encode(), 0, 1,
                                  ctypes.c_double(0.2), str(tmp_path / "out").encode())
    assert rc != 0
    msg = lib.gdbref_roundtrip_last_error().decode()
    assert "gdb200" in msg and ("CUDA" in msg or "device" in msg), msg
    assert not list(tmp_path.iterdir())


# ------------------------------------------------------------------ render() of the plugin under Mitsuba's own host objects
@pytest.mark.gpu
@pytest.mark.parametrize("scene_name,recon", [("cbox_glossy", "L2"), ("cbox_materials", "L1"), ("cbox_env", None), ("cbox_mesh_lights", "L2")])
def test_plugin_render_through_mitsuba_host_objects(oracle, tmp_path, scene_name, recon):
    """GDB200GradientPathIntegrator::render (plugin/gpt_plugin.cpp) called the way Mitsuba's RenderJob calls an integrator:
    a real Scene with sensor, MultiFilm and the gdb200_counter sampler (the reference's own classes, built by
    oracle/_ref/libref_mitsuba.so), registered as Scheduler resources, then MultiFilm::develop writing the five PFM files
    (oracle/ref_plugin_roundtrip.cpp: gdbref_plugin_render).  Nothing of the Python API of this package is on that path; the
    files are compared with the STOCK reference `gpt` integrator on the same Scene and sampler, and "-final" with the reference
    solver on the reference's buffers."""
    from conftest import RefMitsuba
    from gdb200 import pfm
    if not os.path.exists(LIB) or not os.path.exists(RefMitsuba.PATH):
        pytest.skip("needs oracle/_ref/libref_plugin_roundtrip.so and libref_mitsuba.so (builds of /root/reference)")
    lib = ctypes.CDLL(LIB)
    lib.gdbref_roundtrip_last_error.restype = ctypes.c_char_p
    w, h, spp = 48, 40, 6
    desc = getattr(scenes, scene_name)(w, h)
    prm = scenes.default_params(spp=spp, seed=11, ref_uninit_measure=True)
    fov, rfilter = scenes.mitsuba_sensor_args(desc)
    dest = str(tmp_path / "out")
    rc = lib.gdbref_plugin_render(ctypes.byref(desc), ctypes.byref(prm), ctypes.c_double(fov), rfilter.encode(),
                                  int(recon == "L1"), int(recon == "L2"), ctypes.c_double(0.2), dest.encode())
    assert rc == 0, lib.gdbref_roundtrip_last_error().decode()
    got = pfm.load_multifilm(dest)                                  # <dest>-final.pfm, -throughput, -dx, -dy, -direct
    ref = RefMitsuba().gpt(desc, prm, threads=2)                    # the reference's own gpt.cpp on the same scene bytes
    for name in ("-throughput", "-dx", "-dy", "-direct"):
        r32 = ref[name].astype(np.float32)
        scale = max(float(np.abs(r32).mean()), 1e-12)
        assert np.abs(got[name] - r32).max() <= 2e-6 * max(scale, float(np.abs(r32).max())), name      # PFM holds float32
    if recon:
        f32 = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in ref.items()}
        want = oracle.poisson(f32["-dx"], f32["-dy"], f32["-throughput"], f32["-direct"], alpha=0.2, preset=recon + "D")
        exact = oracle.poisson_acc64(f32["-dx"], f32["-dy"], f32["-throughput"], f32["-direct"], alpha=0.2, preset=recon + "D")
        floor = float(np.sqrt(np.mean((want.astype(np.float64) - exact) ** 2)))
        err = float(np.sqrt(np.mean((got["-final"].astype(np.float64) - want) ** 2)))
        assert err <= max(1e-5, 2.0 * floor), (err, floor)
    else:                                                           # no reconstruction: "-final" is the preview accumulated by the tracer
        r32 = ref["-final"].astype(np.float32)
        assert np.abs(got["-final"] - r32).max() <= 2e-6 * float(np.abs(r32).max())
