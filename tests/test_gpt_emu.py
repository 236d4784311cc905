"""The tracer's device routines (gpt_device.cuh / gpt_kernels.cuh), compiled for the host through
tests/emu/cuda_emu.h, against the fp64 CPU oracle on identical scene bytes and sample streams.

This is the CPU-side net under the GPU parity tests (tests/test_gpt_gpu.py): same source as the CUDA
kernels' per-slot bodies, same libm as the oracle, so agreement is expected to the last few ulps
(sum-order of the film accumulation only).  The wavefront scheduling (queues, compaction) is not
covered here; it needs the GPU."""
import math
import os

import numpy as np
import pytest

from gdb200 import scenes

TOL = 1e-12


def close(got, ref):
    for name in ("-throughput", "-dx", "-dy", "-direct", "-final"):
        scale = max(float(np.abs(ref[name]).mean()), 1e-12)
        assert np.abs(got[name] - ref[name]).max() <= TOL * max(scale, 1.0) * 100, name


SCENES = {"cbox_diffuse": lambda w, h: scenes.cbox_diffuse(w, h), "cbox_glossy": lambda w, h: scenes.cbox_glossy(w, h),
          "cbox_glossy_delta": lambda w, h: scenes.cbox_glossy(w, h, delta_variant=True),
          "cbox_materials": lambda w, h: scenes.cbox_materials(w, h),
          "cbox_env": lambda w, h: scenes.cbox_env(w, h),                  # environment emitter + environmentShift
          "cbox_mesh_lights": lambda w, h: scenes.cbox_mesh_lights(w, h),  # mesh emitters, plastic, twosided
          "atrium": lambda w, h: scenes.atrium(w, h, columns=3, segments=8, rings=4),   # > table size: BVH path
          "cbox_point": lambda w, h: scenes.cbox_point(w, h),               # point emitter: the EDiscrete branches of the NEE shift
          "cbox_spot": lambda w, h: scenes.cbox_spot(w, h),                 # spot emitters: cone falloff, only Dirac lights in the scene
          "cbox_dof": lambda w, h: scenes.cbox_dof(w, h),                   # thinlens sensor: aperture samples
          "cbox_sphere_lights": lambda w, h: scenes.cbox_sphere_lights(w, h),   # sphere area emitters (cone / uniform-sphere sampling)
          "cbox_roughglass": lambda w, h: scenes.cbox_roughglass(w, h),     # roughdielectric: refraction half-vector Jacobian, in-BSDF sampler draw
          "cbox_smooth": lambda w, h: scenes.cbox_smooth(w, h)}             # vertex normals (shading != geometric normal), smooth mesh emitter


@pytest.mark.parametrize("scene_name", sorted(SCENES))
def test_device_routines_match_oracle(oracle, emu, scene_name):
    desc = SCENES[scene_name](40, 32)
    p = scenes.default_params(spp=6, seed=3)
    got, cnt = emu.gpt(desc, p)
    ref, _, c2 = oracle.gpt(desc, p)
    close(got, ref)
    assert cnt[3] == c2[0] == 40 * 32 * 6 and cnt[1] == c2[1] and cnt[2] == c2[2]     # samples, rays, path vertices


@pytest.mark.parametrize("kw", [dict(max_depth=2), dict(max_depth=1), dict(rr_depth=2), dict(strict_normals=True),
                                dict(shift_threshold=0.1)])
def test_device_routines_parameters(oracle, emu, kw):
    desc = scenes.cbox_glossy(24, 24)
    p = scenes.default_params(spp=5, seed=11, **kw)
    got, _ = emu.gpt(desc, p)
    ref, _, _ = oracle.gpt(desc, p)
    close(got, ref)


@pytest.mark.parametrize("streams,spp", [(2, 8), (3, 7), (8, 5)])
def test_streams_per_pixel(oracle, emu, streams, spp):
    """Chunked sample streams (gdb200_gpt_params.streams_per_pixel): same film as the oracle's chunk loop,
    every sample accounted for, also when fewer slots than streams are resident (streams dealt dynamically)."""
    desc = scenes.cbox_glossy(20, 16)
    p = scenes.default_params(spp=spp, seed=5)
    p.streams_per_pixel = streams
    ref, _, c2 = oracle.gpt(desc, p)
    for cap in (None, 97):
        if cap is None:
            os.environ.pop("GDB200_MAX_SLOTS", None)
        else:
            os.environ["GDB200_MAX_SLOTS"] = str(cap)
        try:
            got, cnt = emu.gpt(desc, p)
        finally:
            os.environ.pop("GDB200_MAX_SLOTS", None)
        close(got, ref)
        assert cnt[3] == 20 * 16 * spp == c2[0]
        assert cnt[0] == (97 if cap else 20 * 16 * streams)
    one = scenes.default_params(spp=spp, seed=5)
    ref1, _, _ = oracle.gpt(desc, one)
    assert not np.allclose(ref1["-throughput"], ref["-throughput"])     # chunks > 0 really use other streams


@pytest.mark.parametrize("scene_name", ["cbox_glossy", "cbox_mesh_lights"])
def test_bvh_path_gives_the_same_film(oracle, emu, scene_name, monkeypatch):
    """GDB200_FORCE_BVH routes every mesh triangle through the BVH instead of the constant-memory table; the
    oracle tests every triangle, so any traversal or bounds error shows up as a changed hit."""
    monkeypatch.setenv("GDB200_FORCE_BVH", "1")
    desc = SCENES[scene_name](36, 28)
    p = scenes.default_params(spp=5, seed=8)
    got, cnt = emu.gpt(desc, p)
    ref, _, c2 = oracle.gpt(desc, p)
    close(got, ref)
    assert cnt[1] == c2[1] and cnt[2] == c2[2]


@pytest.mark.parametrize("kw", [dict(strict_normals=True), dict(strict_normals=True, max_depth=3), dict(strict_normals=True, shift_threshold=0.2)])
def test_strict_normals_with_shading_normals(oracle, emu, kw):
    """strictNormals only bites when shading and geometric normals differ (smooth-shaded meshes)."""
    desc = scenes.cbox_smooth(30, 26)
    p = scenes.default_params(spp=5, seed=6, **kw)
    got, _ = emu.gpt(desc, p)
    ref, _, _ = oracle.gpt(desc, p)
    close(got, ref)
    loose, _, _ = oracle.gpt(desc, scenes.default_params(spp=5, seed=6, **{**kw, "strict_normals": False}))
    assert np.abs(loose["-throughput"] - ref["-throughput"]).max() > 1e-6


@pytest.mark.parametrize("kw", [dict(max_depth=2), dict(max_depth=3, rr_depth=1), dict(strict_normals=True), dict(shift_threshold=0.1)])
def test_environment_scene_parameters(oracle, emu, kw):
    desc = scenes.cbox_env(28, 24)
    p = scenes.default_params(spp=5, seed=2, **kw)
    got, _ = emu.gpt(desc, p)
    ref, _, _ = oracle.gpt(desc, p)
    close(got, ref)


def test_unsupported_scenes_fail_loudly(emu):
    b = scenes._cornell(8, 8)
    b.shapes[b.sphere((0, 0, 0), 0.2, 0)].emitter = 0
    with pytest.raises(RuntimeError, match="belongs to another shape"):
        emu.gpt(b.build(), scenes.default_params(spp=1))
    b = scenes._cornell(8, 8)
    b.material(type=scenes.BSDF_DIELECTRIC, twosided=True)
    with pytest.raises(RuntimeError, match="transmission component"):
        emu.gpt(b.build(), scenes.default_params(spp=1))
    b = scenes._cornell(8, 8)
    b.envmap(scenes.sky_envmap(8, 4) * 0)
    with pytest.raises(RuntimeError, match="completely black"):
        emu.gpt(b.build(), scenes.default_params(spp=1))


@pytest.mark.parametrize("rfilter", ["gaussian", "tent"])
def test_reconstruction_filters(oracle, emu, rfilter):
    """Film filters other than box (gaussian is Mitsuba's default, film.cpp:89-95): the discretised table of
    rfilter.cpp:37-55 drives the splat footprint and weights of ImageBlock::put."""
    desc = scenes.cbox_diffuse(26, 22, rfilter=rfilter)
    p = scenes.default_params(spp=5, seed=4)
    got, _ = emu.gpt(desc, p)
    ref, wts, _ = oracle.gpt(desc, p)
    close(got, ref)
    box, _, _ = oracle.gpt(scenes.cbox_diffuse(26, 22), p)
    assert np.abs(box["-throughput"] - ref["-throughput"]).max() > 1e-4          # the filter matters
    assert abs(box["-throughput"].mean() - ref["-throughput"].mean()) < 0.05 * box["-throughput"].mean()   # but keeps the energy
    one, _, _ = oracle.gpt(desc, p, threads=1)                                  # oracle band parallelism is safe for the wider footprint
    for k in ref:
        assert np.array_equal(one[k], ref[k])


@pytest.mark.parametrize("scene_name,no_tail,cap", [("cbox_glossy", False, None), ("cbox_mesh_lights", True, None),
                                                    ("cbox_env", True, 100), ("cbox_point", True, None), ("cbox_spot", False, None),
                                                    ("cbox_roughglass", True, None)])
def test_queued_wavefront_kernels(oracle, emu, scene_name, no_tail, cap, monkeypatch):
    """Block mode of the emulation: gpt_generate_kernel / gpt_compact_kernel / gpt_bounce_kernel<2> / gpt_tail_kernel as
    written (shared-memory prologues, __syncthreads, ballots, queue appends), CTAs run by OS threads, with the step loop of
    gdb200_gpt_render restated around them.  Every queue entry is checked each step; GDB200_NO_TAIL keeps all paths on the
    queues to the end (every BSDF-type x shift-stage bucket, incl. plastic), a slot cap deals streams dynamically."""
    if no_tail:
        monkeypatch.setenv("GDB200_NO_TAIL", "1")
    if cap:
        monkeypatch.setenv("GDB200_MAX_SLOTS", str(cap))
    desc = SCENES[scene_name](14, 10)
    p = scenes.default_params(spp=3, seed=3)
    p.streams_per_pixel = 2
    got, cnt = emu.gpt_wavefront(desc, p)
    ref, _, c2 = oracle.gpt(desc, p)
    close(got, ref)
    assert cnt[3] == c2[0] == 14 * 10 * 3 and cnt[1] == c2[1] and cnt[2] == c2[2]
    assert cnt[0] == (cap if cap else 14 * 10 * 2)


@pytest.mark.parametrize("scene_name", sorted(SCENES))
def test_staged_wavefront_kernels(oracle, emu, scene_name):
    """Block mode of the emulation for the STAGED wavefront (csrc/gpt_stages.cuh, the product's default path): stage
    compaction, primary / shade<0..2> / resolve / prepare / generate and the two cast kernels as written (persistent CTAs
    over their queues, warp-aggregated ray appends), the tick loop of renderStaged() restated around them.  Every queue
    entry is checked against the status it must have; the film, the ray count and the path-vertex count must be the
    oracle's."""
    w, h = (12, 8) if scene_name == "atrium" else (14, 10)
    desc = SCENES[scene_name](w, h)
    p = scenes.default_params(spp=3, seed=3)
    p.streams_per_pixel = 2
    got, cnt = emu.gpt_staged(desc, p)
    ref, _, c2 = oracle.gpt(desc, p)
    close(got, ref)
    assert cnt[3] == c2[0] == w * h * 3 and cnt[1] == c2[1] and cnt[2] == c2[2]
    assert cnt[0] == w * h * 2


@pytest.mark.parametrize("scene_name,kw", [("cbox_glossy", dict(max_depth=2)), ("cbox_glossy", dict(max_depth=1)), ("cbox_glossy", dict(rr_depth=2)),
                                           ("cbox_glossy", dict(strict_normals=True)), ("cbox_glossy", dict(shift_threshold=0.1)),
                                           ("cbox_smooth", dict(strict_normals=True)), ("cbox_smooth", dict(strict_normals=True, shift_threshold=0.2)),
                                           ("cbox_env", dict(max_depth=3, rr_depth=1)), ("cbox_env", dict(strict_normals=True)),
                                           ("cbox_materials", dict(shift_threshold=0.1, ref_uninit_measure=True))])
def test_staged_wavefront_parameters(oracle, emu, scene_name, kw):
    desc = SCENES[scene_name](14, 10)
    p = scenes.default_params(spp=4, seed=9, **kw)
    got, cnt = emu.gpt_staged(desc, p, grid=2)
    ref, _, c2 = oracle.gpt(desc, p)
    close(got, ref)
    assert cnt[1] == c2[1] and cnt[2] == c2[2]


@pytest.mark.parametrize("cap,grid", [(37, 1), (100, 5)])
def test_staged_wavefront_deals_streams_to_few_slots(oracle, emu, cap, grid):
    """max_slots below the stream count: slots take the next stream from the atomic counter as they drain."""
    desc = scenes.cbox_glossy(14, 10)
    p = scenes.default_params(spp=5, seed=4, max_slots=cap)
    p.streams_per_pixel = 3
    got, cnt = emu.gpt_staged(desc, p, grid=grid)
    ref, _, c2 = oracle.gpt(desc, p)
    close(got, ref)
    assert cnt[3] == c2[0] == 14 * 10 * 5 and cnt[0] == cap


@pytest.mark.parametrize("scene_name", ["cbox_glossy", "cbox_diffuse", "cbox_mesh_lights", "cbox_smooth", "atrium", "cbox_sphere_lights"])
def test_candidate_selection_never_changes_a_hit_on_the_host(emu, scene_name):
    """gpt_check_culling_kernel on the host: the bounds-based candidate pass of closestPrimitive (table primitives and the
    BVH walk) must give the answers of testing every primitive -- nearest hit and occlusion, for extension,
    visibility-segment and camera-like rays."""
    desc = SCENES[scene_name](32, 24)
    for seed in range(2):
        bad, hits = emu.check_culling(desc, 150_000, seed)
        assert bad == 0, (scene_name, seed, bad)
        assert hits > 50_000


def test_host_validation_follows_the_reference(emu):
    """Scene flattening (csrc/gpt_host.h, shared by the library and the emulation) rejects what the reference's plugins reject."""
    cam = scenes.make_camera(8, 8, (0, 0, 4), (0, 0, 0), (0, 1, 0), 40)
    b = scenes.SceneBuilder(cam)
    b.rectangle((0, 0, 0), (1, 0.2, 0), (0, 1, 0), b.material())          # s . t != 0
    b.point_light((0, 0, 2), (1, 1, 1))
    with pytest.raises(RuntimeError, match="contains shear"):            # rectangle.cpp:108-109
        emu.gpt(b.build(), scenes.default_params(spp=1))
    b = scenes.SceneBuilder(cam)
    b.rectangle((0, 0, 0), (1, 0, 0), (0, 1, 0), b.material())
    e = b.spot_light(scenes.look_at((0, 0, 2), (0, 0, 0), (0, 1, 0)), (1, 1, 1), cutoff_angle=20.0, beam_width=10.0)
    b.emitters[e].beam_width = math.radians(30.0)                         # beamWidth > cutoffAngle: the Assert of spot.cpp:75
    with pytest.raises(RuntimeError, match="cutoffAngle"):
        emu.gpt(b.build(), scenes.default_params(spp=1))
