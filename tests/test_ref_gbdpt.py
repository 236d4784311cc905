"""The pin for the G-BDPT path (BASELINE config 4): outputs of the REFERENCE's own G-BDPT integrator
(src/integrators/gbdpt/{gbdpt,gbdpt_proc,gbdpt_wr}.cpp over src/libbidir, compiled unmodified into
oracle/_ref/libref_gbdpt.so) on scenes built from a gdb200_scene_desc, rendered through a real RenderJob with the
gdb200_counter sampler (oracle/ref_gbdpt_shim.cpp).  tests/golden/ref_gbdpt_golden.npz holds them for boxes without the build
(generator: tests/golden/make_gbdpt_golden.py).  No restatement and no kernel of G-BDPT exist yet; these are the vectors they
will be held to: the seven buffers MultiFilm writes (gbdpt.cpp:164) -- the base image, the four signed finite-difference
gradients and both reconstructions."""
import os

import numpy as np
import pytest

import gdb200  # noqa: F401
from gdb200 import scenes
from conftest import ROOT, RefGbdpt

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_gbdpt_golden.npz")
W, H, SPP, SEED = 20, 16, 3, 5
CASES = {
    "cbox_glossy": dict(max_depth=6),
    "cbox_glossy_no_light_image": dict(scene="cbox_glossy", max_depth=6, light_image=False),
    "cbox_diffuse_depth4": dict(scene="cbox_diffuse", max_depth=4),
    "cbox_materials": dict(max_depth=8),                               # conductor, dielectric, plastic: manifold offsets through specular chains
    "cbox_glossy_delta": dict(scene="cbox_glossy", scene_kw=dict(delta_variant=True), max_depth=6),
    "cbox_sphere_lights_rr2": dict(scene="cbox_sphere_lights", max_depth=5, rr_depth=2),
    "cbox_glossy_threshold": dict(scene="cbox_glossy", max_depth=6, shift_threshold=0.1),
}


def case(name):
    kw = dict(CASES[name])
    light = kw.pop("light_image", True)
    desc = getattr(scenes, kw.pop("scene", name))(W, H, **kw.pop("scene_kw", {}))
    return desc, scenes.default_params(spp=SPP, seed=SEED, **kw), light


def _have_reference():
    return RefGbdpt.available() and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_mitsuba.so")) and not os.environ.get("GDB200_NO_REF")


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_gbdpt_reproduces_the_committed_vectors(name):
    """The compiled reference G-BDPT gives the committed outputs bit for bit (float32 PFM values): the build recipe, the
    driver and the sampler are deterministic -- any thread count, any run."""
    golden = np.load(GOLDEN)
    for k in RefGbdpt.NAMES:
        g = golden[name + k]
        assert g.shape == (H, W, 3) and np.isfinite(g).all(), k
    assert golden[name + "-primal"].mean() > 0.01
    if not _have_reference():
        pytest.skip("needs oracle/_ref/libref_gbdpt.so (a build of /root/reference)")
    desc, prm, light = case(name)
    got = RefGbdpt().render(desc, prm, light_image=light, threads=3)
    for k in RefGbdpt.NAMES:
        assert np.array_equal(got[k], golden[name + k]), (name, k, float(np.abs(got[k] - golden[name + k]).max()))


def test_gbdpt_and_gpt_estimate_the_same_image(oracle):
    """Links the G-BDPT pin to the G-PT path that is already pinned: both integrators are unbiased estimators of the same
    image, so at a moderate sample count the G-BDPT base image and G-PT's throughput + direct agree in the mean, and so do
    the finite-difference gradients as the solver sees them."""
    if not _have_reference():
        pytest.skip("needs oracle/_ref/libref_gbdpt.so (a build of /root/reference)")
    desc = scenes.cbox_diffuse(W, H)
    prm = scenes.default_params(spp=96, seed=2, max_depth=5)
    bd = RefGbdpt().render(desc, prm, threads=4)
    pt, _, _ = oracle.gpt(desc, prm)
    primal_pt = pt["-throughput"] + pt["-direct"]
    assert abs(bd["-primal"].mean() - primal_pt.mean()) <= 0.05 * primal_pt.mean(), (bd["-primal"].mean(), primal_pt.mean())
    # The solver's dx is merged from two one-sided buffers the way gbdpt.cpp:264-280 does it: half the +x gradient of pixel
    # (x, y) minus half the -x gradient of pixel (x+1, y).  G-PT's "-dx" (gpt.cpp:1338-1345) estimates the same difference.
    dx_bd = 0.5 * bd["-gradientPosX"][:, :-1].astype(np.float64) - 0.5 * bd["-gradientNegX"][:, 1:]
    dy_bd = 0.5 * bd["-gradientPosY"][:-1].astype(np.float64) - 0.5 * bd["-gradientNegY"][1:]
    # (G-PT keeps directly visible emitters out of its gradients and adds "-direct" after the reconstruction, gpt.cpp:1445-1462;
    # G-BDPT's buffers hold the whole image, so the differences of "-direct" are added to G-PT's side)
    direct = pt["-direct"]
    dx_pt = pt["-dx"][:, :-1] + (direct[:, 1:] - direct[:, :-1])
    dy_pt = pt["-dy"][:-1] + (direct[1:] - direct[:-1])
    for g_bd, g_pt in ((dx_bd, dx_pt), (dy_bd, dy_pt)):
        noise = float(np.abs(g_pt).mean())
        assert np.abs(g_bd.mean(axis=(0, 1)) - g_pt.mean(axis=(0, 1))).max() <= 0.02 * primal_pt.mean() + 0.25 * noise
        assert np.corrcoef(g_bd.ravel(), g_pt.ravel())[0, 1] > 0.8            # the same edges, pixel by pixel
    # both reconstructions stay close to the base image in the mean (screened Poisson keeps the DC level of the primal)
    assert abs(bd["-L2"].mean() - bd["-primal"].mean()) <= 0.05 * bd["-primal"].mean()
    assert abs(bd["-L1"].mean() - bd["-primal"].mean()) <= 0.10 * bd["-primal"].mean()
