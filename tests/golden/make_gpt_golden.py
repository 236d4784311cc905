"""Generates tests/golden/gpt_golden.npz: the five G-PT buffers of the CPU oracle (oracle/gpt_oracle.cpp) for every test
scene at 20x16, 4 spp, seed 21, 2 sample streams per pixel.  The reference ships no golden image for `gpt`,
so these fixtures pin the oracle against *itself* over time (any later edit that changes its output shows up in
tests/test_gpt_golden.py) and give the GPU test a committed target that does not depend on rebuilding the oracle.

    python tests/golden/make_gpt_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gdb200  # noqa: E402,F401
from gdb200 import scenes  # noqa: E402
from conftest import Oracle  # noqa: E402

W, H, SPP, SEED, STREAMS = 20, 16, 4, 21, 2
SCENES = {
    "cbox_diffuse": lambda: scenes.cbox_diffuse(W, H), "cbox_glossy": lambda: scenes.cbox_glossy(W, H),
    "cbox_glossy_delta": lambda: scenes.cbox_glossy(W, H, delta_variant=True), "cbox_materials": lambda: scenes.cbox_materials(W, H),
    "cbox_env": lambda: scenes.cbox_env(W, H), "cbox_mesh_lights": lambda: scenes.cbox_mesh_lights(W, H),
    "cbox_smooth": lambda: scenes.cbox_smooth(W, H), "cbox_point": lambda: scenes.cbox_point(W, H), "cbox_spot": lambda: scenes.cbox_spot(W, H), "cbox_dof": lambda: scenes.cbox_dof(W, H),
    "cbox_roughglass": lambda: scenes.cbox_roughglass(W, H), "cbox_sphere_lights": lambda: scenes.cbox_sphere_lights(W, H),
    "atrium": lambda: scenes.atrium(W, H, columns=3, segments=8, rings=4), "cbox_gaussian": lambda: scenes.cbox_diffuse(W, H, rfilter="gaussian"),
}


def params():
    p = scenes.default_params(spp=SPP, seed=SEED)
    p.streams_per_pixel = STREAMS
    return p


if __name__ == "__main__":
    orc = Oracle()
    out = {}
    for name, mk in SCENES.items():
        bufs, _, cnt = orc.gpt(mk(), params(), threads=4)
        for k, v in bufs.items():
            out[f"{name}{k}"] = v
        out[f"{name}/counters"] = cnt
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gpt_golden.npz"), **out)
    print("wrote", len(out), "arrays")
