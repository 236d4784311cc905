"""Generates tests/golden/ref_gbdpt_golden.npz: outputs of the REFERENCE's own G-BDPT integrator (compiled from /root/reference
into oracle/_ref/libref_gbdpt.so, driven by oracle/ref_gbdpt_shim.cpp) for the cases of tests/test_ref_gbdpt.py.  Run in a
container that has /root/reference:  python tests/golden/make_gbdpt_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from conftest import RefGbdpt  # noqa: E402
import test_ref_gbdpt as T  # noqa: E402

ref = RefGbdpt()
out = {}
for name in sorted(T.CASES):
    desc, prm, light = T.case(name)
    got = ref.render(desc, prm, light_image=light)
    for k, v in got.items():
        out[name + k] = v
np.savez_compressed(os.path.join(HERE, "ref_gbdpt_golden.npz"), **out)
print("wrote", len(out), "arrays")
