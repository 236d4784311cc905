#!/usr/bin/env python
"""Regenerates tests/golden/ref_gpt_golden.npz and ref_bsdf_golden.npz: the outputs of the REFERENCE's own gpt.cpp and BSDF
plugins (compiled from /root/reference into oracle/_ref by `make -C oracle ref`) for the cases listed in
tests/test_ref_gpt.py and tests/test_ref_mitsuba.py.  Needs /root/reference (to build) or a built oracle/_ref/libref_mitsuba.so.

    python tests/golden/make_ref_golden.py

The generators live in the test modules (their `reference` / `reference_outputs` fixtures write the files when
GDB200_WRITE_GOLDEN is set); this script just runs them that way and reports what was written."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
env = dict(os.environ, GDB200_WRITE_GOLDEN="1")
rc = subprocess.call([sys.executable, "-m", "pytest", "-q", "tests/test_ref_gpt.py", "tests/test_ref_mitsuba.py"], cwd=ROOT, env=env)
for name in ("ref_gpt_golden.npz", "ref_bsdf_golden.npz"):
    path = os.path.join(ROOT, "tests", "golden", name)
    print(name, os.path.getsize(path), "bytes")
sys.exit(rc)
