"""Committed golden buffers of the G-PT tracer (tests/golden/gpt_golden.npz, generator make_gpt_golden.py): the oracle
and the device routines (host build) must keep reproducing them; tests/test_gpt_gpu.py checks the GPU against the same file."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_gpt_golden import SCENES, params  # noqa: E402

GOLDEN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gpt_golden.npz"))
BUFFERS = ("-throughput", "-dx", "-dy", "-direct", "-final")


def check(got, name, rel=1e-10):
    for b in BUFFERS:
        ref = GOLDEN[name + b]
        scale = max(float(np.abs(ref).mean()), 1e-12)
        assert np.abs(got[b] - ref).max() <= rel * max(scale, 1.0), (name, b, float(np.abs(got[b] - ref).max()))


@pytest.mark.parametrize("name", sorted(SCENES))
def test_oracle_reproduces_golden(oracle, name):
    got, _, cnt = oracle.gpt(SCENES[name](), params(), threads=3)
    check(got, name)
    assert np.array_equal(cnt, GOLDEN[name + "/counters"])          # samples, rays, path vertices


@pytest.mark.parametrize("name", sorted(SCENES))
def test_device_routines_reproduce_golden(emu, name):
    got, _ = emu.gpt(SCENES[name](), params())
    check(got, name)
