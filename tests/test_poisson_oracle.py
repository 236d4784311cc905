"""CPU tests: pin the Poisson oracle (oracle/poisson_oracle.c) against
(a) the reference's own solver sources compiled unmodified (oracle/_ref) and
(b) golden outputs of that reference committed under tests/golden/."""
import hashlib
import json
import os

import numpy as np
import pytest

import gdb200  # noqa: F401
from gdb200 import synth
from conftest import ROOT, rmse

GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "poisson_golden.json")))


def _inputs(case):
    d = synth.solver_inputs(case["w"], case["h"], seed=case["seed"], last_col_nonzero=case["last_col_nonzero"])
    sha = hashlib.sha256(b"".join(d[k].tobytes() for k in ("throughput", "dx", "dy", "direct"))).hexdigest()
    return d, sha == case["input_sha256"]


@pytest.mark.parametrize("case", GOLDEN, ids=lambda c: f"{c['w']}x{c['h']}-{c['preset']}-d{int(c['direct'])}")
def test_oracle_matches_reference_golden(oracle, case):
    d, same_bytes = _inputs(case)
    fin = oracle.poisson(d["dx"], d["dy"], d["throughput"], d["direct"] if case["direct"] else None,
                         alpha=case["alpha"], preset=case["preset"])
    if same_bytes:   # identical input bytes => the restatement must be bit-identical to the reference
        assert hashlib.sha256(fin.tobytes()).hexdigest() == case["final_sha256"]
    probe = fin.reshape(-1)[:: max(1, fin.size // 8)][:8]
    np.testing.assert_allclose(probe, case["final_probe"], rtol=0, atol=2e-5)
    assert abs(float(fin.astype(np.float64).mean()) - case["final_mean"]) < 1e-5


@pytest.mark.parametrize("size", [(64, 48), (33, 17), (5, 1), (1, 1)])
@pytest.mark.parametrize("preset", ["L2D", "L1D", "L2Q"])
def test_oracle_bit_exact_vs_compiled_reference(oracle, size, preset):
    if oracle.ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    d = synth.solver_inputs(*size, seed=42, last_col_nonzero=True)
    for direct in (d["direct"], None):
        a = oracle.poisson(d["dx"], d["dy"], d["throughput"], direct, preset=preset)
        b = oracle.poisson_ref(d["dx"], d["dy"], d["throughput"], direct, preset=preset)
        assert np.array_equal(a, b)


def test_oracle_no_throughput_matches_reference(oracle):
    if oracle.ref is None:
        pytest.skip("oracle/_ref not built")
    d = synth.solver_inputs(40, 24, seed=3)
    a = oracle.poisson(d["dx"], d["dy"], None, None, preset="L2D")
    b = oracle.poisson_ref(d["dx"], d["dy"], None, None, preset="L2D")
    assert np.array_equal(a, b)


@pytest.mark.parametrize("preset", ["L2D", "L1D"])
def test_oracle_fixed_point_is_bit_exact(oracle, preset):
    """dx,dy = forward differences of throughput => e == 0, CG breaks at iteration 0 and
    final == throughput + direct bit-exactly (Solver.cpp:440)."""
    d = synth.fixed_point_inputs(96, 40)
    fin = oracle.poisson(d["dx"], d["dy"], d["throughput"], d["direct"], alpha=0.2, preset=preset)
    assert np.array_equal(fin, 1.0 * d["direct"] + d["throughput"])


def test_oracle_denoises(oracle):
    d = synth.solver_inputs(128, 128, seed=1)
    for preset in ("L2D", "L1D"):
        fin = oracle.poisson(d["dx"], d["dy"], d["throughput"], None, preset=preset)
        assert rmse(fin, d["clean"]) < 0.25 * rmse(d["throughput"], d["clean"])


@pytest.mark.parametrize("preset", ["L1Q", "L1L"])
def test_oracle_bit_exact_vs_compiled_reference_slow_presets(oracle, preset):
    """The two remaining presets of Solver::Params::setConfigPreset (Solver.cpp:117-147): L1Q = 64 IRLS x 1000 CG,
    L1L = 7 IRLS x 20000 CG with cgTolerance 1e-20 — pinned on a small image where they finish in a second."""
    if oracle.ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    d = synth.solver_inputs(24, 16, seed=7, last_col_nonzero=True)
    a = oracle.poisson(d["dx"], d["dy"], d["throughput"], d["direct"], preset=preset)
    b = oracle.poisson_ref(d["dx"], d["dy"], d["throughput"], d["direct"], preset=preset)
    assert np.array_equal(a, b)
    assert rmse(a, d["clean"]) < rmse(d["throughput"], d["clean"])


@pytest.mark.parametrize("preset", ["L2D", "L1D"])
def test_metrics_restatement_matches_the_reference(oracle, preset):
    """Solver::evaluateMetricsMTS (Solver.cpp:511-541): the restatement against the reference's own method, bit for bit."""
    import ctypes
    if oracle.ref is None:
        pytest.skip("needs oracle/_ref/libref_poisson.so")
    w, h = 40, 28
    d = synth.solver_inputs(w, h, seed=21)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    res = {}
    for name, fn in (("port", oracle.lib.gdb200_oracle_poisson_metrics), ("ref", oracle.ref.ref_poisson_metrics)):
        final, err, errL = np.empty_like(d["dx"]), np.empty_like(d["dx"]), (ctypes.c_float * 2)()
        assert fn(p(d["dx"]), p(d["dy"]), p(d["throughput"]), p(d["direct"]), w, h, ctypes.c_float(0.2), preset.encode(), p(final), p(err), errL) == 0
        res[name] = (final, err, (errL[0], errL[1]))
    assert np.array_equal(res["port"][0], res["ref"][0]) and np.array_equal(res["port"][1], res["ref"][1])
    assert res["port"][2] == res["ref"][2] and res["ref"][2][0] > 0
