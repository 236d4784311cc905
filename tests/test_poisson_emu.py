"""The Poisson solver's DEVICE code (gradientdomain-mitsuba_b200/csrc/poisson.cu) run on the host: tests/emu/poisson_emu.cpp
compiles the same source with CUDA threads as OS threads, CTAs as processes and "GPUs" as groups of CTA processes sharing one
memory mapping.  Checks, without a GPU: the persistent IRLS/CG kernel against the CPU oracle, the four kernel variants against
each other (bit for bit), the sharded protocol (halo rows through peer pointers, reductions through mailboxes and the relay
slot) with one and with several CTAs per GPU, and that a solve whose peer never starts gives up instead of hanging.
The real kernels are covered on the GPU by tests/test_poisson_gpu.py."""
import os
import struct
import subprocess
import time

import numpy as np
import pytest

from gdb200 import synth
from conftest import rmse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "poisson_emu")
MAX_RANKS = 16
L2D = (1, 0.0, 0.0, 50, 100, 0.0)                 # irlsIterMax, irlsRegInit, irlsRegIter, cgIterMax, cgIterCheck, cgTolerance
SHORT_L1 = (3, 0.05, 0.5, 6, 100, 0.0)            # the L1D recipe with fewer iterations (the emulation pays ~100 barriers per CG step)


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu"), "poisson_emu"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)

    def run(d, w, h, cfg, variant=0, bounds=None, ctas=1, skip_rank=-1, tmp=None, timeout=600, direct=True):
        bounds = [0, h] if bounds is None else bounds
        n = len(bounds) - 1
        head = struct.pack("<6i", w, h, variant, n, ctas, skip_rank) + struct.pack(f"<{MAX_RANKS + 1}i", *(bounds + [0] * (MAX_RANKS + 1 - len(bounds))))
        head += struct.pack("<iffiif", *cfg) + struct.pack("<fi", 0.2, 1 if direct else 0)
        src, dst = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        with open(src, "wb") as f:
            f.write(head)
            for k in ("dx", "dy", "throughput", "direct"):
                f.write(np.ascontiguousarray(d[k], dtype=np.float32).tobytes())
        t0 = time.time()
        r = subprocess.run([EMU, src, dst], capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0 and "pthread_create failed" in r.stderr:
            pytest.skip("this machine does not allow the emulation's 256 OS threads per CTA process")
        assert r.returncode == 0, r.stderr
        raw = open(dst, "rb").read()
        img = np.frombuffer(raw, dtype=np.float32, count=w * h * 3).reshape(h, w, 3).copy()
        tail = np.frombuffer(raw, dtype=np.int32, offset=w * h * 3 * 4).reshape(n, 4)         # per shard: irls, cg, status, last message number
        return img, tail, time.time() - t0
    return run


def test_emulated_kernel_matches_the_oracle(emu, oracle, tmp_path):
    w, h = 64, 32
    d = synth.solver_inputs(w, h, seed=4, last_col_nonzero=True)
    got, tail, _ = emu(d, w, h, L2D, tmp=str(tmp_path))
    want = oracle.poisson(d["dx"], d["dy"], d["throughput"], d["direct"], alpha=0.2, preset="L2D")
    assert tuple(tail[0][:2]) == (1, 50)
    assert rmse(got, want) <= 1e-6, rmse(got, want)


@pytest.mark.parametrize("size,ctas", [((64, 32), 1), ((64, 32), 2), ((52, 23), 1)])
def test_emulated_kernel_variants_return_the_same_bits(emu, tmp_path, size, ctas):
    w, h = size
    d = synth.solver_inputs(w, h, seed=9, last_col_nonzero=True)
    base, tail0, _ = emu(d, w, h, SHORT_L1, variant=0, ctas=ctas, tmp=str(tmp_path))
    assert np.isfinite(base).all() and tuple(tail0[0][:2]) == (3, 18)
    for variant in (1, 2, 3):
        got, tail, _ = emu(d, w, h, SHORT_L1, variant=variant, ctas=ctas, tmp=str(tmp_path))
        assert np.array_equal(got, base), variant
        assert tuple(tail[0][:2]) == (3, 18)


@pytest.mark.parametrize("w,bounds,ctas,variant", [(64, [0, 16, 32], 1, 0), (64, [0, 16, 32], 1, 1), (64, [0, 32, 64], 2, 0), (64, [0, 32, 64], 2, 1),
                                                   (64, [0, 16, 32, 41], 1, 0), (64, [0, 16, 32, 48, 64], 1, 2), (100, [0, 16, 35], 2, 3)])
def test_emulated_sharded_solve(emu, tmp_path, w, bounds, ctas, variant):
    """Several "GPUs" (groups of CTA processes) solve one image: equal to the one-GPU solve up to reduction order, every
    shard takes the same number of iterations, and the result does not depend on the kernel variant."""
    h = bounds[-1]
    d = synth.solver_inputs(w, h, seed=12, last_col_nonzero=True)
    single, _, _ = emu(d, w, h, SHORT_L1, variant=0, tmp=str(tmp_path))
    got, tail, _ = emu(d, w, h, SHORT_L1, variant=variant, bounds=bounds, ctas=ctas, tmp=str(tmp_path))
    assert not (got == -777.0).any(), "every band was written"
    assert all(tuple(t[:2]) == (3, 18) and t[2] == 0 for t in tail), tail
    assert len({int(t[3]) for t in tail}) == 1, "all shards count the same number of reductions"
    assert rmse(got, single) <= 2e-6, rmse(got, single)
    if variant:
        ref, _, _ = emu(d, w, h, SHORT_L1, variant=0, bounds=bounds, ctas=ctas, tmp=str(tmp_path))
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("ctas", [1, 2])
def test_emulated_sharded_solve_gives_up_when_a_peer_is_missing(emu, tmp_path, ctas):
    """Rank 1 never starts.  Rank 0 must not hang: CTA 0 gives up after the time limit (3 s in the emulation), relays what the
    mailbox holds, the other CTAs of its grid keep in step with it (with two CTAs per GPU a divergence deadlocks the grid
    barrier: the failure seen once on the GPU box), and the solve ends with the missing rank in its status word."""
    w, h = 64, 64
    d = synth.solver_inputs(w, h, seed=2)
    got, tail, seconds = emu(d, w, h, SHORT_L1, bounds=[0, 32, 64], ctas=ctas, skip_rank=1, tmp=str(tmp_path), timeout=180)
    assert tail[0][2] == 1 + 1, tail            # status = 1 + the rank whose message never came
    assert 2.5 < seconds < 120, seconds


def test_emulated_early_out_is_taken_by_every_shard_alike(emu, tmp_path):
    """cgTolerance early-outs (the L1L preset's mechanism): the decision is taken from the all-GPU sums, which every GPU adds in
    rank order, so all shards stop at the same CG iteration -- and at the same one as the one-GPU solve here."""
    w, h = 64, 48
    d = synth.solver_inputs(w, h, seed=21)
    cfg = (3, 0.05, 0.5, 12, 2, 30.0)                # check every 2nd iteration against a tolerance that is reached on the way
    single, tail1, _ = emu(d, w, h, cfg, tmp=str(tmp_path))
    got, tail, _ = emu(d, w, h, cfg, bounds=[0, 16, 48], ctas=1, variant=1, tmp=str(tmp_path))
    assert 0 < tail1[0][1] < 36, tail1               # some IRLS iteration did stop early
    assert all(tuple(t[:2]) == tuple(tail1[0][:2]) and t[2] == 0 for t in tail), (tail, tail1)
    assert rmse(got, single) <= 2e-6


def test_emulated_solve_without_direct(emu, oracle, tmp_path):
    w, h = 36, 20
    d = synth.solver_inputs(w, h, seed=8, last_col_nonzero=True)
    got, _, _ = emu(d, w, h, L2D, bounds=[0, 16, 20], tmp=str(tmp_path), direct=False)
    want = oracle.poisson(d["dx"], d["dy"], d["throughput"], np.zeros_like(d["direct"]), alpha=0.2, preset="L2D")
    assert rmse(got, want) <= 1e-6
