"""CPU test (world_size 2, gloo) of the host layer of the sharded Poisson solve (gdb200.poisson.ShardedPoissonSolver): band
arithmetic, the all-gather of the shard handles, connect order, the collective solve call and the gather of the bands.
The C library is replaced by a stand-in (there is no GPU here) whose "kernel" is the CPU oracle solving the whole image and
writing only this rank's band -- exactly the contract of gdb200_poisson_solve_device on a shard (include/gdb200.h).  The real
kernels are covered on the GPU by tests/test_poisson_gpu.py."""
import ctypes
import os
import struct

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gdb200
from gdb200 import poisson, synth
from conftest import Oracle

W, H = 40, 70          # 5 tile rows: bands [0, 32) and [32, 70)


class StandInLibrary:
    """The four shard entry points + solve_device, with the argument order of include/gdb200.h."""

    def __init__(self, oracle, log, real):
        self.oracle, self.log, self.shard, self.real = oracle, log, None, real

    def __getattr__(self, name):          # host-only entry points (gdb200_poisson_preset, ...) stay the real ones
        return getattr(self.real, name)

    def gdb200_poisson_shard_create(self, w, h, y0, y1, rank, n, out_plan):
        self.shard = dict(w=w, h=h, y0=y0, y1=y1, rank=rank, n=n)
        out_plan._obj.value = 0x1000 + rank          # a non-NULL plan handle (byref argument)
        return 0

    def gdb200_poisson_shard_export(self, plan, buf):
        s = self.shard
        blob = bytearray(poisson.SHARD_HANDLE_BYTES)
        struct.pack_into("<q6i", blob, 144, os.getpid(), 0, s["rank"], s["y0"], s["y1"], s["w"], s["h"])   # pid, device, rank, y0, y1, w, h
        ctypes.memmove(buf, bytes(blob), len(blob))
        return 0

    def gdb200_poisson_shard_connect(self, plan, blob, n):
        raw = ctypes.string_at(blob, n * poisson.SHARD_HANDLE_BYTES)
        peers = [struct.unpack_from("<q6i", raw, r * poisson.SHARD_HANDLE_BYTES + 144) for r in range(n)]
        self.log["peers"] = [(p[2], p[3], p[4]) for p in peers]            # (rank, y0, y1) in the order received
        return 0

    def gdb200_poisson_solve_device(self, plan, dx, dy, thr, direct, alpha, cfg, out, stream, stats):
        s = self.shard
        n = s["w"] * s["h"] * 3
        arr = lambda p: np.ctypeslib.as_array((ctypes.c_float * n).from_address(p.value)).reshape(s["h"], s["w"], 3)   # noqa: E731
        full = self.oracle.poisson(arr(dx), arr(dy), arr(thr), arr(direct), alpha=alpha.value, preset="L2D")
        arr(out)[s["y0"]:s["y1"]] = full[s["y0"]:s["y1"]]                  # a shard fills its band only
        return 0

    def gdb200_poisson_plan_destroy(self, plan):
        self.log["destroyed"] = True


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        log = {}
        fake = StandInLibrary(Oracle(), log, poisson.lib())
        poisson.lib = lambda: fake
        d = synth.solver_inputs(W, H, seed=3)
        t = {k: torch.from_numpy(v.copy()) for k, v in d.items()}
        solver = gdb200.ShardedPoissonSolver(W, H)
        assert solver.bounds == [0, 32, 70] and solver.plan.band == (solver.bounds[rank], solver.bounds[rank + 1])
        assert log["peers"] == [(0, 0, 32), (1, 32, 70)], log           # every rank connects to all handles, in rank order
        params = gdb200.SolverParams()
        params.setConfigPreset("L2D")
        out = torch.full((H, W, 3), -1.0)
        solver.solve_device(t["dx"], t["dy"], t["throughput"], t["direct"], 0.2, params.cfg, out)
        y0, y1 = solver.plan.band
        mask = torch.ones(H, dtype=torch.bool); mask[y0:y1] = False
        assert bool((out[mask] == -1.0).all()), "a shard writes its own band only"
        solver.gather(out)
        if rank == 0:
            ret["final"] = out.numpy().copy()
        else:
            assert bool((out[mask] == -1.0).all()), "only the destination receives the other bands"
        solver.close()
        assert log.get("destroyed")
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_solve_host_layer():
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, 29621, ret), nprocs=world, join=True)
        final = ret["final"]
    d = synth.solver_inputs(W, H, seed=3)
    want = Oracle().poisson(d["dx"], d["dy"], d["throughput"], d["direct"], alpha=0.2, preset="L2D")
    np.testing.assert_array_equal(final, want)
