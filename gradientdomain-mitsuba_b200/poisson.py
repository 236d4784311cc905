"""Host-side mirror of ``poisson::Solver`` (reference
src/integrators/poisson_solver/Solver.hpp:48-158) over the gdb200 C ABI.

The method names, argument meaning and call order are the reference's
(``importImagesMTS`` → ``setupBackend`` → ``solveIndirect`` → ``exportImagesMTS``,
call site gpt.cpp:1456-1462) so the parity tests read like a user of the
reference.  The arithmetic runs in the sm_100a kernel of ``csrc/poisson.cu``.
"""
import ctypes

import numpy as np

from ._ffi import lib, check, Stats, PoissonConfig, Gdb200Error


class SolverParams:
    """``poisson::Solver::Params``: alpha + solver configuration (Solver.cpp:57-164)."""

    def __init__(self):
        self.alpha = 0.2            # Solver.cpp:64
        self.verbose = False
        self.logFunc = None
        self.cfg = PoissonConfig()
        self.preset = None
        self.setConfigPreset("L1D")  # Solver.cpp:85

    def setConfigPreset(self, preset):
        """Returns False for an unknown preset, like the reference (Solver.cpp:163)."""
        cfg = PoissonConfig()
        rc = lib().gdb200_poisson_preset(preset.encode(), ctypes.byref(cfg))
        if rc != 0:
            return False
        self.cfg, self.preset = cfg, preset
        return True

    def setLogFunction(self, fn):
        self.logFunc = fn


SHARD_HANDLE_BYTES = 176          # sizeof(gdb200_shard_handle), include/gdb200.h
SHARD_ROW_ALIGN = 16              # the kernel's tile height: bands that are multiples of it leave no partly filled tile row


def shard_bounds(h, n):
    """Row bands [b[r], b[r+1]) of an h-row image for n GPUs: equal numbers of 16-row tile rows, remainder to the first ranks."""
    tiles = -(-h // SHARD_ROW_ALIGN)
    if n < 1 or n > tiles:
        raise Gdb200Error(f"cannot split {h} rows ({tiles} tile rows) over {n} GPUs")
    cuts = [min(h, ((tiles * r) // n) * SHARD_ROW_ALIGN) for r in range(n)] + [h]
    return cuts


class PoissonPlan:
    """Device workspace for one image size (gdb200_poisson_plan); with ``band`` one GPU's share of a sharded solve."""

    def __init__(self, w, h, band=None, rank=0, n_ranks=1):
        self.w, self.h = int(w), int(h)
        self.band = (0, self.h) if band is None else (int(band[0]), int(band[1]))
        self.rank, self.n_ranks = int(rank), int(n_ranks)
        self._h = ctypes.c_void_p()
        if band is None and n_ranks == 1:
            check(lib().gdb200_poisson_plan_create(self.w, self.h, ctypes.byref(self._h)))
        else:
            check(lib().gdb200_poisson_shard_create(self.w, self.h, self.band[0], self.band[1], self.rank, self.n_ranks,
                                                    ctypes.byref(self._h)))

    def export_handle(self):
        """gdb200_poisson_shard_export: the bytes the other ranks need to reach this shard's halo rows and mailbox."""
        buf = ctypes.create_string_buffer(SHARD_HANDLE_BYTES)
        check(lib().gdb200_poisson_shard_export(self._h, buf))
        return buf.raw

    def connect(self, handles):
        """gdb200_poisson_shard_connect with every rank's handle (list of bytes, rank order)."""
        blob = b"".join(handles)
        if len(blob) != SHARD_HANDLE_BYTES * self.n_ranks:
            raise Gdb200Error(f"expected {self.n_ranks} handles of {SHARD_HANDLE_BYTES} bytes")
        check(lib().gdb200_poisson_shard_connect(self._h, ctypes.c_char_p(blob), self.n_ranks))

    def solve_device(self, dx, dy, throughput, direct, alpha, cfg, out, stream=None, stats=None):
        """All image arguments are CUDA device pointers (ints) or objects with
        ``data_ptr()`` (torch tensors); throughput/direct may be None."""
        def ptr(t):
            if t is None:
                return None
            return ctypes.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else int(t))
        check(lib().gdb200_poisson_solve_device(self._h, ptr(dx), ptr(dy), ptr(throughput), ptr(direct),
                                                ctypes.c_float(alpha), ctypes.byref(cfg), ptr(out),
                                                ctypes.c_void_p(stream or 0),
                                                ctypes.byref(stats) if stats is not None else None))

    @property
    def variant(self):
        """Kernel variant the solves of this plan run: 0 streaming, 1 x and Ap resident in shared memory (small images),
        2 x resident, 3 streaming with the shared-tile exchange (include/gdb200.h). Same bits from all of them."""
        return int(lib().gdb200_poisson_plan_variant(self._h))

    @variant.setter
    def variant(self, v):
        check(lib().gdb200_poisson_plan_set_variant(self._h, int(v)))

    def metrics_device(self, err, stream=None):
        """Solver::evaluateMetricsMTS on the x the last solve left in the plan: writes the primal residual image to the device
        buffer ``err`` and returns (errL1, errL2)."""
        l1, l2 = ctypes.c_float(), ctypes.c_float()
        check(lib().gdb200_poisson_metrics_device(self._h, ctypes.c_void_p(err.data_ptr() if hasattr(err, "data_ptr") else int(err)),
                                                  ctypes.byref(l1), ctypes.byref(l2), ctypes.c_void_p(stream or 0)))
        return l1.value, l2.value

    def close(self):
        if self._h:
            lib().gdb200_poisson_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _as_f32(a, w, h, name):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.size != 3 * w * h:
        raise Gdb200Error(f"{name}: expected {3 * w * h} floats for {w}x{h} RGB, got {a.size}")
    return a


def poisson_solve(dx, dy, throughput, direct, w, h, alpha=0.2, preset="L1D", out=None, stats=None):
    """One-shot host-buffer solve (gdb200_poisson_solve). Returns (h, w, 3) float32."""
    dx, dy = _as_f32(dx, w, h, "dx"), _as_f32(dy, w, h, "dy")
    throughput, direct = _as_f32(throughput, w, h, "throughput"), _as_f32(direct, w, h, "direct")
    if out is None:
        out = np.empty((h, w, 3), dtype=np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None  # noqa: E731
    check(lib().gdb200_poisson_solve(p(dx), p(dy), p(throughput), p(direct), w, h, ctypes.c_float(alpha),
                                     preset.encode(), p(out), ctypes.byref(stats) if stats is not None else None))
    return out


class ShardedPoissonSolver:
    """All GPUs of a torch.distributed group solve ONE image together (include/gdb200.h "sharded solve"): rank r owns a band
    of rows; inside the one persistent kernel per GPU the halo rows and the CG reductions travel over NVLink peer memory.
    torch.distributed only carries the 176-byte handles once, at construction."""

    def __init__(self, w, h, group=None, bounds=None):
        import torch.distributed as dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.bounds = shard_bounds(h, self.world) if bounds is None else list(bounds)
        self.plan = PoissonPlan(w, h, band=(self.bounds[self.rank], self.bounds[self.rank + 1]), rank=self.rank, n_ranks=self.world)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.plan.export_handle(), group=group)
        self.plan.connect(handles)
        dist.barrier(group)                      # nobody starts a solve before everybody can be reached

    def solve_device(self, dx, dy, throughput, direct, alpha, cfg, out, stream=None, stats=None):
        """Collective: every rank calls it with the WHOLE images on its own GPU; fills rows bounds[rank]:bounds[rank+1] of out."""
        self.plan.solve_device(dx, dy, throughput, direct, alpha, cfg, out, stream=stream, stats=stats)

    def gather(self, out, dst=0):
        """Rank dst receives the other ranks' bands of ``out`` (h, w, 3 float32 on the GPU) so that its copy is the whole image."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return
        if self.rank == dst:
            bufs = {r: torch.empty_like(out[self.bounds[r]:self.bounds[r + 1]]) for r in range(self.world) if r != dst}
            for req in dist.batch_isend_irecv([dist.P2POp(dist.irecv, t, r, self.group) for r, t in bufs.items()]):
                req.wait()
            for r, t in bufs.items():
                out[self.bounds[r]:self.bounds[r + 1]].copy_(t)
        else:
            mine = out[self.bounds[self.rank]:self.bounds[self.rank + 1]].contiguous()
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, dst, self.group)]):
                req.wait()

    def close(self):
        self.plan.close()


class PoissonSolver:
    """``poisson::Solver`` look-alike."""

    def __init__(self, params):
        self.params = params
        self._imgs = None
        self._plan_ready = False
        self._final = None
        self.stats = Stats()

    def importImagesMTS(self, dx, dy, tp, direct, width, height):
        self._imgs = (_as_f32(dx, width, height, "dx"), _as_f32(dy, width, height, "dy"),
                      _as_f32(tp, width, height, "throughput"), _as_f32(direct, width, height, "direct"),
                      int(width), int(height))

    def setupBackend(self):
        if self._imgs is None:
            raise Gdb200Error("setupBackend() before importImagesMTS()")   # reference asserts, Solver.cpp:259-260
        self._plan_ready = True

    def solveIndirect(self):
        if not self._plan_ready:
            raise Gdb200Error("solveIndirect() before setupBackend()")     # Solver.cpp:376
        dx, dy, tp, direct, w, h = self._imgs
        self._final = poisson_solve(dx, dy, tp, direct, w, h, self.params.alpha, self.params.preset,
                                    stats=self.stats)
        if self.params.logFunc:
            self.params.logFunc("Execution time = %.2f s\n" % (self.stats.device_ms * 1e-3))  # Solver.cpp:500

    def evaluateMetricsMTS(self, err):
        """Solver::evaluateMetricsMTS (Solver.cpp:511-541): fills err (h, w, 3 float32) with the primal block of b - P*x and
        returns (errL1, errL2), the mean |e| and mean |e|^2 over the 3*w*h RGB elements of e."""
        if self._final is None:
            raise Gdb200Error("evaluateMetricsMTS() before solveIndirect()")      # reference asserts m_x, Solver.cpp:513
        L = lib()
        out = np.empty(self._final.shape, dtype=np.float32)
        l1, l2 = ctypes.c_float(), ctypes.c_float()
        check(L.gdb200_poisson_metrics(out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(l1), ctypes.byref(l2)))
        np.copyto(np.asarray(err).reshape(out.shape), out)
        return l1.value, l2.value

    def exportImagesMTS(self, rec):
        if self._final is None:
            raise Gdb200Error("exportImagesMTS() before solveIndirect()")
        np.copyto(np.asarray(rec).reshape(self._final.shape), self._final)
