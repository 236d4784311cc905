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


class PoissonPlan:
    """Device workspace for one image size (gdb200_poisson_plan)."""

    def __init__(self, w, h):
        self.w, self.h = int(w), int(h)
        self._h = ctypes.c_void_p()
        check(lib().gdb200_poisson_plan_create(self.w, self.h, ctypes.byref(self._h)))

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
