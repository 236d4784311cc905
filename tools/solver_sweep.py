"""Poisson-solve-only sweep (BASELINE config 5): device-resident inputs, CUDA-event
time inside the library, algorithmic GB/s (SURVEY.md §8d figures) vs measured HBM peak.

Next to it the GPU bar BASELINE.md §3 names: the REFERENCE's own solver running on its own CUDA backend (BackendCUDA.cu compiled
unmodified for sm_100a into oracle/_ref/libref_poisson_cuda.so, see oracle/Makefile), on the same inputs, timed by the
reference's own timer ("Execution time": Solver::solveIndirect between its CUDA events, Solver.cpp:378,500)."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdb200  # noqa: E402
from gdb200 import synth  # noqa: E402

BYTES = {"L2D": 6900, "L1D": 138684}
peak = 6544.0
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CUDA = os.path.join(ROOT, "oracle", "_ref", "libref_poisson_cuda.so")
ref_cuda = ctypes.CDLL(REF_CUDA) if os.path.exists(REF_CUDA) else None


def reference_cuda_ms(host, w, h, preset):
    """(best ms of 2 runs, result) of the reference solver on its CUDA backend, or (None, None) without that build."""
    if ref_cuda is None:
        return None, None
    out = np.empty((h, w, 3), dtype=np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    best = None
    for _ in range(2):
        sec = ctypes.c_float(-1)
        rc = ref_cuda.ref_poisson_solve_backend(p(host["dx"]), p(host["dy"]), p(host["throughput"]), p(host["direct"]), w, h,
                                                ctypes.c_float(0.2), preset.encode(), b"CUDA", p(out), ctypes.byref(sec))
        if rc != 0 or sec.value < 0:
            return None, None
        best = sec.value * 1e3 if best is None else min(best, sec.value * 1e3)
    return best, out


sizes = [(512, 512), (1024, 1024), (1920, 1080), (3840, 2160), (7680, 4320)]
if len(sys.argv) > 1:
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (w, h) in sizes:
    d = synth.solver_inputs(min(w, 1024), min(h, 1024), seed=1234)
    reps = (-(-h // d["dx"].shape[0]), -(-w // d["dx"].shape[1]), 1)
    host = {k: np.ascontiguousarray(np.tile(v, reps)[:h, :w]) for k, v in d.items()}
    t = {k: torch.from_numpy(v).cuda() for k, v in host.items()}
    out = torch.empty_like(t["dx"])
    plan = gdb200.PoissonPlan(w, h)
    for preset in ("L2D", "L1D"):
        params = gdb200.SolverParams()
        params.setConfigPreset(preset)
        times = []
        for it in range(4 if preset == "L2D" else 3):
            st = gdb200.Stats()
            plan.solve_device(t["dx"], t["dy"], t["throughput"], t["direct"], 0.2, params.cfg, out, stats=st)
            times.append(st.device_ms)
        ms = min(times[1:])
        chosen = plan.variant
        others = {}
        for v in (0, 1, 2, 3):                             # A/B: the same solve by the other kernel variants that fit
            if v == chosen:
                continue
            try:
                plan.variant = v
            except gdb200.Gdb200Error:
                continue
            ts = []
            for it in range(3):
                st = gdb200.Stats()
                plan.solve_device(t["dx"], t["dy"], t["throughput"], t["direct"], 0.2, params.cfg, out, stats=st)
                ts.append(st.device_ms)
            others[str(v)] = round(min(ts[1:]), 3)
        plan.variant = chosen
        st = gdb200.Stats()
        plan.solve_device(t["dx"], t["dy"], t["throughput"], t["direct"], 0.2, params.cfg, out, stats=st)
        gbs = BYTES[preset] * w * h / (ms * 1e-3) / 1e9
        row = {"size": f"{w}x{h}", "preset": preset, "ms": round(ms, 3), "all_ms": [round(x, 3) for x in times],
               "alg_GBs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 3)}
        row["variant"] = chosen
        row["other_variants_ms"] = others
        ref_ms, ref_out = reference_cuda_ms(host, w, h, preset)
        if ref_ms is not None:
            mine = out.cpu().numpy()
            row.update({"reference_cuda_backend_ms": round(ref_ms, 3), "speedup_vs_reference_cuda": round(ref_ms / ms, 2),
                        "reference_cuda_alg_GBs": round(BYTES[preset] * w * h / (ref_ms * 1e-3) / 1e9, 1),
                        "rmse_vs_reference_cuda": float(np.sqrt(np.mean((mine.astype(np.float64) - ref_out) ** 2)))})
        print(json.dumps(row), flush=True)
    plan.close()
    del t, out
    torch.cuda.empty_cache()
