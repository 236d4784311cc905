"""Poisson-solve-only sweep (BASELINE config 5): device-resident inputs, CUDA-event
time inside the library, algorithmic GB/s (SURVEY.md §8d figures) vs measured HBM peak."""
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

sizes = [(512, 512), (1024, 1024), (1920, 1080), (3840, 2160), (7680, 4320)]
if len(sys.argv) > 1:
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (w, h) in sizes:
    d = synth.solver_inputs(min(w, 1024), min(h, 1024), seed=1234)
    reps = (-(-h // d["dx"].shape[0]), -(-w // d["dx"].shape[1]), 1)
    t = {k: torch.from_numpy(np.tile(v, reps)[:h, :w].copy()).cuda() for k, v in d.items()}
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
        gbs = BYTES[preset] * w * h / (ms * 1e-3) / 1e9
        print(json.dumps({"size": f"{w}x{h}", "preset": preset, "ms": round(ms, 3), "all_ms": [round(x, 3) for x in times],
                          "alg_GBs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 3)}), flush=True)
    plan.close()
    del t, out
    torch.cuda.empty_cache()
