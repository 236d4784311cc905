"""Generates tests/golden/poisson_golden.json from the REFERENCE's own solver
(oracle/_ref/libref_poisson.so = src/integrators/poisson_solver/*.cpp compiled
unmodified).  Run in the build container (needs /root/reference):
    python tests/golden/make_poisson_golden.py
The reference ships no golden vectors for this path, so these
are outputs of the reference itself on the seeded inputs of gdb200.synth."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gdb200  # noqa: E402,F401
from gdb200 import synth  # noqa: E402
from conftest import Oracle  # noqa: E402

CASES = [(64, 48, 1234, False), (33, 17, 5, True), (128, 128, 99, False), (1, 1, 3, True), (7, 1, 3, True),
         (1, 9, 3, True), (130, 70, 11, True)]


def main():
    orc = Oracle()
    assert orc.ref is not None, "reference solver not built"
    out = []
    for (w, h, seed, last) in CASES:
        d = synth.solver_inputs(w, h, seed=seed, last_col_nonzero=last)
        for preset in ("L2D", "L1D"):
            for with_direct in (True, False):
                fin = orc.poisson_ref(d["dx"], d["dy"], d["throughput"], d["direct"] if with_direct else None,
                                      alpha=0.2, preset=preset)
                out.append({"w": w, "h": h, "seed": seed, "last_col_nonzero": last, "preset": preset,
                            "direct": with_direct, "alpha": 0.2,
                            "input_sha256": hashlib.sha256(b"".join(d[k].tobytes() for k in
                                                                    ("throughput", "dx", "dy", "direct"))).hexdigest(),
                            "final_sha256": hashlib.sha256(fin.tobytes()).hexdigest(),
                            "final_mean": float(fin.astype(np.float64).mean()),
                            "final_probe": [float(v) for v in fin.reshape(-1)[:: max(1, fin.size // 8)][:8]]})
    with open(os.path.join(ROOT, "tests", "golden", "poisson_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
