#!/usr/bin/env python
"""Offline screened-Poisson reconstruction of a written G-PT render (README.txt:68-74 of the reference: "the gradient
and color buffers are written to disk, so ... reconstruction parameters [can be changed] at will").

    python tools/reconstruct.py <dest> [--preset L1D|L1Q|L1L|L2D|L2Q] [--alpha 0.2] [--out <file.pfm>]

reads <dest>-dx, <dest>-dy, <dest>-throughput and (if present) <dest>-direct as ".pfm" or ".exr" (a MultiFilm writes
OpenEXR by default -- files rendered by the reference load as they are), solves on the GPU through gdb200_poisson_solve
(no CPU path) and writes <dest>-final in the inputs' format (or to --out, by its extension).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import gdb200  # noqa: E402
from gdb200 import pfm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dest")
    ap.add_argument("--preset", default="L1D", choices=["L1D", "L1Q", "L1L", "L2D", "L2Q"])
    ap.add_argument("--alpha", type=float, default=0.2)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    bufs = pfm.load_multifilm(a.dest)
    for need in ("-dx", "-dy", "-throughput"):
        if need not in bufs:
            sys.exit(f"{a.dest}{need}.pfm / .exr not found")
    h, w, _ = bufs["-dx"].shape
    c = lambda x: np.ascontiguousarray(x, dtype=np.float32)  # noqa: E731
    final = gdb200.poisson_solve(c(bufs["-dx"]), c(bufs["-dy"]), c(bufs["-throughput"]),
                                 c(bufs["-direct"]) if "-direct" in bufs else None, w, h, a.alpha, a.preset)
    as_exr = not os.path.exists(a.dest + "-dx.pfm")
    out = a.out or (a.dest + ("-final.exr" if as_exr else "-final.pfm"))
    if out.lower().endswith(".exr"):
        from gdb200 import exr
        exr.write_exr(out, final, "float32" if a.out else "float16")
    else:
        pfm.write_pfm(out, final)
    print(out)


if __name__ == "__main__":
    main()
