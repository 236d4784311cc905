#!/usr/bin/env python
"""Development tool: render a Mitsuba scene file (the subset of gdb200.xmlscene) with the REFERENCE's own G-PT tracer --
oracle/_ref/libref_mitsuba.so, gpt.cpp compiled from the reference tree (oracle/Makefile) -- on the host cores and, where a
GPU is present, with libgdb200 on the same scene bytes and sample streams, and report how the five buffers differ.

    python tools/compare_with_reference.py scene.xml [-D spp=16] [--write dest]

GPTIntegrator.refUninitMeasure (gdb200_gpt_params.flags) is set for the GPU render: gpt.cpp:957 reads an uninitialised value, and this reproduces what the
g++ build of the reference does there.  One sample stream per pixel (the reference has no other mode).
Test infrastructure: this is the only tool that loads anything under oracle/."""
import argparse
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import gdb200  # noqa: E402
from gdb200 import scenes  # noqa: E402

NAMES = ("-final", "-throughput", "-dx", "-dy", "-direct")


def reference_render(desc, prm, threads):
    path = os.path.join(ROOT, "oracle", "_ref", "libref_mitsuba.so")
    if not os.path.exists(path):
        sys.exit("oracle/_ref/libref_mitsuba.so is not built (make -C oracle ref, needs /root/reference)")
    lib = ctypes.CDLL(path)
    lib.gdbref_gpt_last_error.restype = ctypes.c_char_p
    fov, rfilter = scenes.mitsuba_sensor_args(desc)
    out = np.zeros((5, desc.camera.height, desc.camera.width, 3))
    t0 = time.perf_counter()
    rc = lib.gdbref_gpt_render(ctypes.byref(desc), ctypes.byref(prm), ctypes.c_double(fov), rfilter.encode(), threads,
                               out.ctypes.data_as(ctypes.c_void_p))
    if rc:
        sys.exit("reference: " + lib.gdbref_gpt_last_error().decode())
    return dict(zip(NAMES, out)), time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("-D", action="append", default=[], metavar="key=value")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--write", default=None, help="write the reference's buffers as <dest>-ref-*.pfm")
    a = ap.parse_args()
    parsed = gdb200.load_scene(a.scene, dict(kv.split("=", 1) for kv in a.D))
    integ = parsed.integrator()
    integ.reconstructL1 = integ.reconstructL2 = False
    integ.refUninitMeasure = True      # gpt.cpp:957: what the compiled reference does there (include/gdb200.h)
    prm = integ.params(parsed.spp, parsed.seed)
    n = parsed.desc.camera.width * parsed.desc.camera.height * parsed.spp
    ref, dt = reference_render(parsed.desc, prm, a.threads)
    print(f"reference: {n} samples in {dt:.2f} s on {a.threads} threads = {n / dt / 1e6:.3f} Msamples/s")
    if a.write:
        from gdb200 import pfm
        for k, v in ref.items():
            pfm.write_pfm(f"{a.write}-ref{k}.pfm", v)
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if not have_gpu:
        print("no GPU: nothing to compare with")
        return
    t0 = time.perf_counter()
    got = integ.trace(gdb200.Scene(parsed.desc), spp=parsed.spp, seed=parsed.seed)
    dt = time.perf_counter() - t0
    print(f"gdb200:    {n} samples in {dt:.2f} s = {n / dt / 1e6:.1f} Msamples/s (incl. scene upload and download)")
    for k in NAMES:
        scale = max(float(np.abs(ref[k]).mean()), 1e-12)
        diff = np.abs(got[k] - ref[k]).max(axis=2)
        flipped = diff > 1e-7 * scale
        rms = float(np.sqrt(np.mean((got[k] - ref[k])[~flipped] ** 2))) / scale if (~flipped).any() else 0.0
        print(f"  {k:12s} pixels touched by a flipped branch: {int(flipped.sum()):6d} of {flipped.size}   relative RMS of the rest: {rms:.2e}")


if __name__ == "__main__":
    main()
