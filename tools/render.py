#!/usr/bin/env python
"""Stand-alone driver: renders a Mitsuba 0.5 scene file that selects the `gpt` integrator on the GPU and writes the five
MultiFilm buffers, like `mitsuba [-D key=value] [-o dest] scene.xml` does for the reference (src/mitsuba/mitsuba.cpp:60-84,
multifilm.cpp:423-516).

    python tools/render.py scene.xml -o out/render -D spp=64

writes out/render-final, -throughput, -dx, -dy, -direct as .exr (the film's default fileFormat "openexr", float16) or .pfm
(fileFormat "pfm").  Scene subset: gdb200.xmlscene.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdb200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("-o", "--output", default=None, help="destination file name (extension is replaced; default: the scene's name)")
    ap.add_argument("-D", action="append", default=[], metavar="key=value", help="define a $parameter of the scene file")
    ap.add_argument("--streams", type=int, default=None, help="sample streams per pixel (overrides the scene's streamsPerPixel)")
    a = ap.parse_args()
    defines = dict(kv.split("=", 1) for kv in a.D)
    parsed = gdb200.load_scene(a.scene, defines)
    integ = parsed.integrator()
    scene = gdb200.Scene(parsed.desc)
    t0 = time.perf_counter()
    out = integ.render(scene, spp=parsed.spp, seed=parsed.seed, streams=a.streams or parsed.streams)
    dt = time.perf_counter() - t0
    dest = a.output or os.path.splitext(a.scene)[0]
    paths = integ.save(dest, out, parsed.file_format, parsed.component_format)
    st = integ.stats
    print(f"Render time: {dt:.2f} s  ({st.samples / max(st.device_ms, 1e-9) / 1e3:.1f} Msamples/s traced, "
          f"reconstruction {integ.solver_stats.device_ms:.1f} ms)")
    for p in paths:
        print("Writing image to \"%s\" .." % p)


if __name__ == "__main__":
    main()
