"""A/B sweep over builds of libgdb200 (csrc/Makefile: OUT=/B=/GPT_DEFS=) and runtime knobs, one subprocess each:
device-time Msamples/s of gdb200_gpt_render on the bench scenes plus a small parity check against the oracle.

  python tools/variant_sweep.py base fma pf o3 ... [--slots 1048576,2097152] [--streams 1,8] [--cases cbox_glossy:1024:32]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gradientdomain-mitsuba_b200")

CHILD = r"""
import json, os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
import gdb200
from gdb200 import scenes
from conftest import Oracle
cases, streams = %(cases)r, %(streams)r
res = {"variant": %(name)r, "slots": os.environ.get("GDB200_MAX_SLOTS"), "streams": streams}
desc = scenes.cbox_glossy(64, 64)
integ = gdb200.GPTIntegrator(reconstructL1=False, reconstructL2=False)
got = integ.trace(gdb200.Scene(desc), spp=8, seed=3, streams=streams)
ref, _, _ = Oracle().gpt(desc, integ.params(8, 3, streams=streams))
worst, flips = 0.0, 0.0
for k in ("-throughput", "-dx", "-dy", "-direct"):
    scale = max(float(np.abs(ref[k]).mean()), 1e-12)
    diff = np.abs(got[k] - ref[k]).max(axis=2)
    bad = diff > 1e-5 * scale
    flips = max(flips, float(bad.mean()))
    ok = ~bad
    worst = max(worst, float(np.sqrt(np.mean((got[k] - ref[k])[ok] ** 2))) / scale)
res["parity_rel_rmse"], res["flipped_frac"] = worst, flips
for name, n, spp in cases:
    d = getattr(scenes, name)(n, n)
    sc = gdb200.Scene(d)
    best = None
    for rep in range(3):
        integ.trace(sc, spp=spp, seed=0, download=False, preview=False, streams=streams)
        st = integ.stats
        r = st.samples / st.device_ms / 1e3
        if best is None or r > best[0]:
            best = (r, st.device_ms, st.bounce_ms, st.generate_ms, st.bounce_launches, st.launches, st.cast_ms, st.prepare_ms, st.resolve_ms,
                    st.primary_ms, st.compact_ms)
    res["%%s:%%d:%%d" %% (name, n, spp)] = {"Msamples_s": round(best[0], 2), "ms": round(best[1], 1), "shade_ms": round(best[2], 1),
                                        "gen_ms": round(best[3], 1), "steps": best[4], "launches": best[5], "cast_ms": round(best[6], 1),
                                        "prepare_ms": round(best[7], 1), "resolve_ms": round(best[8], 1), "primary_ms": round(best[9], 1),
                                        "compact_ms": round(best[10], 1)}
    sc.close()
print(json.dumps(res), flush=True)
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="+", help="'base' or the suffix x of gradientdomain-mitsuba_b200/libgdb200_x.so")
    ap.add_argument("--slots", default="", help="comma list of GDB200_MAX_SLOTS values (default: library default)")
    ap.add_argument("--streams", default="1", help="comma list of streams_per_pixel values")
    ap.add_argument("--cases", default="cbox_glossy:1024:32,cbox_diffuse:512:64")
    args = ap.parse_args()
    cases = [(c.split(":")[0], int(c.split(":")[1]), int(c.split(":")[2])) for c in args.cases.split(",")]
    for v in args.variants:
        lib = os.path.join(PKG, "libgdb200.so" if v == "base" else f"libgdb200_{v}.so")
        for slots in (args.slots.split(",") if args.slots else [""]):
            for streams in [int(x) for x in args.streams.split(",")]:
                env = dict(os.environ, GDB200_LIBRARY=lib)
                if slots:
                    env["GDB200_MAX_SLOTS"] = slots
                code = CHILD % {"root": ROOT, "cases": cases, "streams": streams, "name": v}
                r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
                out = r.stdout.strip().splitlines()
                print(out[-1] if out and r.returncode == 0 else json.dumps({"variant": v, "error": r.stderr[-600:]}), flush=True)


if __name__ == "__main__":
    main()
