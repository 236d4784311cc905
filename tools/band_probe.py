"""Per-GPU tracer throughput when one rank of an N-GPU run owns 1/N of the image (interleaved 16-row bands),
measured on ONE GPU: predicts the strong-scaling tracer rate without an N-GPU box.
  python tools/band_probe.py [scene:size:spp] [--world 8] [--streams 1,8]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdb200  # noqa: E402
from gdb200 import scenes  # noqa: E402

case = next((a for a in sys.argv[1:] if ":" in a), "cbox_glossy:1024:64")
world = int(sys.argv[sys.argv.index("--world") + 1]) if "--world" in sys.argv else 8
streams = [int(x) for x in sys.argv[sys.argv.index("--streams") + 1].split(",")] if "--streams" in sys.argv else [1, 8]
name, n, spp = case.split(":")[0], int(case.split(":")[1]), int(case.split(":")[2])
scene = gdb200.Scene(getattr(scenes, name)(n, n))
integ = gdb200.GPTIntegrator(reconstructL1=False, reconstructL2=False)
for s in streams:
    for bands in (None, (16, world, 0), (16, world, world - 1)):
        for rep in range(2):
            integ.trace(scene, spp=spp, seed=0, download=False, preview=False, bands=bands, streams=s)
        st = integ.stats
        print(json.dumps({"case": case, "streams": s, "bands": bands, "ms": round(st.device_ms, 1),
                          "Msamples_s_this_rank": round(st.samples / st.device_ms / 1e3, 2),
                          "x_world": round((world if bands else 1) * st.samples / st.device_ms / 1e3, 1), "steps": st.bounce_launches}), flush=True)
