"""Tracer throughput probe: Msamples/s of gdb200_gpt_render (device time, CUDA events inside the library)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdb200  # noqa: E402
from gdb200 import scenes  # noqa: E402

cases = [("cbox_diffuse", 512, 16), ("cbox_diffuse", 512, 64), ("cbox_glossy", 1024, 16), ("cbox_glossy", 1024, 64)]
if len(sys.argv) > 1:
    cases = [(a.split(":")[0], a.split(":")[1], int(a.split(":")[2])) for a in sys.argv[1:] if ":" in a]
streams = int(os.environ.get("GDB200_SWEEP_STREAMS", "1"))
for name, n, spp in cases:
    w, h = (int(x) for x in str(n).split("x")) if "x" in str(n) else (int(n), int(n))
    desc = getattr(scenes, name)(w, h)
    scene = gdb200.Scene(desc)
    integ = gdb200.GPTIntegrator(reconstructL1=False, reconstructL2=False)
    integ.fusedBounce = os.environ.get("GDB200_SWEEP_FUSED") == "1"          # round-1 single-kernel bounce
    integ.maxSlots = int(os.environ.get("GDB200_SWEEP_SLOTS", "0"))
    for rep in range(2):
        integ.trace(scene, spp=spp, seed=0, download=False, streams=streams)
    st = integ.stats
    print(json.dumps({"scene": name, "size": n, "spp": spp, "streams": streams, "triangles": desc.n_triangles, "ms": round(st.device_ms, 2), "launches": st.launches,
                      "Msamples_s": round(st.samples / st.device_ms / 1e3, 2), "rays_per_sample": round(st.rays / st.samples, 2),
                      "avg_depth": round(st.path_vertices / st.samples, 3), "Grays_s": round(st.rays / st.device_ms / 1e6, 2), "steps": st.bounce_launches,
                      "gen_ms": round(st.generate_ms, 1), "compact_ms": round(st.compact_ms, 1), "bounce_ms": round(st.bounce_ms, 1),
                      "fused": integ.fusedBounce, "slots": integ.maxSlots, "cast_ms": round(st.cast_ms, 1), "prepare_ms": round(st.prepare_ms, 1),
                      "resolve_ms": round(st.resolve_ms, 1), "primary_ms": round(st.primary_ms, 1)}), flush=True)
    scene.close()
