"""Wall-clock phases of one end-to-end render through the public API (what bench.py's e2e leg times)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gdb200
from gdb200 import scenes

W = H = 1024
spp, streams = int(os.environ.get("SPP", "256")), 8
desc = scenes.cbox_glossy(W, H)
integ = gdb200.GPTIntegrator(reconstructL1=True, reconstructL2=False)
plan = gdb200.PoissonPlan(W, H)
host = {n: gdb200.pinned_empty((H, W, 3), "float64") for n in gdb200.BUFFER_NAMES}
for it in range(3):
    torch.cuda.synchronize(); t = [time.perf_counter()]
    sc = gdb200.Scene(desc); torch.cuda.synchronize(); t.append(time.perf_counter())
    out = integ.trace(sc, spp=spp, seed=0, preview=False, streams=streams, out=host); torch.cuda.synchronize(); t.append(time.perf_counter())
    fin = integ.reconstruct(sc, plan); torch.cuda.synchronize(); t.append(time.perf_counter())
    host["-final"][...] = fin; t.append(time.perf_counter())
    sc.close(); torch.cuda.synchronize(); t.append(time.perf_counter())
    names = ["scene_create", "trace+develop+D2H", "reconstruct+D2H", "copy_final", "scene_close"]
    print(it, {n: round(1e3 * (b - a), 1) for n, a, b in zip(names, t, t[1:])}, "device_ms", round(integ.stats.device_ms, 1), "total", round(1e3 * (t[-1] - t[0]), 1), flush=True)
