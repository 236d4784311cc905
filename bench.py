#!/usr/bin/env python
"""bench.py — G-PT Msamples/s (+ Poisson-solve ms) of the gdb200 hot path on N B200s.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line.
A "step" is one pass of the hot path over the BASELINE workload: trace the five G-PT buffers of the
configured scene at its full sample count, merge image strips across ranks (N>1), develop the film
and run the screened-Poisson reconstruction — everything GradientPathIntegrator::render does after
scene loading (reference gpt.cpp:1358-1480).

Workload at every N: BASELINE.json configs[1] — "Cornell box + glossy sphere, 1024x1024, 256 spp,
G-PT L1" (synthetic scene gdb200.scenes.cbox_glossy; the reference ships no scenes).  N>1 shards the
image into row strips (strong scaling: the image is fixed).

value : whole-job Msamples/s, scene + accumulators resident in HBM, no host copies in the timed region.
e2e   : the same through the public API (gdb200.Scene + GPTIntegrator.render) with host buffers:
        scene upload, trace, develop, 5 fp64 buffers + the reconstructed image copied back.
--impl reference: the reference's own tracer (gpt.cpp + the Mitsuba sources it runs on, compiled into oracle/_ref; the CPU
        restatement under oracle/ if that build is absent) on all host cores on a bounded sample of the same workload, plus
        the reference's own solver (oracle/_ref) for the solve time.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene builder, width, height, spp, reconstruction)
    "gpt-c2": ("cbox_glossy", 1024, 1024, 256, "L1"),
    "gpt-c1": ("cbox_diffuse", 512, 512, 64, "L2"),
    "gpt-c3": ("atrium_c3", 1920, 1080, 512, "L1"),     # BASELINE configs[2] stand-in: 258 k triangles (BVH path), sky-map lit
}
# algorithmic bytes of one reconstruction per pixel
SOLVER_BYTES = {"L1": 138684, "L2": 6900}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.rows, self.proc = device, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def load_oracle():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libgdb200_oracle.so"))
    ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_poisson_mt.so")
    ref = ctypes.CDLL(ref_path) if os.path.exists(ref_path) else None
    return lib, ref


# The reference's own tracer compiled from the reference tree (oracle/Makefile): the -O3 -march=x86-64-v3 build is the CPU
# baseline (what a user's Mitsuba would be built like); the -O2 -ffp-contract=off build exists for parity
# and is timed only if the fast one is absent.
REF_MITSUBA = [(os.path.join(ROOT, "oracle", "_ref", "libref_mitsuba_fast.so"), "-O3 -march=x86-64-v3"),
               (os.path.join(ROOT, "oracle", "_ref", "libref_mitsuba.so"), "-O2 -ffp-contract=off (parity build)")]
CPU_KIND_NOTE = {"reference": "tracer = the reference's gpt.cpp and the Mitsuba sources it runs on, compiled from the reference tree "
                              "(oracle/_ref/{lib}, {flags}), 32x32 blocks dealt to one std::thread per core, timed over the block loop "
                              "(Mitsuba's 'Render time'; scene / kd-tree construction and film development excluded)",
                 "port": "tracer = CPU restatement oracle/gpt_oracle.cpp with OpenMP over row bands (oracle/_ref/libref_mitsuba*.so not built)"}


def cpu_tracer_rate(desc, params_fn, spp, threads):
    """(Msamples/s, seconds, kind, note) of the reference tracer on the host cores, on desc at `spp`.  kind "reference": the
    reference's own gpt.cpp + the Mitsuba sources it runs on, compiled from the reference tree into oracle/_ref
    (one sample stream per pixel -- the reference has no other mode); kind "port": the CPU restatement
    oracle/gpt_oracle.cpp (OpenMP over row bands), only when no such library was built.  A library that is there but fails
    is an error, not a reason to time something else."""
    from gdb200 import scenes
    for path, flags in REF_MITSUBA:
        if not os.path.exists(path):
            continue
        import numpy as np
        ref = ctypes.CDLL(path)
        ref.gdbref_gpt_last_error.restype = ctypes.c_char_p
        prm = params_fn(spp)
        prm.streams_per_pixel = 1
        fov, rfilter = scenes.mitsuba_sensor_args(desc)
        out = np.zeros((5, desc.camera.height, desc.camera.width, 3))
        sec = ctypes.c_double()
        rc = ref.gdbref_gpt_render_timed(ctypes.byref(desc), ctypes.byref(prm), ctypes.c_double(fov), rfilter.encode(), int(threads),
                                         out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(sec))
        if rc != 0:
            raise RuntimeError(f"{path}: {ref.gdbref_gpt_last_error().decode()}")
        dt = sec.value
        return (desc.camera.width * desc.camera.height * spp / dt / 1e6, dt, "reference",
                CPU_KIND_NOTE["reference"].format(lib=os.path.basename(path), flags=flags))
    lib, _ = load_oracle()
    prm = params_fn(spp)
    B = scenes.Buffers()
    cnt = (ctypes.c_double * 3)()
    t0 = time.perf_counter()
    rc = lib.gdb200_oracle_gpt_render(ctypes.byref(desc), ctypes.byref(prm), ctypes.byref(B), None, cnt, threads)
    dt = time.perf_counter() - t0
    assert rc == 0
    return cnt[0] / dt / 1e6, dt, "port", CPU_KIND_NOTE["port"]


def cpu_solver_seconds(w, h, preset):
    """The reference's own solver sources (oracle/_ref, all cores via OMP_NUM_THREADS) on synthetic buffers."""
    import numpy as np
    from gdb200 import synth
    _, ref = load_oracle()
    if ref is None:
        return None
    d = synth.solver_inputs(w, h, seed=1234)
    out = np.empty_like(d["dx"])
    sec = ctypes.c_float()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    t0 = time.perf_counter()
    ref.ref_poisson_solve(p(d["dx"]), p(d["dy"]), p(d["throughput"]), p(d["direct"]), w, h, ctypes.c_float(0.2),
                          preset.encode(), p(out), ctypes.byref(sec))
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gdb200", choices=["gdb200", "reference"])
    ap.add_argument("--workload", default="gpt-c2", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override the workload's sample count (invalidates the headline)")
    ap.add_argument("--streams", type=int, default=16, help="sample streams per pixel (gdb200_gpt_params.streams_per_pixel); "
                    "fixed for every N so the film does not depend on the GPU count")
    ap.add_argument("--cpu-spp", type=int, default=16, help="samples/pixel of the bounded CPU-baseline sample (cpu_baseline leg and "
                    "every step of --impl reference)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    scene_name, W, H, spp, recon = WORKLOADS[args.workload]
    if args.spp:
        spp = args.spp
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))

    import gdb200
    from gdb200 import scenes, tiles
    desc = getattr(scenes, scene_name)(W, H)
    integ = gdb200.GPTIntegrator(reconstructL1=(recon == "L1"), reconstructL2=(recon == "L2"), reconstructAlpha=0.2)
    config = {"workload": f"{scene_name} {W}x{H} @ {spp} spp, G-PT {recon} reconstruction (BASELINE configs[1])"
                          if args.workload == "gpt-c2" else f"{scene_name} {W}x{H} @ {spp} spp, G-PT {recon}",
              "scene": "synthetic Cornell box + GGX spheres (gdb200.scenes)", "sampler": f"gdb200_counter seed 0, {args.streams} sample streams per pixel",
              "maxDepth": -1, "rrDepth": 5, "shiftThreshold": 0.001, "alpha": 0.2,
              "parallelism": f"{world} cost-balanced row strips, one NCCL all-reduce of the strip-boundary rows + strips sent to rank 0 (develop + solve)" if world > 1 else "1 GPU",
              "l2_flush": "per-step working set (2 KB of wavefront state per resident path slot, 8 M slots = 16 GB, + ray queues + 168 MB film) exceeds the 126 MB L2"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        cpu_spp = max(1, args.cpu_spp)
        rates = []
        for i in range(args.warmup + args.steps):
            r, dt, kind, note = cpu_tracer_rate(desc, lambda s: integ.params(s, 0), cpu_spp, cores)
            if i >= args.warmup:
                rates.append((r, dt))
        val = W * H * cpu_spp * len(rates) / sum(dt for _, dt in rates) / 1e6          # samples / time over the timed steps
        solve_s = cpu_solver_seconds(W, H, recon + "D")
        line = {"impl": "reference", "metric": "gpt_msamples_per_s", "value": round(val, 4), "unit": "Msamples/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(1e3 * sum(dt for _, dt in rates) / len(rates), 2), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "poisson_solve_ms": round(solve_s * 1e3, 1) if solve_s else None,
                "cpu_baseline": {"value": round(val, 4), "unit": "Msamples/s", "cores": cores, "kind": kind,
                                 "spp": cpu_spp,
                                 "sample": f"{scene_name} {W}x{H} @ {cpu_spp} spp per step (of {spp}); " + note + ", solver = reference sources"},
                "e2e": {"value": round(val, 4), "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ gdb200 arm (GPU)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    gdb200.lib().gdb200_set_device(local_rank)
    scene = gdb200.Scene(desc)
    plan = gdb200.PoissonPlan(W, H) if rank == 0 else None
    # N > 1: contiguous row strips.  A sample reaches at most halo rows beyond its strip, so ONE all-reduce of
    # the packed rows around the strip boundaries completes the film; rank 0 then receives the strip interiors, develops and
    # solves.  The strips are cost-balanced during the warm-up steps (tiles.rebalance): rows differ in cost, and rank 0 traces
    # less because it alone develops and solves afterwards.
    bounds = tiles.even_bounds(H, world)
    halo = tiles.halo_rows(desc.rfilter_radius)
    acc = scene.accumulators() if world > 1 else None
    phase = {"exchange_ms": 0.0, "boundary_allreduce_ms": 0.0, "gather_ms": 0.0, "develop_ms": 0.0, "solve_wall_ms": 0.0}
    agg = {"bounce_ms": 0.0, "generate_ms": 0.0, "compact_ms": 0.0, "state_bytes": 0.0, "bounce_launches": 0, "trace_ms": 0.0, "path_bounces": 0.0,
           "cast_ms": 0.0, "prepare_ms": 0.0, "resolve_ms": 0.0, "primary_ms": 0.0,
           "solve_ms": 0.0, "trace_wall_ms": 0.0, "launches": 0, "samples": 0.0, "rays": 0.0, "exchange_bytes": 0}

    row_cost = [1.0] * H                 # per-row tracing cost learnt over the warm-up steps (tiles.rebalance)

    def step(timed, balance=False):
        nonlocal bounds
        rows = (bounds[rank], bounds[rank + 1]) if world > 1 else None
        wall0 = time.perf_counter()
        integ.trace(scene, spp=spp, seed=0, rows=rows, download=False, preview=False, streams=args.streams)   # "-final" comes from the reconstruction
        wall1 = time.perf_counter()          # the call returns when the strip is traced (host side of the call included)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        nb = 0
        evb = torch.cuda.Event(enable_timing=True)
        if world > 1:
            nb = tiles.exchange_boundaries(acc, world, halo=halo, bounds=bounds)
            evb.record()
            nb += tiles.gather_strips(acc, rank, world, bounds=bounds, first_buffer=1)
        else:
            evb.record()
        ev[1].record()
        if balance and world > 1:
            torch.cuda.synchronize()
        wall2 = time.perf_counter()
        if world > 1 and rank == 0:
            scene.develop(download=False)
        ev[2].record()
        if rank == 0:
            integ.reconstruct(scene, plan, download=False)
        ev[3].record()
        if balance and world > 1:          # warm-up only: every rank learns every strip's tracing time and every rank's tail
            torch.cuda.synchronize()           # (wall clock of the calls, not device time: what decides when a rank reaches the exchange)
            mine = torch.tensor([(wall1 - wall0) * 1e3, (time.perf_counter() - wall2) * 1e3], device="cuda", dtype=torch.float64)
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            bounds = tiles.rebalance(bounds, [float(t[0]) for t in every], [float(t[1]) for t in every], row_cost=row_cost)
        if timed:
            torch.cuda.synchronize()
            phase["exchange_ms"] += ev[0].elapsed_time(ev[1]); phase["develop_ms"] += ev[1].elapsed_time(ev[2])
            phase["boundary_allreduce_ms"] += ev[0].elapsed_time(evb); phase["gather_ms"] += evb.elapsed_time(ev[1])
            phase["solve_wall_ms"] += ev[2].elapsed_time(ev[3])
            st = integ.stats
            for k in ("bounce_ms", "generate_ms", "compact_ms", "state_bytes", "bounce_launches", "samples", "rays", "path_bounces",
                      "cast_ms", "prepare_ms", "resolve_ms", "primary_ms"):
                agg[k] += getattr(st, k)
            agg["trace_ms"] += st.device_ms
            agg["trace_wall_ms"] += (wall1 - wall0) * 1e3
            agg["launches"] += st.launches + 1 + (1 if rank == 0 else 0)
            if rank == 0:
                agg["solve_ms"] += integ.solver_stats.device_ms
            if world > 1:
                agg["exchange_bytes"] += nb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False, balance=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(True)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    samples = torch.tensor([agg["samples"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(samples, op=dist.ReduceOp.SUM)
    clocks = sampler.stop() if rank == 0 else None
    total_ms, total_samples = float(ms.item()), float(samples.item())
    rank_trace_ms = [agg["trace_ms"] / args.steps]
    rank_trace_wall_ms = [agg["trace_wall_ms"] / args.steps]
    if world > 1:                                   # every rank's mean tracing time per step: names the limiting rank / phase
        mine = torch.tensor([agg["trace_ms"] / args.steps, agg["trace_wall_ms"] / args.steps], device="cuda", dtype=torch.float64)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        rank_trace_ms = [round(float(t[0].item()), 1) for t in every]
        rank_trace_wall_ms = [round(float(t[1].item()), 1) for t in every]

    # ------------------------------------------------------------------ end to end through the public API
    h2d = ctypes.sizeof(scenes.SceneDesc) + desc.n_shapes * ctypes.sizeof(scenes.Shape) + desc.n_materials * ctypes.sizeof(scenes.Material) \
        + desc.n_emitters * ctypes.sizeof(scenes.Emitter) + desc.n_vertices * 24 + desc.n_triangles * 12
    d2h = 5 * W * H * 3 * 8 + W * H * 3 * 4
    e2e_steps = max(1, min(2, args.steps))
    t0 = 0.0
    # host side of the end-to-end call: page-locked result buffers and the solver workspace are allocated once, like a
    # renderer that keeps its film between frames; scene upload, tracing, develop, solve and every D2H copy are timed
    host_out = {n: gdb200.pinned_empty((H, W, 3), "float64") for n in gdb200.BUFFER_NAMES} if rank == 0 else None
    # the end-to-end leg starts after ~15 s of sustained load: sample the clocks again so that a power/thermal drop between
    # the two legs is visible next to the number it affects ("clocks_e2e")
    sampler_e2e = ClockSampler(local_rank)
    if rank == 0:
        sampler_e2e.start()
    for it in range(e2e_steps + 1):               # iteration 0 is the end-to-end warm-up (first-use allocations), not timed
        if it == 1:
            barrier()
            t0 = time.perf_counter()
        if world == 1:
            sc = gdb200.Scene(desc)                      # scene upload (H2D) is part of the user-visible call
            out = integ.render(sc, spp=spp, seed=0, streams=args.streams, out=host_out, plan=plan)   # trace + develop + D2H of 5 buffers + solve + D2H of final
            sc.close()
        else:
            integ.trace(scene, spp=spp, seed=0, rows=(bounds[rank], bounds[rank + 1]), download=False, preview=False, streams=args.streams)
            tiles.exchange_boundaries(acc, world, halo=halo, bounds=bounds)
            tiles.gather_strips(acc, rank, world, bounds=bounds, first_buffer=1)
            if rank == 0:
                out = scene.develop(download=True, out=host_out)
                out["-final"][...] = integ.reconstruct(scene, plan, download=True)
        barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = W * H * spp * e2e_steps / float(e2e_s.item()) / 1e6
    clocks_e2e = sampler_e2e.stop() if rank == 0 else None

    if rank == 0:
        peaks, peak_src = measured_peaks()
        # Dominant kernel: the shade stage of the staged wavefront = gpt_stage_kernel<SK_SHADE0|1|2>, three specialisations of one
        # stage launched back to back every tick (csrc/gpt_stages.cuh).  Algorithmic bytes = SURVEY §8d record sizes per
        # path-bounce (read + write), counted on the device by that stage; time = the GPU's nanosecond timer, stamped by the kernels that open and close
        # the stage's launches (gdb200_stats.bounce_ms), summed over the timed region.
        shade_launches = 3 * max(1, agg["bounce_launches"])
        bounce_avg_ms = agg["bounce_ms"] / shade_launches
        achieved = agg["state_bytes"] / max(agg["bounce_ms"], 1e-9) / 1e6          # GB/s
        family_ms = agg["bounce_ms"] + agg["prepare_ms"] + agg["resolve_ms"] + agg["cast_ms"]
        traffic = None          # DRAM bytes per launch: ncu's per-path-bounce figure for this stage (profiles/) x this run's bounces per launch
        try:
            per_bounce = json.load(open(os.path.join(ROOT, "profiles", "r02_gpt_shade_summary.json")))["dram_bytes_per_path_bounce"]
            traffic = round(per_bounce * agg["path_bounces"] / shade_launches)
        except Exception:
            pass
        solve_ms = agg["solve_ms"] / args.steps
        line = {"metric": "gpt_msamples_per_s", "value": round(total_samples / total_ms / 1e3, 3), "unit": "Msamples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 2),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "tracer_msamples_per_s": round(agg["samples"] * world / max(agg["trace_ms"], 1e-9) / 1e3, 3) if world == 1
                else round(total_samples / max(agg["trace_ms"], 1e-9) / 1e3, 3),
                "poisson_solve_ms": round(solve_ms, 3), "poisson_preset": recon + "D",
                "poisson_roofline": {"bound": "hbm", "achieved": round(SOLVER_BYTES[recon] * W * H / max(solve_ms, 1e-9) / 1e6, 1),
                                     "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                     "frac": round(SOLVER_BYTES[recon] * W * H / max(solve_ms, 1e-9) / 1e6 / peaks["hbm_gbs"], 3),
                                     "kernel_variant": plan.variant,
                                     "note": "working set fits the 126 MB L2 at this size and x / Ap stay in shared memory (kernel variant 1): grade the solver roofline on tools/solver_sweep.py 4K/8K"},
                "rays_per_sample": round(agg["rays"] / max(agg["samples"], 1), 2),
                "e2e": {"value": round(e2e_val, 3), "unit": "Msamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": int(agg["launches"]),
                "clocks": clocks, "clocks_e2e": clocks_e2e,
                "roofline": {"bound": "hbm", "kernel": "gpt_stage_kernel<shade0|shade1|shade2>", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": traffic,
                             "peak_source": peak_src, "avg_launch_ms": round(bounce_avg_ms, 4),
                             "share_of_step": round(agg["bounce_ms"] / max(agg["trace_ms"], 1e-9), 3),
                             "bounce_family": {"kernels": "prepare + shade + resolve stages + both cast kernels",
                                               "achieved": round(agg["state_bytes"] / max(family_ms, 1e-9) / 1e6, 1),
                                               "frac": round(agg["state_bytes"] / max(family_ms, 1e-9) / 1e6 / peaks["hbm_gbs"], 4),
                                               "share_of_step": round(family_ms / max(agg["trace_ms"], 1e-9), 3)},
                             "note": "algorithmic bytes = SURVEY §8d wavefront record sizes per path-bounce, counted on device; the stages "
                                     "move 32-byte records of scattered slots, which HBM3e serves at ~2.5-3 TB/s (profiles/r02_stage_kernels_ncu.txt)"},
                "tracer_ms": {k: round(agg[k + "_ms"] / args.steps, 1) for k in ("bounce", "cast", "prepare", "resolve", "primary", "generate", "compact")}}
        if world > 1:
            line["exchange_bytes_per_step"] = int(agg["exchange_bytes"] / args.steps)
            line["strip_bounds"] = bounds
            line["rank_trace_ms"] = rank_trace_ms
            line["rank_trace_call_wall_ms"] = rank_trace_wall_ms
        line["rank0_phase_ms"] = {k: round(v / args.steps, 2) for k, v in phase.items()}
        if world == 1:
            try:                                  # a reported baseline: it must never cost the measured line
                rate, dt, kind, note = cpu_tracer_rate(desc, lambda s: integ.params(s, 0), args.cpu_spp, cores)
                line["cpu_baseline"] = {"value": round(rate, 4), "unit": "Msamples/s", "cores": cores, "kind": kind, "spp": args.cpu_spp,
                                        "sample": f"{scene_name} {W}x{H} @ {args.cpu_spp} spp (of {spp}), {dt:.1f} s, " + note}
            except Exception as e:                # noqa: BLE001
                line["cpu_baseline"] = {"value": None, "unit": "Msamples/s", "cores": cores, "kind": "unavailable", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
