"""Sharded Poisson solve across the GPUs of one node, one process per GPU (CUDA IPC peer memory):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/shard_solve_check.py [WxH ...]
Every rank holds the whole synthetic input (in the render flow: after the all-gather of the developed buffers), solves its band,
rank 0 gathers the bands, compares with its own single-GPU solve of the same input and prints one JSON line per size / preset:
device time = max over ranks of the kernel time (they run in lock step), algorithmic GB/s summed over the GPUs."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gdb200  # noqa: E402
from gdb200 import synth  # noqa: E402

BYTES = {"L2D": 6900, "L1D": 138684}


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peak = 6544.0
    try:
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        pass
    sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(1024, 1024), (3840, 2160), (7680, 4320)]
    for (w, h) in sizes:
        d = synth.solver_inputs(min(w, 1024), min(h, 1024), seed=1234)
        reps = (-(-h // d["dx"].shape[0]), -(-w // d["dx"].shape[1]), 1)
        t = {k: torch.from_numpy(np.ascontiguousarray(np.tile(v, reps)[:h, :w])).cuda() for k, v in d.items()}
        solver = gdb200.ShardedPoissonSolver(w, h)
        single = gdb200.PoissonPlan(w, h) if rank == 0 else None
        for preset in ("L2D", "L1D"):
            params = gdb200.SolverParams()
            params.setConfigPreset(preset)
            out = torch.zeros_like(t["dx"])
            times = []
            for it in range(3):
                st = gdb200.Stats()
                dist.barrier()
                solver.solve_device(t["dx"], t["dy"], t["throughput"], t["direct"], 0.2, params.cfg, out, stats=st)
                times.append(st.device_ms)
            ms = torch.tensor([min(times[1:])], device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            solver.gather(out)
            if rank == 0:
                ref = torch.empty_like(out)
                best = None
                for it in range(2):
                    st = gdb200.Stats()
                    single.solve_device(t["dx"], t["dy"], t["throughput"], t["direct"], 0.2, params.cfg, ref, stats=st)
                    best = st.device_ms if best is None else min(best, st.device_ms)
                diff = (out.double() - ref.double())
                gbs = BYTES[preset] * w * h / (ms.item() * 1e-3) / 1e9
                print(json.dumps({"size": f"{w}x{h}", "preset": preset, "gpus": world, "bands": solver.bounds, "variant": solver.plan.variant,
                                  "ms": round(ms.item(), 3), "one_gpu_ms": round(best, 3), "speedup": round(best / ms.item(), 2),
                                  "alg_GBs_all_gpus": round(gbs, 1), "frac_of_n_x_measured_peak": round(gbs / (peak * world), 3),
                                  "rmse_vs_one_gpu": float(torch.sqrt((diff * diff).mean())), "iters": [st.irls_iters, st.cg_iters]}), flush=True)
            dist.barrier()
        solver.close()
        if single is not None:
            single.close()
        del t
        torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
