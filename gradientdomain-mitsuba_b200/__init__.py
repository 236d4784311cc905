"""gdb200 — host-side mirror of the reference's interfaces for the G-PT +
screened-Poisson hot path, over the C ABI in ``include/gdb200.h``.

Everything here calls the sm_100a CUDA library ``libgdb200.so`` through ctypes.
There is no CPU implementation in this package: if the library is missing or no
CUDA device is present the calls raise.
"""
from ._ffi import lib, Gdb200Error, Stats, library_path, pinned_empty, release_workspace  # noqa: F401
from .poisson import PoissonSolver, SolverParams, poisson_solve, PoissonPlan, ShardedPoissonSolver, shard_bounds  # noqa: F401
from .gpt import GPTIntegrator, Scene, BUFFER_NAMES  # noqa: F401
from . import scenes, synth, pfm, exr, xmlscene, meshio  # noqa: F401
from .xmlscene import load_scene  # noqa: F401
