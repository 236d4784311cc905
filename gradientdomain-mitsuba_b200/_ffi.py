"""ctypes binding of include/gdb200.h (the C-ABI drop-in boundary)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


class Gdb200Error(RuntimeError):
    """Raised for any non-zero status of the C ABI (the plugin shim turns the
    same status into Log(EError), which throws — reference logger.cpp:100-148)."""


class Stats(ctypes.Structure):
    _fields_ = [("device_ms", ctypes.c_double), ("h2d_ms", ctypes.c_double), ("d2h_ms", ctypes.c_double),
                ("launches", ctypes.c_int), ("irls_iters", ctypes.c_int), ("cg_iters", ctypes.c_int),
                ("reserved0", ctypes.c_int), ("samples", ctypes.c_double), ("rays", ctypes.c_double),
                ("path_vertices", ctypes.c_double), ("state_bytes", ctypes.c_double), ("bounce_ms", ctypes.c_double),
                ("generate_ms", ctypes.c_double), ("compact_ms", ctypes.c_double), ("path_bounces", ctypes.c_double),
                ("bounce_launches", ctypes.c_int), ("reserved1", ctypes.c_int),
                ("cast_ms", ctypes.c_double), ("prepare_ms", ctypes.c_double), ("resolve_ms", ctypes.c_double),
                ("primary_ms", ctypes.c_double), ("rays_cast", ctypes.c_double)]


class PoissonConfig(ctypes.Structure):
    _fields_ = [("irlsIterMax", ctypes.c_int), ("irlsRegInit", ctypes.c_float), ("irlsRegIter", ctypes.c_float),
                ("cgIterMax", ctypes.c_int), ("cgIterCheck", ctypes.c_int), ("cgTolerance", ctypes.c_float)]


def library_path():
    # GDB200_LIBRARY: developer switch to A/B-test another build of the same CUDA library
    return os.environ.get("GDB200_LIBRARY") or os.path.join(_HERE, "libgdb200.so")


_lib = None


def lib():
    """Load libgdb200.so (built in-tree by __graft_entry__.build()). No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise Gdb200Error(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(gdb200 has no CPU fallback)")
    L = ctypes.CDLL(path)
    c_float_p = ctypes.POINTER(ctypes.c_float)
    vp = ctypes.c_void_p
    L.gdb200_version.restype = ctypes.c_int
    L.gdb200_last_error.restype = ctypes.c_char_p
    L.gdb200_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.gdb200_set_device.argtypes = [ctypes.c_int]
    L.gdb200_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.gdb200_host_free.argtypes = [vp]
    L.gdb200_poisson_preset.argtypes = [ctypes.c_char_p, ctypes.POINTER(PoissonConfig)]
    L.gdb200_poisson_plan_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
    L.gdb200_poisson_plan_destroy.argtypes = [vp]
    L.gdb200_poisson_plan_destroy.restype = None
    L.gdb200_poisson_shard_create.argtypes = [ctypes.c_int] * 6 + [ctypes.POINTER(vp)]
    L.gdb200_poisson_shard_export.argtypes = [vp, vp]
    L.gdb200_poisson_shard_connect.argtypes = [vp, vp, ctypes.c_int]
    L.gdb200_poisson_plan_set_variant.argtypes = [vp, ctypes.c_int]
    L.gdb200_poisson_plan_variant.argtypes = [vp]
    L.gdb200_poisson_metrics_device.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), vp]
    L.gdb200_poisson_metrics.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    L.gdb200_poisson_solve_device.argtypes = [vp, vp, vp, vp, vp, ctypes.c_float, ctypes.POINTER(PoissonConfig),
                                              vp, vp, ctypes.POINTER(Stats)]
    L.gdb200_poisson_solve.argtypes = [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                       ctypes.c_char_p, vp, ctypes.POINTER(Stats)]
    _check_abi(L, path)
    _lib = L
    return L


def _check_abi(L, path):
    """The ctypes mirrors in this package must have the layout the library was compiled with."""
    from . import scenes
    mine = [ctypes.sizeof(t) for t in (Stats, PoissonConfig, scenes.Camera, scenes.Shape, scenes.Material, scenes.Emitter,
                                        scenes.EnvMap, scenes.SceneDesc, scenes.GPTParams, scenes.Buffers)]
    if not hasattr(L, "gdb200_abi_sizes"):
        raise Gdb200Error(f"{path} predates gdb200_abi_sizes(): rebuild it (python -c 'import __graft_entry__ as g; g.build()')")
    theirs = (ctypes.c_int * 16)()
    n = L.gdb200_abi_sizes(theirs, 16)
    if n != len(mine) or list(theirs[:n]) != mine:
        raise Gdb200Error(f"{path} was built from another include/gdb200.h (struct sizes {list(theirs[:n])} vs {mine}): rebuild it")


def check(rc):
    if rc != 0:
        raise Gdb200Error(f"gdb200 error {rc}: {lib().gdb200_last_error().decode(errors='replace')}")


def release_workspace():
    """gdb200_release_workspace(): frees the per-device wavefront scratch memory that renders keep between calls."""
    L = lib()
    L.gdb200_release_workspace.restype = None
    L.gdb200_release_workspace()


def pinned_empty(shape, dtype):
    """numpy array on page-locked host memory (gdb200_host_alloc) so device<->host copies of film buffers run at full
    PCIe / NVLink-C2C bandwidth; freed when the array is garbage-collected."""
    import weakref
    import numpy as np
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    ptr = ctypes.c_void_p()
    check(lib().gdb200_host_alloc(ctypes.byref(ptr), max(nbytes, 1)))
    buf = (ctypes.c_char * max(nbytes, 1)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib().gdb200_host_free, ctypes.c_void_p(ptr.value))
    return arr
