"""Host-side mirror of the reference's `gpt` integrator plugin
(GradientPathIntegrator, src/integrators/gpt/gpt.cpp:1191-1211, 1358-1480) over the C ABI.

Same parameter names, defaults and validation errors as the XML plugin; `render()` does what
GradientPathIntegrator::render does after scene loading: trace the five buffers
("-final", "-throughput", "-dx", "-dy", "-direct", gpt.cpp:1380), then — unless both
reconstructL1 and reconstructL2 are off — run the screened-Poisson reconstruction on the
developed fp32 buffers (gpt.cpp:1415-1477) and replace "-final" with its result.
All arithmetic runs in libgdb200.so on the GPU.
"""
import ctypes

import numpy as np

from ._ffi import lib, check, Stats, PoissonConfig, Gdb200Error
from . import scenes as _scenes

BUFFER_NAMES = ("-final", "-throughput", "-dx", "-dy", "-direct")   # gpt.cpp:1380


def _bind(L):
    vp = ctypes.c_void_p
    if getattr(L, "_gpt_bound", False):
        return
    L.gdb200_scene_create.argtypes = [ctypes.POINTER(_scenes.SceneDesc), ctypes.POINTER(vp)]
    L.gdb200_scene_destroy.argtypes = [vp]
    L.gdb200_scene_destroy.restype = None
    L.gdb200_gpt_render.argtypes = [vp, ctypes.POINTER(_scenes.GPTParams), ctypes.POINTER(_scenes.Buffers),
                                    ctypes.POINTER(Stats)]
    L.gdb200_gpt_solver_inputs.argtypes = [vp] + [ctypes.POINTER(vp)] * 4
    L.gdb200_gpt_accumulators.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t)]
    L.gdb200_gpt_develop.argtypes = [vp, ctypes.POINTER(_scenes.Buffers)]
    L.gdb200_cancel.argtypes = [vp]
    L.gdb200_cancel.restype = None
    L._gpt_bound = True


def device_view(ptr, shape, typestr="<f8"):
    """Zero-copy torch view of a raw device pointer owned by libgdb200 (plumbing for NCCL)."""
    import torch

    class _Dev:
        __cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Dev(), device="cuda")


class Scene:
    """Device-resident flattened scene (gdb200_scene)."""

    def __init__(self, desc):
        L = lib()
        _bind(L)
        self.desc = desc
        self.width, self.height = desc.camera.width, desc.camera.height
        self._h = ctypes.c_void_p()
        check(L.gdb200_scene_create(ctypes.byref(desc), ctypes.byref(self._h)))

    def accumulators(self):
        """torch view [5, h, w, 4] (fp64: R,G,B,weight) of the raw film accumulators."""
        ptr, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
        check(lib().gdb200_gpt_accumulators(self._h, ctypes.byref(ptr), ctypes.byref(nbytes)))
        return device_view(ptr.value, (5, self.height, self.width, 4))

    def develop(self, download=True, out=None):
        """Re-develop after the accumulators were merged across GPUs (gdb200_gpt_develop)."""
        out, B = ({} if out is None else out), _scenes.Buffers()
        if download:
            for field, name in (("preview_final", "-final"), ("throughput", "-throughput"), ("dx", "-dx"),
                                ("dy", "-dy"), ("direct", "-direct")):
                if name not in out:
                    out[name] = np.empty((self.height, self.width, 3), dtype=np.float64)
                setattr(B, field, out[name].ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        check(lib().gdb200_gpt_develop(self._h, ctypes.byref(B)))
        return out

    def close(self):
        if self._h:
            lib().gdb200_scene_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GPTIntegrator:
    """`<integrator type="gpt">` look-alike."""

    def __init__(self, maxDepth=-1, minDepth=-1, rrDepth=5, strictNormals=False, hideEmitters=False,
                 shiftThreshold=0.001, reconstructL1=True, reconstructL2=False, reconstructAlpha=0.2):
        # validation and messages of gpt.cpp:1203-1210, integrator.cpp:221,224
        if reconstructL1 and reconstructL2:
            raise Gdb200Error("Disable 'reconstructL1' or 'reconstructL2': Cannot display two reconstructions at a time!")
        if reconstructAlpha <= 0.0:
            raise Gdb200Error("'reconstructAlpha' must be set to a value greater than zero!")
        if maxDepth <= 0 and maxDepth != -1:
            raise Gdb200Error("'maxDepth' must be set to -1 (infinite) or a value greater than zero!")
        if rrDepth <= 0:
            raise Gdb200Error("'rrDepth' must be set to a value greater than zero!")
        self.maxDepth, self.rrDepth, self.strictNormals = maxDepth, rrDepth, strictNormals
        self.minDepth = 1                      # read, then forced to 1 (gpt.cpp:1369)
        self.hideEmitters = hideEmitters
        self.shiftThreshold = shiftThreshold
        self.reconstructL1, self.reconstructL2, self.reconstructAlpha = reconstructL1, reconstructL2, reconstructAlpha
        self.stats, self.solver_stats = Stats(), Stats()
        # not XML parameters of the reference: gdb200_gpt_params.flags / max_slots (include/gdb200.h)
        self.refUninitMeasure = False      # reproduce the compiled reference at gpt.cpp:957 (parity tests against it)
        self.fusedBounce = False           # round-1 single-kernel bounce (A/B measurements)
        self.maxSlots = 0                  # resident path slots, 0 = library default

    def params(self, spp, seed=0, rows=None, bands=None, preview=True, streams=1):
        p = _scenes.default_params(spp=spp, seed=seed, max_depth=self.maxDepth, rr_depth=self.rrDepth,
                                   shift_threshold=self.shiftThreshold, strict_normals=self.strictNormals)
        if rows is not None:
            p.y_begin, p.y_end = rows
        p.skip_preview = 0 if preview else 1
        if bands is not None:                      # (band_rows, band_count, band_index)
            p.band_rows, p.band_count, p.band_index = bands
        p.streams_per_pixel = streams              # sample streams per pixel (gdb200_gpt_params.streams_per_pixel)
        p.flags = (_scenes.GPT_REF_UNINIT_MEASURE if self.refUninitMeasure else 0) | (_scenes.GPT_FUSED_BOUNCE if self.fusedBounce else 0)
        p.max_slots = self.maxSlots
        return p

    def trace(self, scene, spp, seed=0, rows=None, download=True, bands=None, preview=True, streams=1, out=None):
        """The sampling part of render(): returns the developed fp64 buffers (h,w,3).  `out`: optional dict of
        preallocated (h,w,3) float64 arrays to download into (e.g. gdb200.pinned_empty buffers reused across renders)."""
        if self.hideEmitters:   # gpt.cpp:1362-1365
            raise Gdb200Error("Option 'hideEmitters' not implemented for Gradient-Domain Path Tracing!")
        h, w = scene.height, scene.width
        out = {} if out is None else out
        B = _scenes.Buffers()
        if download:
            for field, name in (("preview_final", "-final"), ("throughput", "-throughput"), ("dx", "-dx"),
                                ("dy", "-dy"), ("direct", "-direct")):
                if name not in out:
                    out[name] = np.empty((h, w, 3), dtype=np.float64)
                a = out[name]
                if a.shape != (h, w, 3) or a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
                    raise Gdb200Error(f"output buffer {name} must be a C-contiguous float64 array of shape {(h, w, 3)}")
                setattr(B, field, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        p = self.params(spp, seed, rows, bands, preview, streams)
        check(lib().gdb200_gpt_render(scene._h, ctypes.byref(p), ctypes.byref(B), ctypes.byref(self.stats)))
        return out

    def reconstruct(self, scene, plan=None, download=True):
        """gpt.cpp:1415-1477 on the device-resident developed buffers. Returns fp32 (h,w,3) or None."""
        if not (self.reconstructL1 or self.reconstructL2):
            return None
        import torch
        from .poisson import PoissonPlan
        L = lib()
        ptrs = [ctypes.c_void_p() for _ in range(4)]
        check(L.gdb200_gpt_solver_inputs(scene._h, *[ctypes.byref(x) for x in ptrs]))
        d_dx, d_dy, d_thr, d_direct = (x.value for x in ptrs)
        cfg = PoissonConfig()
        check(L.gdb200_poisson_preset(b"L1D" if self.reconstructL1 else b"L2D", ctypes.byref(cfg)))   # gpt.cpp:1447-1451
        own = plan is None
        plan = plan or PoissonPlan(scene.width, scene.height)
        out = torch.empty((scene.height, scene.width, 3), dtype=torch.float32, device="cuda")
        plan.solve_device(d_dx, d_dy, d_thr, d_direct, float(self.reconstructAlpha), cfg, out, stats=self.solver_stats)
        res = out.cpu().numpy() if download else out
        if own:
            plan.close()
        return res

    def render(self, scene, spp, seed=0, streams=1, out=None, plan=None):
        """Returns {"-final","-throughput","-dx","-dy","-direct"} like the five multifilm buffers.  `out` / `plan`:
        optional preallocated host buffers (see trace) and solver workspace to reuse across renders."""
        out = self.trace(scene, spp, seed, preview=not (self.reconstructL1 or self.reconstructL2), streams=streams, out=out)
        final = self.reconstruct(scene, plan)
        if final is not None:
            out["-final"][...] = final                   # setBitmapMulti(reconstruction, BUFFER_FINAL), gpt.cpp:1468-1475
        return out

    def save(self, dest, buffers, file_format="pfm", component_format="float16"):
        """MultiFilm::develop (multifilm.cpp:423-516): writes <dest>-final, -throughput, -dx, -dy, -direct as ".pfm"
        (float32 RGB, bottom-up scanlines) or, for file_format "openexr" (the film's default), ".exr" (RGB, float16 or
        float32, ZIP) and returns the paths."""
        if file_format == "openexr":
            from . import exr
            return exr.save_multifilm(dest, buffers, component_format)
        if file_format != "pfm":
            raise ValueError("file_format must be \"openexr\" or \"pfm\"")
        from . import pfm
        return pfm.save_multifilm(dest, buffers)
