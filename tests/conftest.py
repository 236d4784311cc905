import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


# GPU tests whose code changed after this round's last hardware run.  They are collected last, so that with `-x` a surprise in
# one of them cannot hide the results of the suites that have been green on a B200 in their final form, and they carry a
# hard time limit (they synchronise kernels of several shards: a mistake there shows up as a hang, not as a wrong number).
# Round 2: the sharded solve's wait loop got a template flag (only CTA 0 may time out) after the last run that had GPU budget.
NOT_YET_RUN_ON_HARDWARE = ("test_sharded_solve", "test_one_rank_shard_is_the_plain_plan", "test_shard_arguments_are_checked")


def pytest_collection_modifyitems(config, items):
    late = [it for it in items if any(tag in it.nodeid for tag in NOT_YET_RUN_ON_HARDWARE)]
    if late:
        items[:] = [it for it in items if it not in late] + late
        for it in late:
            it.add_marker(pytest.mark.timeout(300, method="thread"))


def _make(target):
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), target], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


class Oracle:
    """ctypes view of oracle/libgdb200_oracle.so (the CPU restatement) and, when
    built, oracle/_ref/libref_poisson.so (the reference's own solver sources)."""

    def __init__(self):
        path = os.path.join(ROOT, "oracle", "libgdb200_oracle.so")
        if not os.path.exists(path):
            _make("restatement")
        self.lib = ctypes.CDLL(path)
        ref_path = os.path.join(ROOT, "oracle", "_ref", "libref_poisson.so")
        if not os.path.exists(ref_path) and os.path.isdir(REFERENCE):
            _make("ref")
        self.ref = ctypes.CDLL(ref_path) if os.path.exists(ref_path) else None

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None

    def poisson(self, dx, dy, thr, direct, alpha=0.2, preset="L1D"):
        h, w, _ = dx.shape
        out = np.empty_like(dx)
        rc = self.lib.gdb200_oracle_poisson_solve(self._p(dx), self._p(dy), self._p(thr), self._p(direct),
                                                  w, h, ctypes.c_float(alpha), preset.encode(), self._p(out))
        assert rc == 0
        return out

    def poisson_acc64(self, dx, dy, thr, direct, alpha=0.2, preset="L1D"):
        """Same algorithm with exact fp64 reduction sums: a yardstick for how far reduction
        order alone moves the result (not the reference's arithmetic)."""
        h, w, _ = dx.shape
        out = np.empty_like(dx)
        rc = self.lib.gdb200_oracle_poisson_solve_acc64(self._p(dx), self._p(dy), self._p(thr), self._p(direct),
                                                        w, h, ctypes.c_float(alpha), preset.encode(), self._p(out))
        assert rc == 0
        return out

    def gpt(self, desc, params, threads=8):
        """CPU oracle of the G-PT tracer. Returns (buffers dict like the integrator's, weights[5,h,w], counters[3])."""
        from gdb200 import scenes
        w, h = desc.camera.width, desc.camera.height
        names = (("throughput", "-throughput"), ("dx", "-dx"), ("dy", "-dy"), ("direct", "-direct"), ("preview_final", "-final"))
        out = {n: np.zeros((h, w, 3)) for _, n in names}
        B = scenes.Buffers()
        for field, n in names:
            setattr(B, field, out[n].ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        wts, cnt = np.zeros((5, h, w)), np.zeros(3)
        rc = self.lib.gdb200_oracle_gpt_render(ctypes.byref(desc), ctypes.byref(params), ctypes.byref(B),
                                               self._p(wts), self._p(cnt), threads)
        assert rc == 0
        return out, wts, cnt

    def path(self, desc, params, threads=8):
        w, h = desc.camera.width, desc.camera.height
        out = np.zeros((h, w, 3))
        assert self.lib.gdb200_oracle_path_render(ctypes.byref(desc), ctypes.byref(params), self._p(out), threads) == 0
        return out

    def poisson_ref(self, dx, dy, thr, direct, alpha=0.2, preset="L1D"):
        h, w, _ = dx.shape
        out = np.empty_like(dx)
        sec = ctypes.c_float()
        rc = self.ref.ref_poisson_solve(self._p(dx), self._p(dy), self._p(thr), self._p(direct), w, h,
                                        ctypes.c_float(alpha), preset.encode(), self._p(out), ctypes.byref(sec))
        assert rc == 0
        return out


@pytest.fixture(scope="session")
def oracle():
    return Oracle()


def rmse(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)))


class Emu:
    """ctypes view of tests/emu/libgdb200_emu.so: the tracer's per-slot *device* routines compiled for the
    host (test infrastructure, see tests/emu/gpt_emu.cpp), so device code can be checked against the oracle
    without a GPU."""

    def __init__(self):
        subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "emu")], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        self.lib = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libgdb200_emu.so"))
        self.lib.gdb200_emu_last_error.restype = ctypes.c_char_p

    def gpt(self, desc, params):
        from gdb200 import scenes
        w, h = desc.camera.width, desc.camera.height
        names = (("throughput", "-throughput"), ("dx", "-dx"), ("dy", "-dy"), ("direct", "-direct"), ("preview_final", "-final"))
        out = {n: np.zeros((h, w, 3)) for _, n in names}
        B = scenes.Buffers()
        for field, n in names:
            setattr(B, field, out[n].ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        cnt = np.zeros(6)
        rc = self.lib.gdb200_emu_gpt_render(ctypes.byref(desc), ctypes.byref(params), ctypes.byref(B),
                                            cnt.ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError(self.lib.gdb200_emu_last_error().decode())
        return out, cnt     # cnt: done slots, rays, path vertices, samples, state bytes, path bounces

    def gpt_wavefront(self, desc, params):
        """Block mode: the queued wavefront kernels (generate / compact / bounce / tail) with every CTA run by OS threads."""
        from gdb200 import scenes
        w, h = desc.camera.width, desc.camera.height
        names = (("throughput", "-throughput"), ("dx", "-dx"), ("dy", "-dy"), ("direct", "-direct"), ("preview_final", "-final"))
        out = {n: np.zeros((h, w, 3)) for _, n in names}
        B = scenes.Buffers()
        for field, n in names:
            setattr(B, field, out[n].ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        cnt = np.zeros(6)
        rc = self.lib.gdb200_emu_gpt_render_wavefront(ctypes.byref(desc), ctypes.byref(params), ctypes.byref(B),
                                                      cnt.ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError(self.lib.gdb200_emu_last_error().decode())
        return out, cnt


    def check_culling(self, desc, n_rays, seed):
        """(mismatches, hits): candidate selection of the intersection searches vs exhaustive testing on random rays."""
        out = (ctypes.c_ulonglong * 2)()
        rc = self.lib.gdb200_emu_check_culling(ctypes.byref(desc), int(n_rays), ctypes.c_ulonglong(seed), out)
        if rc != 0:
            raise RuntimeError(self.lib.gdb200_emu_last_error().decode())
        return int(out[0]), int(out[1])

    def gpt_staged(self, desc, params, grid=3):
        """Block mode of the STAGED wavefront (csrc/gpt_stages.cuh): stage / cast / compact kernels as written, `grid`
        persistent CTAs per launch run by OS threads."""
        from gdb200 import scenes
        w, h = desc.camera.width, desc.camera.height
        names = (("throughput", "-throughput"), ("dx", "-dx"), ("dy", "-dy"), ("direct", "-direct"), ("preview_final", "-final"))
        out = {n: np.zeros((h, w, 3)) for _, n in names}
        B = scenes.Buffers()
        for field, n in names:
            setattr(B, field, out[n].ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        cnt = np.zeros(6)
        rc = self.lib.gdb200_emu_gpt_render_staged(ctypes.byref(desc), ctypes.byref(params), ctypes.byref(B),
                                                   cnt.ctypes.data_as(ctypes.c_void_p), int(grid))
        if rc != 0:
            raise RuntimeError(self.lib.gdb200_emu_last_error().decode())
        return out, cnt


@pytest.fixture(scope="session")
def emu():
    return Emu()


class RefMitsuba:
    """ctypes view of oracle/_ref/libref_mitsuba.so: the REFERENCE's own G-PT integrator (src/integrators/gpt/gpt.cpp) with
    the scene, kd-tree, shapes, emitters, BSDFs, sensor, film and filters it runs on, compiled from /root/reference
    (recipe: oracle/Makefile) and driven from a gdb200_scene_desc (oracle/ref_gpt_shim.cpp)."""
    PATH = os.path.join(ROOT, "oracle", "_ref", "libref_mitsuba.so")

    def __init__(self):
        if not os.path.exists(self.PATH) and os.path.isdir(REFERENCE):
            _make("_ref/libref_mitsuba.so")
        self.lib = ctypes.CDLL(self.PATH)
        self.lib.gdbref_gpt_last_error.restype = ctypes.c_char_p

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH) or os.path.isdir(REFERENCE)

    def gpt(self, desc, params, threads=1):
        from gdb200 import scenes
        fov, rfilter = scenes.mitsuba_sensor_args(desc)
        h, w = desc.camera.height, desc.camera.width
        out = np.zeros((5, h, w, 3))
        rc = self.lib.gdbref_gpt_render(ctypes.byref(desc), ctypes.byref(params), ctypes.c_double(fov), rfilter.encode(), int(threads),
                                        out.ctypes.data_as(ctypes.c_void_p))
        if rc:
            raise RuntimeError(self.lib.gdbref_gpt_last_error().decode())
        return dict(zip(("-final", "-throughput", "-dx", "-dy", "-direct"), out))

    def li(self, desc, params):
        """GradientPathIntegrator::Li averaged per pixel (what Oracle.path restates)."""
        from gdb200 import scenes
        fov, rfilter = scenes.mitsuba_sensor_args(desc)
        out = np.zeros((desc.camera.height, desc.camera.width, 3))
        rc = self.lib.gdbref_gpt_li_render(ctypes.byref(desc), ctypes.byref(params), ctypes.c_double(fov), rfilter.encode(),
                                           out.ctypes.data_as(ctypes.c_void_p))
        if rc:
            raise RuntimeError(self.lib.gdbref_gpt_last_error().decode())
        return out


class RefGbdpt:
    """ctypes view of oracle/_ref/libref_gbdpt.so: the REFERENCE's own G-BDPT integrator (src/integrators/gbdpt over
    src/libbidir) rendered through a real RenderJob on Mitsuba's scheduler (oracle/ref_gbdpt_shim.cpp); returns the seven
    buffers MultiFilm writes for it (gbdpt.cpp:164) as float32 arrays."""
    PATH = os.path.join(ROOT, "oracle", "_ref", "libref_gbdpt.so")
    NAMES = ("-L1", "-L2", "-gradientNegY", "-gradientNegX", "-gradientPosX", "-gradientPosY", "-primal")

    def __init__(self):
        if not os.path.exists(self.PATH) and os.path.isdir(REFERENCE):
            _make("_ref/libref_gbdpt.so")
        ctypes.CDLL(RefMitsuba.PATH, mode=ctypes.RTLD_GLOBAL)
        self.lib = ctypes.CDLL(self.PATH)
        self.lib.gdbref_gbdpt_last_error.restype = ctypes.c_char_p

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH) or os.path.isdir(REFERENCE)

    def render(self, desc, params, light_image=True, alpha=0.2, threads=4):
        import tempfile
        from gdb200 import scenes, pfm
        fov, rfilter = scenes.mitsuba_sensor_args(desc)
        with tempfile.TemporaryDirectory() as d:
            dest = os.path.join(d, "out")
            rc = self.lib.gdbref_gbdpt_render(ctypes.byref(desc), ctypes.byref(params), ctypes.c_double(fov), rfilter.encode(),
                                              int(light_image), ctypes.c_double(alpha), int(threads), dest.encode())
            if rc:
                raise RuntimeError(self.lib.gdbref_gbdpt_last_error().decode())
            return {n: pfm.read_pfm(dest + n + ".pfm") for n in self.NAMES}
