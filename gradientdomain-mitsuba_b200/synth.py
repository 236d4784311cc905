"""Deterministic synthetic inputs for the solver sweep.

Pure integer hashing + Box-Muller in numpy, so the same bytes come out on every
machine and numpy version (the golden checksums under tests/golden depend on it).
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform(shape, seed, stream):
    """float64 uniforms in (0,1), a pure function of (index, seed, stream)."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        key = _splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        bits = _splitmix64(idx ^ key)
    return ((bits >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def normal(shape, seed, stream):
    u1 = uniform(shape, seed, 2 * stream)
    u2 = uniform(shape, seed, 2 * stream + 1)
    return (np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).reshape(shape)


def clean_image(w, h):
    """clean(x,y,c) = 0.5 + 0.25 sin(0.02x + 0.01y + c) + 0.2 [(x//64 + y//64) odd]."""
    y, x, c = np.meshgrid(np.arange(h), np.arange(w), np.arange(3), indexing="ij")
    return 0.5 + 0.25 * np.sin(0.02 * x + 0.01 * y + c) + 0.2 * (((x // 64) + (y // 64)) % 2)


def solver_inputs(w, h, seed=1234, noise_primal=0.2, noise_grad=0.02, last_col_nonzero=False):
    """Returns dict of float32 (h,w,3) arrays: clean, throughput, dx, dy, direct."""
    clean = clean_image(w, h)
    thr = clean + noise_primal * normal((h, w, 3), seed, 0)
    dx = np.zeros_like(clean)
    dy = np.zeros_like(clean)
    dx[:, :-1] = clean[:, 1:] - clean[:, :-1]
    dy[:-1] = clean[1:] - clean[:-1]
    ndx = noise_grad * normal((h, w, 3), seed, 1)
    ndy = noise_grad * normal((h, w, 3), seed, 2)
    if not last_col_nonzero:   # the renderer leaves a (weight-N) value there; one variant keeps it
        ndx[:, -1] = 0.0
        ndy[-1] = 0.0
    dx += ndx
    dy += ndy
    direct = 0.05 * uniform((h, w, 3), seed, 7).reshape(h, w, 3)
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)  # noqa: E731
    return {"clean": f(clean), "throughput": f(thr), "dx": f(dx), "dy": f(dy), "direct": f(direct)}


def fixed_point_inputs(w, h, seed=7):
    """dx,dy = fp32 forward differences of throughput (last column/row 0): the
    solver must return throughput+direct bit-exactly."""
    thr = np.ascontiguousarray(0.5 + 0.3 * normal((h, w, 3), seed, 0), dtype=np.float32)
    dx = np.zeros_like(thr)
    dy = np.zeros_like(thr)
    dx[:, :-1] = thr[:, 1:] - thr[:, :-1]
    dy[:-1] = thr[1:] - thr[:-1]
    direct = np.ascontiguousarray(uniform((h, w, 3), seed, 3).reshape(h, w, 3), dtype=np.float32)
    return {"throughput": thr, "dx": dx, "dy": dy, "direct": direct}
