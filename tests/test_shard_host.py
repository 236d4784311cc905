"""Host side of the sharded Poisson solve (include/gdb200.h "sharded solve"): band arithmetic, the handle layout the Python
layer ships between ranks, loud failure without a GPU.  The kernels themselves are covered by tests/test_poisson_gpu.py
(shards sharing one GPU, and one shard per GPU when the box has several)."""
import os
import subprocess
import sys

import pytest

import gdb200
from gdb200 import poisson

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("h,n", [(1024, 1), (1024, 2), (1024, 8), (2160, 8), (4320, 8), (1080, 3), (100, 7), (17, 2), (16, 1)])
def test_shard_bounds_cover_the_image_in_tile_rows(h, n):
    b = gdb200.shard_bounds(h, n)
    assert len(b) == n + 1 and b[0] == 0 and b[-1] == h
    assert all(b[i] < b[i + 1] for i in range(n)), b
    assert all(v % poisson.SHARD_ROW_ALIGN == 0 for v in b[:-1]), b           # only the last band may end inside a tile row
    rows = [b[i + 1] - b[i] for i in range(n)]
    assert max(rows) - min(rows) <= 2 * poisson.SHARD_ROW_ALIGN, rows          # even up to one tile row (+ the ragged tail)


def test_shard_bounds_refuses_more_gpus_than_tile_rows():
    with pytest.raises(gdb200.Gdb200Error, match="cannot split"):
        gdb200.shard_bounds(40, 4)
    with pytest.raises(gdb200.Gdb200Error, match="cannot split"):
        gdb200.shard_bounds(40, 0)


def test_handle_size_matches_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "gdb200.h"\nint main(void){printf("%zu", sizeof(gdb200_shard_handle));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    assert int(subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout) == poisson.SHARD_HANDLE_BYTES


def test_connect_checks_the_number_of_handles_before_touching_the_library():
    class Fake(gdb200.PoissonPlan):
        def __init__(self):            # no device: only the host-side check is exercised
            self.n_ranks, self._h = 3, None
    with pytest.raises(gdb200.Gdb200Error, match="expected 3 handles"):
        Fake().connect([b"\0" * poisson.SHARD_HANDLE_BYTES] * 2)


def test_shard_creation_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    code = ("import gdb200\n"
            "try:\n"
            "    gdb200.PoissonPlan(64, 64, band=(0, 32), rank=0, n_ranks=2)\n"
            "except gdb200.Gdb200Error as e:\n"
            "    print('ERR', e)\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert "ERR" in out.stdout, out.stdout + out.stderr
