"""Image-tile sharding of the G-PT tracer across the GPUs of one box.

The reference shards 32x32 blocks over worker threads / TCP peers and merges whole blocks
under a mutex (gpt_proc.cpp:137-149).  Here each rank traces the samples whose BASE pixel
lies in its horizontal strip into a full-size accumulator film; a sample only ever splats
into rows y-2..y+2 (4 neighbours + the box filter's 1e-5 overhang), so after tracing

  1. ONE all-reduce(sum) over the packed rows around the strip boundaries completes the
     halo rows (throughput/dx/dy/direct/final value+weight planes), and
  2. rank 0 gathers the strip interiors, develops the film and runs the global Poisson solve.

Everything here is device-agnostic torch.distributed plumbing (NCCL on GPUs, gloo in the CPU
tests); the arithmetic stays in libgdb200.
"""
import torch
import torch.distributed as dist

HALO = 2   # rows a sample can reach above/below its base pixel row with the box filter (radius 0.5 + 1e-5)


def halo_rows(rfilter_radius):
    """Rows a sample can reach above/below its base pixel row: the neighbour splat lands one pixel away and the filter
    footprint spans floor(radius + 0.5) more (ImageBlock::put, imageblock.h:167-176).  Box: 2, gaussian (radius 2): 3."""
    import math
    return 1 + int(math.floor(rfilter_radius + 0.5))


def strip_rows(height, rank, world):
    """Contiguous row range [y0, y1) of `rank` (balanced to within one row)."""
    base, rem = divmod(height, world)
    y0 = rank * base + min(rank, rem)
    return y0, y0 + base + (1 if rank < rem else 0)


def boundary_rows(height, world, halo=HALO):
    """Row indices touched by more than one rank: +-halo around every strip boundary."""
    rows = []
    for r in range(1, world):
        y = strip_rows(height, r, world)[0]
        rows.extend(range(max(0, y - halo), min(height, y + halo)))
    return sorted(set(rows))


def exchange_boundaries(acc, world, group=None, halo=HALO):
    """acc: [5, H, W, 4] accumulator film of this rank. Sums the boundary rows over all ranks
    in place with a single all-reduce of the packed rows (halo = halo_rows(filter radius))."""
    if world <= 1:
        return 0
    rows = boundary_rows(acc.shape[1], world, halo)
    if not rows:
        return 0
    idx = torch.as_tensor(rows, device=acc.device)
    packed = acc.index_select(1, idx).contiguous()
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    acc.index_copy_(1, idx, packed)
    return packed.numel() * packed.element_size()


def gather_strips(acc, rank, world, group=None):
    """After exchange_boundaries: rank 0 receives every rank's strip interior so that its film is
    the complete image. Returns the bytes this rank sent."""
    if world <= 1:
        return 0
    h = acc.shape[1]
    sizes = [strip_rows(h, r, world) for r in range(world)]
    y0, y1 = sizes[rank]
    mine = acc[:, y0:y1].contiguous()
    if rank == 0:
        bufs = [torch.empty((acc.shape[0], b - a, acc.shape[2], acc.shape[3]), dtype=acc.dtype, device=acc.device)
                for (a, b) in sizes]
        dist.gather(mine, bufs, dst=0, group=group)
        for (a, b), t in zip(sizes[1:], bufs[1:]):
            acc[:, a:b].copy_(t)
        return 0
    dist.gather(mine, None, dst=0, group=group)
    return mine.numel() * mine.element_size()


def band_spec(rank, world, band_rows=16):
    """(band_rows, band_count, band_index) for gdb200_gpt_params: interleaved row bands balance the
    per-strip cost differences of contiguous strips (paths under the light are short, floor paths long)."""
    return (band_rows, world, rank)


def band_owned_rows(height, rank, world, band_rows=16):
    return [y for y in range(height) if (y // band_rows) % world == rank]


def exchange_all(acc, world, group=None):
    """Interleaved bands put a strip boundary every band_rows rows, so the halo rows are a large part of
    the film: sum the whole accumulator film with ONE all-reduce; every rank then holds the full image."""
    if world <= 1:
        return 0
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc.numel() * acc.element_size()
