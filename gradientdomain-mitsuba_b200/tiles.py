"""Image-tile sharding of the G-PT tracer across the GPUs of one box.

The reference shards 32x32 blocks over worker threads / TCP peers and merges whole blocks
under a mutex (gpt_proc.cpp:137-149).  Here each rank traces the samples whose BASE pixel
lies in its horizontal strip into a full-size accumulator film; a sample only ever splats
into rows y-2..y+2 (4 neighbours + the box filter's 1e-5 overhang), so after tracing

  1. ONE all-reduce(sum) over the packed rows around the strip boundaries completes the
     halo rows (throughput/dx/dy/direct/final value+weight planes), and
  2. rank 0 receives the strip interiors, develops the film and runs the global Poisson solve.

Strips need not be equal: rebalance() moves the boundaries so that every rank's tracing time plus the work it alone does
afterwards (rank 0: develop + solve) is the same, from times measured on the previous render.

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


def even_bounds(height, world):
    """Strip boundaries [y_0 = 0, y_1, ..., y_world = height] of strip_rows."""
    return [strip_rows(height, r, world)[0] for r in range(world)] + [height]


def boundary_rows(height, world, halo=HALO, bounds=None):
    """Row indices touched by more than one rank: +-halo around every strip boundary (bounds: explicit strip boundaries,
    default the even split)."""
    bounds = even_bounds(height, world) if bounds is None else bounds
    rows = []
    for y in bounds[1:-1]:
        rows.extend(range(max(0, y - halo), min(height, y + halo)))
    return sorted(set(rows))


def rebalance(bounds, cost_ms, extra_ms=None, min_rows=4, row_cost=None):
    """Cost-balanced strip boundaries.  cost_ms[r]: measured tracing time of rank r on its current strip [bounds[r], bounds[r+1])
    (floor rows cost more than rows under the light, so even strips finish at different times); extra_ms[r]: work that rank r
    alone does after the exchange (rank 0: develop + the Poisson solve) and that the other ranks would otherwise wait out.
    Returns boundaries for which cost + extra is equal on all ranks.  Deterministic: every rank computes the same boundaries
    from the all-gathered times.

    row_cost: optional list of per-row cost estimates (length = image height) that is UPDATED IN PLACE and carried from one
    call to the next: inside each strip the estimates are scaled so that they add up to the strip's measured time, so the
    structure learnt from earlier, different boundaries is kept (without it the cost is taken as constant inside a strip and a
    strip that straddles cheap and expensive rows keeps missing its target)."""
    world = len(bounds) - 1
    extra = [0.0] * world if extra_ms is None else [float(e) for e in extra_ms]
    height = bounds[-1]
    if row_cost is None:
        row_cost = [1.0] * height
    elif len(row_cost) != height:
        raise ValueError(f"row_cost has {len(row_cost)} entries for {height} rows")
    for r in range(world):
        a, b = bounds[r], bounds[r + 1]
        have = sum(row_cost[a:b])
        scale = max(float(cost_ms[r]), 1e-6) / have if have > 0 else 0.0
        for y in range(a, b):
            row_cost[y] = row_cost[y] * scale if have > 0 else max(float(cost_ms[r]), 1e-6) / max(1, b - a)
    total = sum(row_cost) + sum(extra)
    target = total / world                       # finishing time of every rank
    # boundary r where the running cost reaches the sum of the first r+1 tracing budgets, rounded to the nearer row (cutting
    # each strip greedily below its own budget would push every strip's remainder onto the last rank)
    new, y, run, goal = [0], 0, 0.0, 0.0
    for r in range(world - 1):
        goal += max(target - extra[r], 0.0)      # tracing time ranks 0..r may spend together
        lo, hi = new[-1] + min_rows, height - (world - 1 - r) * min_rows
        while y < hi and (y < lo or run + 0.5 * row_cost[y] <= goal):
            run += row_cost[y]
            y += 1
        new.append(y)
    new.append(height)
    return new


def exchange_boundaries(acc, world, group=None, halo=HALO, bounds=None):
    """acc: [5, H, W, 4] accumulator film of this rank. Sums the boundary rows over all ranks
    in place with a single all-reduce of the packed rows (halo = halo_rows(filter radius))."""
    if world <= 1:
        return 0
    rows = boundary_rows(acc.shape[1], world, halo, bounds)
    if not rows:
        return 0
    idx = torch.as_tensor(rows, device=acc.device)
    packed = acc.index_select(1, idx).contiguous()
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    acc.index_copy_(1, idx, packed)
    return packed.numel() * packed.element_size()


def gather_strips(acc, rank, world, group=None, bounds=None, first_buffer=0):
    """After exchange_boundaries: rank 0 receives every rank's strip interior so that its film is
    the complete image. Returns the bytes this rank sent.  first_buffer = 1 leaves out the "-final" preview accumulator
    (buffer 0), which stays empty when a reconstruction will overwrite it (skip_preview)."""
    if world <= 1:
        return 0
    acc = acc[first_buffer:]
    h = acc.shape[1]
    bounds = even_bounds(h, world) if bounds is None else bounds
    sizes = [(bounds[r], bounds[r + 1]) for r in range(world)]
    y0, y1 = sizes[rank]
    if rank == 0:                                  # strips may differ in height (rebalance): point-to-point, batched so all are in flight at once
        bufs = [torch.empty((acc.shape[0], b0 - a0, acc.shape[2], acc.shape[3]), dtype=acc.dtype, device=acc.device) for (a0, b0) in sizes[1:]]
        ops = [dist.P2POp(dist.irecv, t, r, group) for r, t in zip(range(1, world), bufs)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for (a0, b0), t in zip(sizes[1:], bufs):
            acc[:, a0:b0].copy_(t)
        return 0
    mine = acc[:, y0:y1].contiguous()
    for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, 0, group)]):
        req.wait()
    return mine.numel() * mine.element_size()
