// gdb200 screened-Poisson reconstruction for sm_100a: the whole IRLS-over-CG
// solve of the reference (src/integrators/poisson_solver/Solver.cpp:374-509,
// vector ops Backend.cpp:154-376) as ONE persistent cooperative kernel.
//
// Design (not a port of BackendCUDA.cu's 10 small kernels + host syncs):
//   * planar SoA fp32 planes with a 4-pixel pitch, every access a 16-byte
//     LDG/STG.128; interleaved RGB only at the import/export edges (fused).
//   * 64x16-pixel thread tiles, tiles dealt round-robin to a grid of exactly the
//     co-resident CTA count, so the 5-point stencil's halo rows are L1/L2 hits and
//     each plane crosses HBM once per phase.
//   * per CG iteration two phases / two grid barriers instead of three kernels:
//       A: p = r + b*p_old (recomputed on the halo), x += a_prev*p_old,
//          Ap = P'W2P p, pAp            (reads r,p_old,x,w2; writes p,x,Ap)
//       B: r -= a*Ap, rz                (reads r,Ap; writes r)
//     = 120 B/pixel against the reference's 132.
//   * `e = b - Px` is never stored: w2 and r = P'W2 e recompute it from b and x
//     (IRLS prologue 156 B/pixel against the reference's 336).
//   * reductions: fp32 per thread, fp64 warp/block/grid tree in a fixed order
//     (deterministic, independent of timing); w2 normalisation, alpha/beta and
//     the cgIterCheck/cgTolerance early-out all stay on the device.
// Per-pixel arithmetic keeps the reference's operation order and is compiled
// with -fmad=false, so differences against the CPU reference come from
// reduction order only.
#ifndef GDB200_EMU          // tests/emu/poisson_emu.cpp compiles the device part of this source for the host (test infrastructure)
#include "common.h"
#include <cooperative_groups.h>
#endif
#include <cfloat>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <unistd.h>

namespace cg = cooperative_groups;

namespace gdb200 {

constexpr int kThreads = 256;
constexpr int kTileGX  = 16;   // groups of 4 pixels per tile row  (64 px)
constexpr int kTileY   = 16;   // rows per tile
static_assert(kTileGX * kTileY == kThreads, "one thread per 4-pixel group");
// Resident variant (images of <= kResTiles tiles per CTA, e.g. 1024x1024 on 148 SMs x 4 CTAs): a CTA works on the same tiles in
// every phase, so the CG vectors only their own thread ever touches -- x and Ap -- stay in shared memory for the whole CG loop
// (2 tiles x 2 vectors x 3 channels x 256 threads x 16 B = 48 KB per CTA): 72 instead of 120 B of L2 traffic per pixel and CG
// iteration, and a working set (r, p, p', w2) that fits one L2 partition.  Same arithmetic, same reduction order, same bits.
// Kernel variants: 0 = streaming (everything through L2), 1 = x and Ap resident for <= 2 tiles per CTA, 2 = x resident for
// <= 4 tiles per CTA (1920x1080 on a B200: 96 instead of 120 B), 3 = streaming with the shared-tile exchange of phase A.
constexpr int kResBytes = 2 * 2 * 3 * kThreads * 16;      // 48 KB either way: 2 tiles x {x, Ap} or 4 tiles x {x}
__host__ __device__ constexpr int res_tiles(int mode) { return mode == 1 ? 2 : mode == 2 ? 4 : 0; }
__host__ __device__ constexpr bool res_x(int mode) { return mode == 1 || mode == 2; }
__host__ __device__ constexpr bool res_ap(int mode) { return mode == 1; }

enum Plane {
    B0 = 0, BX = 3, BY = 6,      // b = [alpha*throughput; dx; dy]   (Solver.cpp:321-329)
    X = 9, R = 12, PA = 15, PB = 18, AP = 21,
    W0 = 24, WX = 25, WY = 26,   // w2, normalised
    V0 = 27, VX = 28, VY = 29,   // 1/(|e|+reg) before normalisation
    kPlanes = 30
};

constexpr int kMaxRanks = 16;

// One reduction message of a rank: its three partial sums, each split into two 8-byte words that carry 32 bits of the double
// and the 32-bit number of the reduction (stores of 8 bytes are atomic, so a word is either old or complete: the receiver needs
// no separate flag and the sender no fence between data and flag -- the "LL" idea of NCCL's low-latency protocol).
struct __align__(64) Mail {
    unsigned long long w[8];          // w[2c] = hi32(v_c) : number, w[2c+1] = lo32(v_c) : number; w[6], w[7] unused
};

struct PoissonArgs {
    int W, H, Wp, Gx, tilesX, nTiles, aosVec;
    // Row band [y0, y1) of the image this GPU solves (the whole image on one GPU).  The planes hold rows y0-1 .. y1: local row
    // 0 and local row y1-y0+1 are halo rows, refreshed by the neighbouring GPU with peer stores (see "sharded solve").
    int y0, y1;
    int rank, nRanks;
    float *peerUp, *peerDown;                 // plane arrays of the GPUs holding the bands above / below (peer memory) or NULL
    size_t peerUpElems, peerDownElems;        // their plane sizes in floats
    int peerUpRows;                           // rows of the band above (its lower halo is its local row peerUpRows + 1)
    Mail *mail[kMaxRanks];                    // mail[r]: rank r's mailbox [2][kMaxRanks] (+ its two relay slots); mail[rank] is local
    unsigned epochBase;                       // reductions of all earlier solves: message numbers never repeat within 2^32
    unsigned *status;                         // [0] != 0: a peer did not arrive in time (the solve gives up instead of hanging);
                                              // [1]: number of the last reduction (the next solve's epochBase)
    float alpha;
    gdb200_poisson_config cfg;
    float *plane[kPlanes];
    const float *in_dx, *in_dy, *in_thr, *in_direct;
    float *out_final;
    double *red;      // [2][grid][3]
    int *iters;       // [0] IRLS iterations, [1] CG iterations
};

struct F4 { float v[4]; };

__device__ __forceinline__ F4 ld4(const float *p)
{
    float4 t = *reinterpret_cast<const float4 *>(p);
    return F4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ void st4(float *p, const F4 &a)
{
    *reinterpret_cast<float4 *>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
__device__ __forceinline__ F4 zero4() { return F4{{0.f, 0.f, 0.f, 0.f}}; }

// Deterministic sum of three per-thread doubles over the grid -- and, in a sharded solve, over the grids of all GPUs; the result
// is returned to THREAD 0 of every CTA only (it derives the CG scalars and publishes them in shared memory, see CgScalars).
// Contains one grid barrier.  After the barrier EVERY thread fetches its share of the per-CTA partials (<= 3 independent loads
// per component at 592 CTAs: one L2 round trip instead of a 19-step loop on one warp), then a fixed-order warp / block tree.
// The shared scratch is indexed by the call's parity so that a call needs no trailing __syncthreads.
//
// Sharded solve (nRanks > 1): CTA 0 of every GPU then posts its GPU's sums into every GPU's mailbox with 32-byte peer stores
// over NVLink, and thread 0 of every CTA waits for the nRanks messages of this reduction in its OWN GPU's mailbox (local
// polling) and adds them in rank order: the same bits on every GPU, so all of them take the same branches.  The message is also
// the barrier that publishes the halo rows pushed during the phase (push_row: plain peer stores; the grid barrier orders them
// before CTA 0's system-scope fence, which orders them before the message).  Two mailbox slots by parity: a GPU can be at most
// one reduction ahead of the slowest one.
struct SyncState {
    int parity = 0;
    unsigned epoch = 0;               // number of reductions so far (+ PoissonArgs::epochBase = the message number)
    bool dead = false;                // a wait timed out: stop waiting, the host reports the failure
};

template <bool SYS> __device__ __forceinline__ void post_mail(Mail *m, const double v[3], unsigned long long number)
{
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v[c]);
        const unsigned long long hi = (bits & 0xffffffff00000000ull) | number, lo = (bits << 32) | number;
#ifdef GDB200_EMU
        __atomic_store_n(&m->w[2 * c], hi, __ATOMIC_RELAXED);
        __atomic_store_n(&m->w[2 * c + 1], lo, __ATOMIC_RELAXED);
#else
        if (SYS) {
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(&m->w[2 * c]), "l"(hi) : "memory");
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(&m->w[2 * c + 1]), "l"(lo) : "memory");
        } else {
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(&m->w[2 * c]), "l"(hi) : "memory");
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(&m->w[2 * c + 1]), "l"(lo) : "memory");
        }
#endif
    }
}

// Polls a message slot until all six words carry `number`; false (and dead = true) after kShardTimeoutNs.  A dead solve stops
// waiting: it runs to its end on whatever is in the mailboxes and the host reports the failure.
#ifdef GDB200_EMU
constexpr unsigned long long kShardTimeoutNs = 3000000000ull;       // the host emulation's peer-missing test should not take long
#else
constexpr unsigned long long kShardTimeoutNs = 8000000000ull;
#endif
__device__ __forceinline__ unsigned long long global_timer_ns()
{
#ifdef GDB200_EMU
    return emu_timer_ns();
#else
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#endif
}
// BOUNDED = false (the relay slot): only CTA 0 may give up -- it then relays whatever the mailbox holds, so that every CTA of the
// grid still computes with the same numbers and takes the same branches.  (CTAs that timed out on their own read different
// garbage, left the CG loop at different iterations and deadlocked the grid barrier: seen once on the GPU box.)
template <bool SYS, bool BOUNDED> __device__ __forceinline__ bool wait_mail(const Mail *m, unsigned long long number, double v[3], bool &dead)
{
    unsigned long long w[6], since = 0;
    unsigned spins = 0;
    bool ok = true;
    for (;;) {
#pragma unroll
        for (int i = 0; i < 6; i += 2) {
#ifdef GDB200_EMU
            w[i] = __atomic_load_n(&m->w[i], __ATOMIC_RELAXED); w[i + 1] = __atomic_load_n(&m->w[i + 1], __ATOMIC_RELAXED);
#else
            if (SYS) asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[i]), "=l"(w[i + 1]) : "l"(&m->w[i]) : "memory");
            else     asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[i]), "=l"(w[i + 1]) : "l"(&m->w[i]) : "memory");
#endif
        }
        bool all = true;
#pragma unroll
        for (int i = 0; i < 6; i++) all = all && (w[i] & 0xffffffffull) == number;
        if (all || dead) break;
        if (BOUNDED && (++spins & 1023u) == 0) {
            const unsigned long long now = global_timer_ns();
            if (since == 0) since = now;
            else if (now - since > kShardTimeoutNs) { dead = true; ok = false; break; }
        }
        __nanosleep(32);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) v[c] = __longlong_as_double((long long)((w[2 * c] & 0xffffffff00000000ull) | (w[2 * c + 1] >> 32)));
    return ok;
}

template <bool SHARD>
__device__ void grid_sum3(cg::grid_group &grid, const PoissonArgs &a, SyncState &sync, double a0, double a1,
                          double a2, float out[3])
{
    __shared__ double s_part[2][kThreads / 32][3];
    __shared__ double s_tot[2][kThreads / 32][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, par = sync.parity;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) { s_part[par][warp][0] = a0; s_part[par][warp][1] = a1; s_part[par][warp][2] = a2; }
    __syncthreads();
    const int G = (int)gridDim.x;
    double *slot = a.red + (size_t)par * G * 3;      // [3][G]
    if (threadIdx.x < 3) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) s += s_part[par][w][threadIdx.x];
        slot[(size_t)threadIdx.x * G + blockIdx.x] = s;
    }
    grid.sync();
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int b = threadIdx.x; b < G; b += kThreads) {
        t0 += __ldcg(slot + b);
        t1 += __ldcg(slot + G + b);
        t2 += __ldcg(slot + 2 * G + b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        t0 += __shfl_xor_sync(0xffffffffu, t0, o);
        t1 += __shfl_xor_sync(0xffffffffu, t1, o);
        t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    if (lane == 0) { s_tot[par][warp][0] = t0; s_tot[par][warp][1] = t1; s_tot[par][warp][2] = t2; }
    __syncthreads();
    sync.parity ^= 1;
    sync.epoch++;
    if (threadIdx.x != 0) return;
    t0 = t1 = t2 = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) { t0 += s_tot[par][w][0]; t1 += s_tot[par][w][1]; t2 += s_tot[par][w][2]; }
    if (SHARD) {
        const unsigned long long number = a.epochBase + sync.epoch;        // 32 bits
        Mail *relay = a.mail[a.rank] + (2 + par) * kMaxRanks;              // this GPU's own two relay slots, behind the mailbox
        if (blockIdx.x == 0) {
            // Everything this GPU pushed into its neighbours' halo rows during the phase happens-before the grid barrier above;
            // this fence orders it before the message (cumulativity: one system-scope fence after an intra-GPU barrier, the
            // shape of cooperative groups' own multi-device barrier).
            __threadfence_system();
            const double tv[3] = {t0, t1, t2};
            for (int r = 0; r < a.nRanks; r++) post_mail<true>(a.mail[r] + par * kMaxRanks + a.rank, tv, number);
            double sum[3] = {0.0, 0.0, 0.0};
            for (int r = 0; r < a.nRanks; r++) {
                double v[3];
                if (!wait_mail<true, true>(a.mail[a.rank] + par * kMaxRanks + r, number, v, sync.dead)) atomicExch(a.status, 1u + (unsigned)r);
                sum[0] += v[0]; sum[1] += v[1]; sum[2] += v[2];
            }
            // ONE system-scope acquire per GPU (592 of them, one per CTA, cost 6 us per reduction: they queue up per SM), then
            // the grid total goes to the other CTAs through a local relay slot at GPU scope.
            __threadfence_system();
            post_mail<false>(relay, sum, number);
            t0 = sum[0]; t1 = sum[1]; t2 = sum[2];
        } else {
            double v[3];
            wait_mail<false, false>(relay, number, v, sync.dead);
            __threadfence();         // acquire at GPU scope: the halo rows read after the caller's __syncthreads are the pushed ones
            t0 = v[0]; t1 = v[1]; t2 = v[2];
        }
    }
    out[0] = (float)t0; out[1] = (float)t1; out[2] = (float)t2;
}

struct TileIter {
    int y, x0, idx;
    bool valid;
};

template <bool SHARD = true>
__device__ __forceinline__ TileIter tile_thread(const PoissonArgs &a, int tile)
{
    const int y0 = SHARD ? a.y0 : 0, y1 = SHARD ? a.y1 : a.H;     // one GPU: the band is the image
    const int tx = tile % a.tilesX;
    int ty = tile / a.tilesX;
    if (SHARD) {
        // the band's first and last tile rows first: their peer stores are long done when the phase ends
        const int tilesY = a.nTiles / a.tilesX;
        ty = ty == 0 ? 0 : (ty == 1 ? tilesY - 1 : ty - 1);
    }
    const int gx = tx * kTileGX + (threadIdx.x % kTileGX);
    TileIter t;
    t.y = y0 + ty * kTileY + (threadIdx.x / kTileGX);
    t.x0 = gx * 4;
    t.valid = gx < a.Gx && t.y < y1;
    t.idx = (t.y - y0 + 1) * a.Wp + t.x0;
    return t;
}

// The 4-pixel groups of the band's two halo rows (y0-1 if it exists, then y1 if it exists), dealt over the whole grid.
template <bool SHARD, class F> __device__ __forceinline__ void for_halo_rows(const PoissonArgs &a, bool above, bool below, F f)
{
    if (!SHARD) return;
    const int n = a.Gx * 2;
    for (int g = blockIdx.x * kThreads + threadIdx.x; g < n; g += gridDim.x * kThreads) {
        const bool up = g < a.Gx;
        if (up ? !(above && a.y0 > 0) : !(below && a.y1 < a.H)) continue;
        TileIter t;
        t.y = up ? a.y0 - 1 : a.y1;
        t.x0 = (up ? g : g - a.Gx) * 4;
        t.valid = true;
        t.idx = (t.y - a.y0 + 1) * a.Wp + t.x0;
        f(t);
    }
}

// Sharded solve: a value this GPU wrote into the first / last row of its band goes to the neighbour's halo row as well (a
// 16-byte store over NVLink, fire and forget; made visible by the reduction that ends the phase, grid_sum3).
template <bool SHARD>
__device__ __forceinline__ void push_row(const PoissonArgs &a, const TileIter &t, int plane, const F4 &v)
{
    if (!SHARD) return;
    if (a.peerUp && t.y == a.y0) st4(a.peerUp + a.peerUpElems * plane + (size_t)(a.peerUpRows + 1) * a.Wp + t.x0, v);
    if (a.peerDown && t.y == a.y1 - 1) st4(a.peerDown + a.peerDownElems * plane + t.x0, v);
}

// ---- interleaved RGB <-> planar ------------------------------------------------------------
__device__ __forceinline__ void load_rgb4(const float *aos, const PoissonArgs &a, const TileIter &t,
                                          F4 &r, F4 &g, F4 &b)
{
    const size_t base = ((size_t)t.y * a.W + t.x0) * 3;
    if (a.aosVec) {
        F4 f0 = ld4(aos + base), f1 = ld4(aos + base + 4), f2 = ld4(aos + base + 8);
        r = F4{{f0.v[0], f0.v[3], f1.v[2], f2.v[1]}};
        g = F4{{f0.v[1], f1.v[0], f1.v[3], f2.v[2]}};
        b = F4{{f0.v[2], f1.v[1], f2.v[0], f2.v[3]}};
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool in = t.x0 + j < a.W;
            r.v[j] = in ? aos[base + 3 * j + 0] : 0.f;
            g.v[j] = in ? aos[base + 3 * j + 1] : 0.f;
            b.v[j] = in ? aos[base + 3 * j + 2] : 0.f;
        }
    }
}

__device__ __forceinline__ void store_rgb4(float *aos, const PoissonArgs &a, const TileIter &t,
                                           const F4 &r, const F4 &g, const F4 &b)
{
    const size_t base = ((size_t)t.y * a.W + t.x0) * 3;
    if (a.aosVec) {
        st4(aos + base,     F4{{r.v[0], g.v[0], b.v[0], r.v[1]}});
        st4(aos + base + 4, F4{{g.v[1], b.v[1], r.v[2], g.v[2]}});
        st4(aos + base + 8, F4{{b.v[2], r.v[3], g.v[3], b.v[3]}});
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (t.x0 + j < a.W) {
                aos[base + 3 * j + 0] = r.v[j];
                aos[base + 3 * j + 1] = g.v[j];
                aos[base + 3 * j + 2] = b.v[j];
            }
    }
}

// ---- phase: import  (Solver.cpp:321-337) ----------------------------------------------------
__device__ __forceinline__ void import_group(const PoissonArgs &a, const TileIter &t)
{
    {
        F4 c[3];
        if (a.in_thr) load_rgb4(a.in_thr, a, t, c[0], c[1], c[2]);
        else c[0] = c[1] = c[2] = zero4();
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            F4 b0;
#pragma unroll
            for (int j = 0; j < 4; j++) b0.v[j] = c[ch].v[j] * a.alpha;
            st4(a.plane[B0 + ch] + t.idx, b0);
            st4(a.plane[X + ch] + t.idx, c[ch]);
            st4(a.plane[PA + ch] + t.idx, zero4());
        }
        load_rgb4(a.in_dx, a, t, c[0], c[1], c[2]);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) st4(a.plane[BX + ch] + t.idx, c[ch]);
        load_rgb4(a.in_dy, a, t, c[0], c[1], c[2]);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) st4(a.plane[BY + ch] + t.idx, c[ch]);
    }
}

template <bool SHARD>
__device__ void phase_import(const PoissonArgs &a)
{
    for (int tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
        const TileIter t = tile_thread<SHARD>(a, tile);
        if (t.valid) import_group(a, t);
    }
    // sharded: the inputs are whole images on every GPU, so each one imports its own halo rows
    for_halo_rows<SHARD>(a, true, true, [&](const TileIter &t) { import_group(a, t); });
}

// e = b - P x at this thread's 4 pixels, one channel (Backend.cpp:165-186, :256-272 with a=-1).
__device__ __forceinline__ void residual4(const PoissonArgs &a, const TileIter &t, int ch,
                                          F4 &e0, F4 &ex, F4 &ey)
{
    const float *x = a.plane[X + ch];
    const F4 xc = ld4(x + t.idx);
    const F4 b0 = ld4(a.plane[B0 + ch] + t.idx);
    const F4 bx = ld4(a.plane[BX + ch] + t.idx);
    const F4 by = ld4(a.plane[BY + ch] + t.idx);
    const float xr = (t.x0 + 4 < a.W) ? x[t.idx + 4] : 0.f;
    const F4 xd = (t.y != a.H - 1) ? ld4(x + t.idx + a.Wp) : zero4();
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int xx = t.x0 + j;
        const float xi = xc.v[j];
        const float nx = (j == 3) ? xr : xc.v[j + 1];
        const float p0 = xi * a.alpha;
        const float p1 = (xx != a.W - 1) ? nx - xi : 0.f;
        const float p2 = (t.y != a.H - 1) ? xd.v[j] - xi : 0.f;
        e0.v[j] = b0.v[j] - p0;
        ex.v[j] = bx.v[j] - p1;
        ey.v[j] = by.v[j] - p2;
    }
}

// ---- phase: w2 numerators 1/(|e|+reg) and their sum (Backend.cpp:351-366) -------------------
template <bool SHARD>
__device__ void phase_weights(const PoissonArgs &a, bool first, float reg, double &sum)
{
    for (int tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
        const TileIter t = tile_thread<SHARD>(a, tile);
        if (!t.valid) continue;
        F4 v0, vx, vy;
        if (first) {                                   // Solver.cpp:391-392: w2 = 1
#pragma unroll
            for (int j = 0; j < 4; j++) v0.v[j] = vx.v[j] = vy.v[j] = (t.x0 + j < a.W) ? 1.f : 0.f;
        } else {
            F4 e0[3], ex[3], ey[3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) residual4(a, t, ch, e0[ch], ex[ch], ey[ch]);
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool in = t.x0 + j < a.W;
                const float l0 = sqrtf(e0[0].v[j] * e0[0].v[j] + e0[1].v[j] * e0[1].v[j] + e0[2].v[j] * e0[2].v[j]);
                const float lx = sqrtf(ex[0].v[j] * ex[0].v[j] + ex[1].v[j] * ex[1].v[j] + ex[2].v[j] * ex[2].v[j]);
                const float ly = sqrtf(ey[0].v[j] * ey[0].v[j] + ey[1].v[j] * ey[1].v[j] + ey[2].v[j] * ey[2].v[j]);
                v0.v[j] = in ? 1.0f / (l0 + reg) : 0.f;
                vx.v[j] = in ? 1.0f / (lx + reg) : 0.f;
                vy.v[j] = in ? 1.0f / (ly + reg) : 0.f;
                part += v0.v[j] + vx.v[j] + vy.v[j];
            }
            sum += (double)part;
        }
        st4(a.plane[V0] + t.idx, v0);
        st4(a.plane[VX] + t.idx, vx);
        st4(a.plane[VY] + t.idx, vy);
    }
    // sharded: the weight of the vertical gradient in the row above the band (phase_rhs / phase A read it as "wyu"), from the
    // x halo rows -- same expression, not part of this GPU's share of the sum
    for_halo_rows<SHARD>(a, true, false, [&](const TileIter &t) {
        F4 vy;
        if (first) {
#pragma unroll
            for (int j = 0; j < 4; j++) vy.v[j] = (t.x0 + j < a.W) ? 1.f : 0.f;
        } else {
            F4 e0[3], ex[3], ey[3];
#pragma unroll
            for (int ch = 0; ch < 3; ch++) residual4(a, t, ch, e0[ch], ex[ch], ey[ch]);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float ly = sqrtf(ey[0].v[j] * ey[0].v[j] + ey[1].v[j] * ey[1].v[j] + ey[2].v[j] * ey[2].v[j]);
                vy.v[j] = (t.x0 + j < a.W) ? 1.0f / (ly + reg) : 0.f;
            }
        }
        st4(a.plane[VY] + t.idx, vy);
    });
}

// ---- phase: w2 = coef*v, r = P' W2 (b - Px), rz = r.r  (Backend.cpp:368-372, :190-217) ------
template <bool SHARD>
__device__ void phase_rhs(const PoissonArgs &a, float coef, double rz[3])
{
    float acc[3] = {0.f, 0.f, 0.f};
    for (int tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
        const TileIter t = tile_thread<SHARD>(a, tile);
        if (!t.valid) continue;
        const bool hasL = t.x0 > 0, hasU = t.y > 0;
        F4 w0 = ld4(a.plane[V0] + t.idx), wx = ld4(a.plane[VX] + t.idx), wy = ld4(a.plane[VY] + t.idx);
        F4 wyu = hasU ? ld4(a.plane[VY] + t.idx - a.Wp) : zero4();
        float wxl = hasL ? a.plane[VX][t.idx - 1] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) { w0.v[j] *= coef; wx.v[j] *= coef; wy.v[j] *= coef; wyu.v[j] *= coef; }
        wxl *= coef;
        st4(a.plane[W0] + t.idx, w0);
        st4(a.plane[WX] + t.idx, wx);
        st4(a.plane[WY] + t.idx, wy);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float *x = a.plane[X + ch];
            F4 e0, ex, ey;
            residual4(a, t, ch, e0, ex, ey);
            const F4 xc = ld4(x + t.idx);
            // ex at the left neighbour and ey at the upper neighbour, recomputed.
            float exl = 0.f;
            if (hasL) exl = a.plane[BX + ch][t.idx - 1] - (xc.v[0] - x[t.idx - 1]);
            F4 eyu = zero4();
            if (hasU) {
                const F4 xu = ld4(x + t.idx - a.Wp), byu = ld4(a.plane[BY + ch] + t.idx - a.Wp);
#pragma unroll
                for (int j = 0; j < 4; j++) eyu.v[j] = byu.v[j] - (xc.v[j] - xu.v[j]);
            }
            F4 r;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int xx = t.x0 + j;
                float v = w0.v[j] * e0.v[j] * a.alpha;
                if (xx != 0)        v += (j == 0 ? wxl : wx.v[j - 1]) * (j == 0 ? exl : ex.v[j - 1]);
                if (xx != a.W - 1)  v -= wx.v[j] * ex.v[j];
                if (t.y != 0)       v += wyu.v[j] * eyu.v[j];
                if (t.y != a.H - 1) v -= wy.v[j] * ey.v[j];
                if (xx >= a.W) v = 0.f;
                r.v[j] = v;
                acc[ch] += v * v;
            }
            st4(a.plane[R + ch] + t.idx, r);
            push_row<SHARD>(a, t, R + ch, r);
        }
    }
    for_halo_rows<SHARD>(a, true, false, [&](const TileIter &t) {
        F4 wyu = ld4(a.plane[VY] + t.idx);
#pragma unroll
        for (int j = 0; j < 4; j++) wyu.v[j] *= coef;
        st4(a.plane[WY] + t.idx, wyu);
    });
    rz[0] = acc[0]; rz[1] = acc[1]; rz[2] = acc[2];
}

// Resident x (which = 0) / Ap (which = 1) of this thread's 4 pixels in the CTA's k-th tile; T = tiles the variant holds.
template <int T> __device__ __forceinline__ F4 res_ld(const float4 *res, int k, int which, int ch)
{
    const float4 t = res[((which * T + k) * 3 + ch) * kThreads + threadIdx.x];
    return F4{{t.x, t.y, t.z, t.w}};
}
template <int T> __device__ __forceinline__ void res_st(float4 *res, int k, int which, int ch, const F4 &v)
{
    res[((which * T + k) * 3 + ch) * kThreads + threadIdx.x] = make_float4(v.v[0], v.v[1], v.v[2], v.v[3]);
}

// ---- phase A: p = r + b*p_old; x += a_prev*p_old; Ap = A p; pAp ---------------------------
// (Backend.cpp:325-347 of the previous iteration fused with :221-252 of this one.)
template <bool SHARD>
__device__ void phase_cg_a(const PoissonArgs &a, int pOld, int pNew, const float aPrev[3],
                           const float beta[3], double pAp[3])
{
    const float alphaSqr = a.alpha * a.alpha;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
        const TileIter t = tile_thread<SHARD>(a, tile);
        if (!t.valid) continue;
        const bool hasL = t.x0 > 0, hasR = t.x0 + 4 < a.W, hasU = t.y > 0, hasD = t.y < a.H - 1;
        const F4 w0 = ld4(a.plane[W0] + t.idx), wx = ld4(a.plane[WX] + t.idx), wy = ld4(a.plane[WY] + t.idx);
        const F4 wyu = hasU ? ld4(a.plane[WY] + t.idx - a.Wp) : zero4();
        const float wxl = hasL ? a.plane[WX][t.idx - 1] : 0.f;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float *r = a.plane[R + ch], *po = a.plane[pOld + ch];
            const float b = beta[ch], al = aPrev[ch];
            const F4 rc = ld4(r + t.idx), pc = ld4(po + t.idx);
            F4 xv = ld4(a.plane[X + ch] + t.idx);
            F4 c, u = zero4(), d = zero4();
            float l = 0.f, rr = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                xv.v[j] += pc.v[j] * al;
                c.v[j] = rc.v[j] + pc.v[j] * b;
            }
            if (hasU) {
                const F4 r2 = ld4(r + t.idx - a.Wp), p2 = ld4(po + t.idx - a.Wp);
#pragma unroll
                for (int j = 0; j < 4; j++) u.v[j] = r2.v[j] + p2.v[j] * b;
            }
            if (hasD) {
                const F4 r2 = ld4(r + t.idx + a.Wp), p2 = ld4(po + t.idx + a.Wp);
#pragma unroll
                for (int j = 0; j < 4; j++) d.v[j] = r2.v[j] + p2.v[j] * b;
            }
            if (hasL) l = r[t.idx - 1] + po[t.idx - 1] * b;
            if (hasR) rr = r[t.idx + 4] + po[t.idx + 4] * b;
            F4 Ap;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int xx = t.x0 + j;
                const float xi = c.v[j];
                float v = w0.v[j] * xi * alphaSqr;
                if (xx != 0)        v += (j == 0 ? wxl : wx.v[j - 1]) * (xi - (j == 0 ? l : c.v[j - 1]));
                if (xx != a.W - 1)  v += wx.v[j] * (xi - (j == 3 ? rr : c.v[j + 1]));
                if (t.y != 0)       v += wyu.v[j] * (xi - u.v[j]);
                if (t.y != a.H - 1) v += wy.v[j] * (xi - d.v[j]);
                if (xx >= a.W) v = 0.f;
                Ap.v[j] = v;
                acc[ch] += xi * v;
            }
            st4(a.plane[X + ch] + t.idx, xv);
            st4(a.plane[pNew + ch] + t.idx, c);
            st4(a.plane[AP + ch] + t.idx, Ap);
            push_row<SHARD>(a, t, pNew + ch, c);
        }
    }
    pAp[0] = acc[0]; pAp[1] = acc[1]; pAp[2] = acc[2];
}

// Phase A of the resident variant.  x and Ap live in shared memory (res); the new search direction c = r + b*p_old is handed
// to the neighbouring threads through a shared tile (one channel at a time: 18 x 72 floats) instead of being recomputed from
// r and p_old re-read through L1 -- with 48 KB of the CTA's shared memory taken, L1 no longer holds the tile's rows (measured:
// hit rate 25 % instead of 48 %, more L2 traffic than the streaming variant saves).  Only the tile's outer ring is recomputed
// from global r / p_old.  Same expressions on the same values as phase_cg_a => same bits.
constexpr int kHaloPitch = kTileGX * 4 + 8;     // 4 floats of margin left and right keep the float4 rows 16-byte aligned
template <int MODE, bool SHARD>
__device__ void phase_cg_a_res(const PoissonArgs &a, int pOld, int pNew, const float aPrev[3],
                               const float beta[3], double pAp[3], float4 *res, bool xResident)
{
    constexpr int T = res_tiles(MODE);
    __shared__ __align__(16) float s_c[kTileY + 2][kHaloPitch];
    const float alphaSqr = a.alpha * a.alpha;
    const int lx = threadIdx.x % kTileGX, ly = threadIdx.x / kTileGX;
    float acc[3] = {0.f, 0.f, 0.f};
    int k = -1;
    for (int tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
        k++;
        const TileIter t = tile_thread<SHARD>(a, tile);
        const bool hasL = t.x0 > 0, hasR = t.x0 + 4 < a.W, hasU = t.y > 0, hasD = t.y < a.H - 1;
        F4 w0 = zero4(), wx = zero4(), wy = zero4(), wyu = zero4();
        float wxl = 0.f;
        if (t.valid) {
            w0 = ld4(a.plane[W0] + t.idx); wx = ld4(a.plane[WX] + t.idx); wy = ld4(a.plane[WY] + t.idx);
            if (hasU) wyu = ld4(a.plane[WY] + t.idx - a.Wp);
            if (hasL) wxl = a.plane[WX][t.idx - 1];
        }
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float *r = a.plane[R + ch], *po = a.plane[pOld + ch];
            const float b = beta[ch], al = aPrev[ch];
            F4 c = zero4();
            if (t.valid) {
                const F4 rc = ld4(r + t.idx), pc = ld4(po + t.idx);
                F4 xv = (res_x(MODE) && xResident) ? res_ld<T>(res, k, 0, ch) : ld4(a.plane[X + ch] + t.idx);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    xv.v[j] += pc.v[j] * al;
                    c.v[j] = rc.v[j] + pc.v[j] * b;
                }
                if (res_x(MODE)) res_st<T>(res, k, 0, ch, xv);
                else st4(a.plane[X + ch] + t.idx, xv);
                st4(a.plane[pNew + ch] + t.idx, c);
                push_row<SHARD>(a, t, pNew + ch, c);
                *reinterpret_cast<float4 *>(&s_c[ly + 1][4 + 4 * lx]) = make_float4(c.v[0], c.v[1], c.v[2], c.v[3]);
                // the ring around the tile, from the neighbouring tiles' r and p_old
                if (ly == 0 && hasU) {
                    const F4 r2 = ld4(r + t.idx - a.Wp), p2 = ld4(po + t.idx - a.Wp);
                    *reinterpret_cast<float4 *>(&s_c[0][4 + 4 * lx]) =
                        make_float4(r2.v[0] + p2.v[0] * b, r2.v[1] + p2.v[1] * b, r2.v[2] + p2.v[2] * b, r2.v[3] + p2.v[3] * b);
                }
                if (ly == kTileY - 1 && hasD) {
                    const F4 r2 = ld4(r + t.idx + a.Wp), p2 = ld4(po + t.idx + a.Wp);
                    *reinterpret_cast<float4 *>(&s_c[kTileY + 1][4 + 4 * lx]) =
                        make_float4(r2.v[0] + p2.v[0] * b, r2.v[1] + p2.v[1] * b, r2.v[2] + p2.v[2] * b, r2.v[3] + p2.v[3] * b);
                }
                if (lx == 0 && hasL) s_c[ly + 1][3] = r[t.idx - 1] + po[t.idx - 1] * b;
                if (lx == kTileGX - 1 && hasR) s_c[ly + 1][4 + 4 * kTileGX] = r[t.idx + 4] + po[t.idx + 4] * b;
            }
            __syncthreads();
            if (t.valid) {
                F4 u = zero4(), d = zero4();
                float l = 0.f, rr = 0.f;
                if (hasU) { const float4 q = *reinterpret_cast<const float4 *>(&s_c[ly][4 + 4 * lx]); u = F4{{q.x, q.y, q.z, q.w}}; }
                if (hasD) { const float4 q = *reinterpret_cast<const float4 *>(&s_c[ly + 2][4 + 4 * lx]); d = F4{{q.x, q.y, q.z, q.w}}; }
                if (hasL) l = s_c[ly + 1][3 + 4 * lx];
                if (hasR) rr = s_c[ly + 1][8 + 4 * lx];
                F4 Ap;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int xx = t.x0 + j;
                    const float xi = c.v[j];
                    float v = w0.v[j] * xi * alphaSqr;
                    if (xx != 0)        v += (j == 0 ? wxl : wx.v[j - 1]) * (xi - (j == 0 ? l : c.v[j - 1]));
                    if (xx != a.W - 1)  v += wx.v[j] * (xi - (j == 3 ? rr : c.v[j + 1]));
                    if (t.y != 0)       v += wyu.v[j] * (xi - u.v[j]);
                    if (t.y != a.H - 1) v += wy.v[j] * (xi - d.v[j]);
                    if (xx >= a.W) v = 0.f;
                    Ap.v[j] = v;
                    acc[ch] += xi * v;
                }
                if (res_ap(MODE)) res_st<T>(res, k, 1, ch, Ap);
                else st4(a.plane[AP + ch] + t.idx, Ap);
            }
            __syncthreads();
        }
    }
    pAp[0] = acc[0]; pAp[1] = acc[1]; pAp[2] = acc[2];
}

// ---- phase B: r -= a*Ap; rz = r.r   (Backend.cpp:296-321) ----------------------------------
template <int MODE, bool SHARD>
__device__ void phase_cg_b(const PoissonArgs &a, const float al[3], double rz[3], const float4 *res)
{
    float acc[3] = {0.f, 0.f, 0.f};
    int k = -1;
    for (int tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
        k++;
        const TileIter t = tile_thread<SHARD>(a, tile);
        if (!t.valid) continue;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            F4 r = ld4(a.plane[R + ch] + t.idx);
            const F4 Ap = res_ap(MODE) ? res_ld<res_tiles(MODE)>(res, k, 1, ch) : ld4(a.plane[AP + ch] + t.idx);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float ri = r.v[j] - Ap.v[j] * al[ch];
                r.v[j] = ri;
                acc[ch] += ri * ri;
            }
            st4(a.plane[R + ch] + t.idx, r);
            push_row<SHARD>(a, t, R + ch, r);
        }
    }
    rz[0] = acc[0]; rz[1] = acc[1]; rz[2] = acc[2];
}

// ---- phase: pending x += a*p of the last CG iteration --------------------------------------
template <int MODE, bool SHARD>
__device__ void phase_flush_x(const PoissonArgs &a, int pCur, const float al[3], const float4 *res, bool xResident)
{
    int k = -1;
    for (int tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
        k++;
        const TileIter t = tile_thread<SHARD>(a, tile);
        if (!t.valid) continue;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            F4 x = (res_x(MODE) && xResident) ? res_ld<res_tiles(MODE)>(res, k, 0, ch) : ld4(a.plane[X + ch] + t.idx);
            const F4 p = ld4(a.plane[pCur + ch] + t.idx);
#pragma unroll
            for (int j = 0; j < 4; j++) x.v[j] += p.v[j] * al[ch];
            st4(a.plane[X + ch] + t.idx, x);
            push_row<SHARD>(a, t, X + ch, x);
        }
    }
}

// ---- phase: final = 1*direct + x  (Solver.cpp:561-567), with the last x update folded in ---
template <int MODE, bool SHARD>
__device__ void phase_export(const PoissonArgs &a, int pCur, const float al[3], const float4 *res, bool xResident)
{
    int k = -1;
    for (int tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
        k++;
        const TileIter t = tile_thread<SHARD>(a, tile);
        if (!t.valid) continue;
        F4 x[3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            x[ch] = (res_x(MODE) && xResident) ? res_ld<res_tiles(MODE)>(res, k, 0, ch) : ld4(a.plane[X + ch] + t.idx);
            const F4 p = ld4(a.plane[pCur + ch] + t.idx);
#pragma unroll
            for (int j = 0; j < 4; j++) x[ch].v[j] += p.v[j] * al[ch];
            st4(a.plane[X + ch] + t.idx, x[ch]);                  // the solved x stays in the plan (gdb200_poisson_metrics_device)
        }
        if (a.in_direct) {
            F4 d[3];
            load_rgb4(a.in_direct, a, t, d[0], d[1], d[2]);
#pragma unroll
            for (int ch = 0; ch < 3; ch++)
#pragma unroll
                for (int j = 0; j < 4; j++) x[ch].v[j] = 1.0f * d[ch].v[j] + x[ch].v[j];
        }
        store_rgb4(a.out_final, a, t, x[0], x[1], x[2]);
    }
}

// The CG scalars are uniform over the grid: thread 0 of each CTA derives them from the grid sums (every CTA the same bits) and
// keeps them in shared memory -- 16 registers per thread less than carrying them through the phases (the kernel is capped at
// 64 registers for 4 CTAs/SM and was spilling in the CG loop: 80 -> 156 bytes of spills cost 3 ms of 26 at 1024x1024).
struct CgScalars {
    float rz[2][3];      // r.r of the last two CG iterations; [cur] is the newer one (Solver.cpp:466 swaps pointers)
    float aPrev[3];      // step of the CG iteration whose x update is pending
    float beta[3], al[3];
    float coef;
};

template <int MODE, bool SHARD>
__global__ void __launch_bounds__(kThreads, 4) poisson_irls_cg_kernel(const PoissonArgs a)
{
#ifdef GDB200_EMU
    float4 *s_res = emu_dynamic_shared();
#else
    extern __shared__ float4 s_res[];     // variants 1, 2: [x | Ap][tile][channel][thread], kResBytes
#endif
    __shared__ CgScalars sc;
    bool xResident = false;               // the current x is in s_res, not in the X planes (uniform over the grid)
    cg::grid_group grid = cg::this_grid();
    SyncState sync;
    int cgTotal = 0, irlsDone = 0;
    const bool lead = threadIdx.x == 0;

    phase_import<SHARD>(a);
    if (lead) sc.aPrev[0] = sc.aPrev[1] = sc.aPrev[2] = 0.f;
    grid.sync();

    int pCur = PA;                       // plane set holding the current search direction
    int cur = 0;                         // sc.rz[cur]: r.r of the newest residual

    for (int irls = 0; irls < a.cfg.irlsIterMax; irls++) {
        if (irls > 0) {                                   // apply the pending x update first
            phase_flush_x<MODE, SHARD>(a, pCur, sc.aPrev, s_res, xResident);
            xResident = false;
            if (SHARD) {                                      // a reduction of nothing: publishes the pushed x halo rows
                float none[3];
                grid_sum3<SHARD>(grid, a, sync, 0.0, 0.0, 0.0, none);
                __syncthreads();
            } else grid.sync();
        }
        double part[3];
        float tot[3];
        if (irls == 0) {
            double dummy = 0.0;
            phase_weights<SHARD>(a, true, 0.f, dummy);
            if (lead) sc.coef = 1.0f;
            grid.sync();
        } else {
            const float reg = a.cfg.irlsRegInit * powf(a.cfg.irlsRegIter, (float)(irls - 1));   // Solver.cpp:395
            double s = 0.0;
            phase_weights<SHARD>(a, false, reg, s);
            grid_sum3<SHARD>(grid, a, sync, s, 0.0, 0.0, tot);
            if (lead) sc.coef = (float)(3 * a.W * a.H) / tot[0];      // (float)w2->numElems, Backend.cpp:368
            __syncthreads();
        }
        phase_rhs<SHARD>(a, sc.coef, part);
        grid_sum3<SHARD>(grid, a, sync, part[0], part[1], part[2], tot);
        if (lead) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                sc.rz[cur][c] = tot[c];
                sc.aPrev[c] = 0.f;
                sc.beta[c] = 0.f;                         // first direction: p = r (Solver.cpp:405)
            }
        }
        __syncthreads();

        for (int cgi = 0;; cgi++) {
            if (cgi % a.cfg.cgIterCheck == 0 || cgi == a.cfg.cgIterMax) {   // Solver.cpp:411-445
                const float errL2W = sc.rz[cur][0] + sc.rz[cur][1] + sc.rz[cur][2];
                if (cgi == a.cfg.cgIterMax || errL2W <= a.cfg.cgTolerance) break;
            }
            cur ^= 1;                                                       // Solver.cpp:466: rz <-> rz2
            const int pNew = (pCur == PA) ? PB : PA;
            if (MODE != 0) phase_cg_a_res<MODE, SHARD>(a, pCur, pNew, sc.aPrev, sc.beta, part, s_res, xResident);
            else phase_cg_a<SHARD>(a, pCur, pNew, sc.aPrev, sc.beta, part);
            xResident = res_x(MODE);
            grid_sum3<SHARD>(grid, a, sync, part[0], part[1], part[2], tot);
            pCur = pNew;
            if (lead) {
#pragma unroll
                for (int c = 0; c < 3; c++) sc.al[c] = sc.rz[cur ^ 1][c] / fmaxf(tot[c], FLT_MIN);   // Backend.cpp:309
            }
            __syncthreads();
            phase_cg_b<MODE, SHARD>(a, sc.al, part, s_res);
            grid_sum3<SHARD>(grid, a, sync, part[0], part[1], part[2], tot);
            if (lead) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    sc.rz[cur][c] = tot[c];
                    sc.beta[c] = tot[c] / fmaxf(sc.rz[cur ^ 1][c], FLT_MIN);                         // Backend.cpp:339
                    sc.aPrev[c] = sc.al[c];
                }
            }
            __syncthreads();
            cgTotal++;
        }
        irlsDone++;
    }
    phase_export<MODE, SHARD>(a, pCur, sc.aPrev, s_res, xResident);
    if (blockIdx.x == 0 && lead) {
        a.iters[0] = irlsDone; a.iters[1] = cgTotal;
        if (SHARD) a.status[1] = a.epochBase + sync.epoch;
    }
}

// ---- Solver::evaluateMetricsMTS (Solver.cpp:511-541) on the x a solve left in the plan: e = b - P x; the primal block of e
// goes out as an image, sum |e_i| and sum |e_i|^2 over the 3n RGB elements of e as two doubles (the host divides by 3n).
__global__ void __launch_bounds__(kThreads) poisson_metrics_kernel(const PoissonArgs a, float *err, double *sums)
{
    double s1 = 0.0, s2 = 0.0;
    for (int tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
        const TileIter t = tile_thread(a, tile);
        if (!t.valid) continue;
        F4 e0[3], ex[3], ey[3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) residual4(a, t, ch, e0[ch], ex[ch], ey[ch]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (t.x0 + j >= a.W) continue;
            const F4 *blk[3] = {e0, ex, ey};
#pragma unroll
            for (int b = 0; b < 3; b++) {
                const float r = blk[b][0].v[j], g = blk[b][1].v[j], bl = blk[b][2].v[j];
                const float l2 = r * r + g * g + bl * bl;
                s1 += (double)sqrtf(l2); s2 += (double)l2;
            }
        }
        store_rgb4(err, a, t, e0[0], e0[1], e0[2]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sums[0], s1); atomicAdd(&sums[1], s2); }
}

}  // namespace gdb200

#ifndef GDB200_EMU
// =================================================================================== host ====

struct gdb200_poisson_plan {
    int device = 0, w = 0, h = 0, wp = 0, grid = 0;
    float *planes = nullptr;
    double *red = nullptr;
    int *iters = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evHost[4] = {nullptr, nullptr, nullptr, nullptr};   // copy timing of the host-pointer entry point (created on first use)
    gdb200::PoissonArgs last;          // geometry, alpha and planes of the last solve (gdb200_poisson_metrics_device)
    bool solved = false;
    int variant = 0;                   // kernel variant the solves of this plan run (poisson_irls_cg_kernel<variant>)
    long long nTiles = 0;
    int occ = 0, sms = 0;
    // sharded solve: this plan covers rows [y0, y1) of the image as rank `rank` of `nRanks`
    int y0 = 0, y1 = 0, rank = 0, nRanks = 1;
    size_t planeElems = 0;             // floats per plane: (y1 - y0 + 2) rows of wp
    gdb200::Mail *mail = nullptr;      // [2][kMaxRanks] written by the peers, then [2][kMaxRanks] of local relay slots
    unsigned *status = nullptr;
    unsigned epochBase = 16;           // message number the next solve starts after (mailboxes start zeroed)
    struct Peer { float *planes = nullptr; gdb200::Mail *mail = nullptr; int y0 = 0, y1 = 0; bool opened[2] = {false, false}; };
    Peer peer[gdb200::kMaxRanks];
    // staging for the host-pointer entry point
    float *d_in[4] = {nullptr, nullptr, nullptr, nullptr};
    float *d_out = nullptr;
};

using namespace gdb200;

static void *variant_kernel(int v, bool shard)
{
    switch (v) {
    case 1: return shard ? (void *)poisson_irls_cg_kernel<1, true> : (void *)poisson_irls_cg_kernel<1, false>;
    case 2: return shard ? (void *)poisson_irls_cg_kernel<2, true> : (void *)poisson_irls_cg_kernel<2, false>;
    case 3: return shard ? (void *)poisson_irls_cg_kernel<3, true> : (void *)poisson_irls_cg_kernel<3, false>;
    default: return shard ? (void *)poisson_irls_cg_kernel<0, true> : (void *)poisson_irls_cg_kernel<0, false>;
    }
}
static size_t variant_smem(int v) { return res_tiles(v) ? kResBytes : 0; }

// A variant fits when it keeps the streaming variant's residency (the grid was sized for it) and every CTA's tiles fit.
static bool variant_fits(const gdb200_poisson_plan *p, int v)
{
    if (v == 0) return true;
    if (variant_smem(v) && cudaFuncSetAttribute(variant_kernel(v, p->nRanks > 1), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)variant_smem(v)) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, variant_kernel(v, p->nRanks > 1), kThreads, variant_smem(v)) != cudaSuccess) { cudaGetLastError(); return false; }
    if ((long long)occ * p->sms < p->grid) return false;
    return res_tiles(v) == 0 || p->nTiles <= (long long)res_tiles(v) * p->grid;
}

extern "C" {

int gdb200_poisson_preset(const char *preset, gdb200_poisson_config *c)
{
    if (!preset || !c) return set_error(GDB200_ERR_ARGUMENT, "preset/out_cfg is NULL");
    // Base config, Solver.cpp:92-100.
    c->irlsIterMax = 1; c->irlsRegInit = 0.f; c->irlsRegIter = 0.f;
    c->cgIterMax = 1; c->cgIterCheck = 100; c->cgTolerance = 0.f;
    if (!strcmp(preset, "L1D")) { c->irlsIterMax = 20; c->irlsRegInit = 0.05f; c->irlsRegIter = 0.5f; c->cgIterMax = 50; return GDB200_OK; }
    if (!strcmp(preset, "L1Q")) { c->irlsIterMax = 64; c->irlsRegInit = 1.0f; c->irlsRegIter = 0.7f; c->cgIterMax = 1000; return GDB200_OK; }
    if (!strcmp(preset, "L1L")) { c->irlsIterMax = 7; c->irlsRegInit = 1.0e-4f; c->irlsRegIter = 1.0e-1f; c->cgIterMax = 20000; c->cgTolerance = 1.0e-20f; return GDB200_OK; }
    if (!strcmp(preset, "L2D")) { c->cgIterMax = 50; return GDB200_OK; }
    if (!strcmp(preset, "L2Q")) { c->cgIterMax = 500; return GDB200_OK; }
    return set_error(GDB200_ERR_ARGUMENT, "unknown solver preset '%s' (expected L1D, L1Q, L1L, L2D or L2Q)", preset);
}

void gdb200_poisson_plan_destroy(gdb200_poisson_plan *p)
{
    if (!p) return;
    for (int r = 0; r < kMaxRanks; r++) {
        if (p->peer[r].opened[0]) cudaIpcCloseMemHandle(p->peer[r].planes);
        if (p->peer[r].opened[1]) cudaIpcCloseMemHandle(p->peer[r].mail);
    }
    cudaFree(p->planes); cudaFree(p->red); cudaFree(p->iters); cudaFree(p->mail); cudaFree(p->status);
    for (float *d : p->d_in) cudaFree(d);
    cudaFree(p->d_out);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    for (cudaEvent_t e : p->evHost) if (e) cudaEventDestroy(e);
    delete p;
}

int gdb200_poisson_plan_create(int w, int h, gdb200_poisson_plan **out)
{
    return gdb200_poisson_shard_create(w, h, 0, h, 0, 1, out);
}

int gdb200_poisson_shard_create(int w, int h, int y0, int y1, int rank, int n_ranks, gdb200_poisson_plan **out)
{
    if (!out) return set_error(GDB200_ERR_ARGUMENT, "out_plan is NULL");
    *out = nullptr;
    if (w <= 0 || h <= 0 || (long long)w * h > (1LL << 29))
        return set_error(GDB200_ERR_ARGUMENT, "invalid image size %dx%d", w, h);
    if (n_ranks < 1 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks)
        return set_error(GDB200_ERR_ARGUMENT, "rank %d of %d (at most %d GPUs share a solve)", rank, n_ranks, kMaxRanks);
    if (y0 < 0 || y1 > h || y0 >= y1)
        return set_error(GDB200_ERR_ARGUMENT, "row band [%d, %d) of a %dx%d image", y0, y1, w, h);
    if ((rank == 0) != (y0 == 0) || (rank == n_ranks - 1) != (y1 == h))
        return set_error(GDB200_ERR_ARGUMENT, "bands must cover the image in rank order: rank %d of %d has rows [%d, %d) of %d",
                         rank, n_ranks, y0, y1, h);
    DeviceInfo di;
    if (int rc = device_info(&di)) return rc;
    gdb200_poisson_plan *p = new gdb200_poisson_plan;
    p->device = di.device; p->w = w; p->h = h; p->wp = (w + 3) & ~3;
    p->y0 = y0; p->y1 = y1; p->rank = rank; p->nRanks = n_ranks;
    const size_t planeElems = (size_t)p->wp * (y1 - y0 + 2);          // + one halo row above and below
    p->planeElems = planeElems;
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, variant_kernel(0, n_ranks > 1), kThreads, 0);
    if (e != cudaSuccess || occ < 1) {
        delete p;
        return set_error(GDB200_ERR_CUDA, "poisson kernel not launchable on this device: %s "
                         "(library is built for sm_100a only)", cudaGetErrorString(e));
    }
    const int tilesX = (p->wp / 4 + kTileGX - 1) / kTileGX, tilesY = (y1 - y0 + kTileY - 1) / kTileY;
    const long long nTiles = (long long)tilesX * tilesY;
    p->grid = (int)std::min<long long>(nTiles, (long long)occ * di.sms);
    p->nTiles = nTiles;
    p->occ = occ; p->sms = di.sms;
    // the fastest variant that fits: x and Ap resident, else x resident, else streaming
    p->variant = 0;
    for (int v = 1; v <= 2 && p->variant == 0; v++)
        if (variant_fits(p, v)) p->variant = v;
#define PLAN_CUDA(call) do { cudaError_t e2 = (call); if (e2 != cudaSuccess) { gdb200_poisson_plan_destroy(p); \
        return set_error(GDB200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e2)); } } while (0)
    PLAN_CUDA(cudaMalloc(&p->planes, planeElems * kPlanes * sizeof(float)));
    PLAN_CUDA(cudaMemset(p->planes, 0, planeElems * kPlanes * sizeof(float)));
    PLAN_CUDA(cudaMalloc(&p->red, sizeof(double) * 2 * 3 * p->grid));
    PLAN_CUDA(cudaMalloc(&p->iters, sizeof(int) * 2));
    PLAN_CUDA(cudaMalloc(&p->mail, sizeof(Mail) * 4 * kMaxRanks));
    PLAN_CUDA(cudaMemset(p->mail, 0, sizeof(Mail) * 4 * kMaxRanks));
    PLAN_CUDA(cudaMalloc(&p->status, sizeof(unsigned) * 2));
    PLAN_CUDA(cudaMemset(p->status, 0, sizeof(unsigned) * 2));
    p->peer[rank].planes = p->planes; p->peer[rank].mail = p->mail; p->peer[rank].y0 = y0; p->peer[rank].y1 = y1;
    PLAN_CUDA(cudaEventCreate(&p->ev0));
    PLAN_CUDA(cudaEventCreate(&p->ev1));
#undef PLAN_CUDA
    *out = p;
    return GDB200_OK;
}

int gdb200_poisson_solve_device(gdb200_poisson_plan *p, const float *d_dx, const float *d_dy,
                                const float *d_thr, const float *d_direct, float alpha,
                                const gdb200_poisson_config *cfg, float *d_out, void *stream,
                                gdb200_stats *stats)
{
    if (!p || !d_dx || !d_dy || !cfg || !d_out)
        return set_error(GDB200_ERR_ARGUMENT, "plan, dx, dy, cfg and out_final are required");
    // the plan's workspace lives on p->device: run there whatever the caller's current device is, and give it back
    struct Bind { int prev = -1; ~Bind() { if (prev >= 0) cudaSetDevice(prev); } } bind;
    if (cudaGetDevice(&bind.prev) != cudaSuccess) { bind.prev = -1; cudaGetLastError(); }
    if (bind.prev != p->device) GDB_CUDA(cudaSetDevice(p->device)); else bind.prev = -1;
    PoissonArgs a;
    memset(&a, 0, sizeof(a));
    a.W = p->w; a.H = p->h; a.Wp = p->wp; a.Gx = p->wp / 4;
    a.tilesX = (a.Gx + kTileGX - 1) / kTileGX;
    a.nTiles = a.tilesX * ((p->y1 - p->y0 + kTileY - 1) / kTileY);
    a.y0 = p->y0; a.y1 = p->y1; a.rank = p->rank; a.nRanks = p->nRanks;
    if (p->nRanks > 1) {
        for (int r = 0; r < p->nRanks; r++) {
            if (!p->peer[r].mail) return set_error(GDB200_ERR_ARGUMENT, "sharded solve: rank %d is not connected (gdb200_poisson_shard_connect)", r);
            a.mail[r] = p->peer[r].mail;
        }
        if (p->rank > 0) {
            const gdb200_poisson_plan::Peer &up = p->peer[p->rank - 1];
            if (up.y1 != p->y0) return set_error(GDB200_ERR_ARGUMENT, "sharded solve: rank %d ends at row %d, rank %d starts at %d", p->rank - 1, up.y1, p->rank, p->y0);
            a.peerUp = up.planes; a.peerUpRows = up.y1 - up.y0; a.peerUpElems = (size_t)p->wp * (up.y1 - up.y0 + 2);
        }
        if (p->rank < p->nRanks - 1) {
            const gdb200_poisson_plan::Peer &dn = p->peer[p->rank + 1];
            if (dn.y0 != p->y1) return set_error(GDB200_ERR_ARGUMENT, "sharded solve: rank %d ends at row %d, rank %d starts at %d", p->rank, p->y1, p->rank + 1, dn.y0);
            a.peerDown = dn.planes; a.peerDownElems = (size_t)p->wp * (dn.y1 - dn.y0 + 2);
        }
    }
    a.epochBase = p->epochBase;
    a.status = p->status;
    a.aosVec = (p->w % 4 == 0) &&
               ((((uintptr_t)d_dx | (uintptr_t)d_dy | (uintptr_t)d_thr | (uintptr_t)d_direct | (uintptr_t)d_out) & 15) == 0);
    // Params::sanitize (Solver.cpp:168-178) and m_P.alpha (Solver.cpp:319).
    a.alpha = d_thr ? fmaxf(alpha, 0.f) : 0.f;
    a.cfg = *cfg;
    a.cfg.irlsIterMax = std::max(cfg->irlsIterMax, 1);
    a.cfg.irlsRegInit = fmaxf(cfg->irlsRegInit, 0.f);
    a.cfg.irlsRegIter = fmaxf(cfg->irlsRegIter, 0.f);
    a.cfg.cgIterMax = std::max(cfg->cgIterMax, 1);
    a.cfg.cgIterCheck = std::max(cfg->cgIterCheck, 1);
    a.cfg.cgTolerance = fmaxf(cfg->cgTolerance, 0.f);
    const size_t planeElems = p->planeElems;
    for (int i = 0; i < kPlanes; i++) a.plane[i] = p->planes + planeElems * i;
    a.in_dx = d_dx; a.in_dy = d_dy; a.in_thr = d_thr; a.in_direct = d_direct; a.out_final = d_out;
    a.red = p->red; a.iters = p->iters;

    cudaStream_t s = (cudaStream_t)stream;
    p->last = a; p->solved = true;
    if (stats) GDB_CUDA(cudaEventRecord(p->ev0, s));
    void *kargs[] = {&a};
    GDB_CUDA(cudaLaunchCooperativeKernel(variant_kernel(p->variant, p->nRanks > 1), dim3(p->grid), dim3(kThreads), kargs, variant_smem(p->variant), s));
    if (stats || p->nRanks > 1) {
        GDB_CUDA(cudaEventRecord(p->ev1, s));
        GDB_CUDA(cudaEventSynchronize(p->ev1));
    }
    if (p->nRanks > 1) {
        unsigned st2[2] = {0, 0};
        GDB_CUDA(cudaMemcpy(st2, p->status, sizeof(st2), cudaMemcpyDeviceToHost));
        const unsigned st = st2[0];
        p->epochBase = st2[1];
        if (st) {
            GDB_CUDA(cudaMemset(p->status, 0, sizeof(unsigned)));
            return set_error(GDB200_ERR_CUDA, "sharded solve: rank %d never saw the reduction message of rank %u (peer not solving, or peer memory not reachable)",
                             p->rank, st - 1);
        }
    }
    if (stats) {
        float ms = 0.f;
        GDB_CUDA(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
        int it[2];
        GDB_CUDA(cudaMemcpy(it, p->iters, sizeof(it), cudaMemcpyDeviceToHost));
        stats->device_ms = ms; stats->launches = 1; stats->irls_iters = it[0]; stats->cg_iters = it[1];
    }
    return GDB200_OK;
}

/* ---- sharded solve: wiring the GPUs together ------------------------------------------------------------------------------ */
int gdb200_poisson_shard_export(gdb200_poisson_plan *p, void *out_handle)
{
    if (!p || !out_handle) return set_error(GDB200_ERR_ARGUMENT, "plan and out_handle are required");
    gdb200_shard_handle h;
    memset(&h, 0, sizeof(h));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle layout");
    cudaIpcMemHandle_t m;
    GDB_CUDA(cudaIpcGetMemHandle(&m, p->planes)); memcpy(h.planes, &m, 64);
    GDB_CUDA(cudaIpcGetMemHandle(&m, p->mail));   memcpy(h.mail, &m, 64);
    h.y0 = p->y0; h.y1 = p->y1; h.rank = p->rank; h.w = p->w; h.h = p->h;
    h.planes_ptr = (unsigned long long)(uintptr_t)p->planes; h.mail_ptr = (unsigned long long)(uintptr_t)p->mail;
    h.device = p->device; h.pid = (long long)getpid();
    memcpy(out_handle, &h, sizeof(h));
    return GDB200_OK;
}

int gdb200_poisson_shard_connect(gdb200_poisson_plan *p, const void *handles, int n)
{
    if (!p || !handles) return set_error(GDB200_ERR_ARGUMENT, "plan and handles are required");
    if (n != p->nRanks) return set_error(GDB200_ERR_ARGUMENT, "%d handles for a solve shared by %d ranks", n, p->nRanks);
    struct Bind { int prev = -1; ~Bind() { if (prev >= 0) cudaSetDevice(prev); } } bind;
    if (cudaGetDevice(&bind.prev) != cudaSuccess) { bind.prev = -1; cudaGetLastError(); }
    if (bind.prev != p->device) GDB_CUDA(cudaSetDevice(p->device)); else bind.prev = -1;
    const gdb200_shard_handle *hs = (const gdb200_shard_handle *)handles;
    for (int r = 0; r < n; r++) {
        const gdb200_shard_handle &h = hs[r];
        if (h.rank != r || h.w != p->w || h.h != p->h)
            return set_error(GDB200_ERR_ARGUMENT, "handle %d describes rank %d of a %dx%d solve (expected rank %d, %dx%d)", r, h.rank, h.w, h.h, r, p->w, p->h);
        if (r == p->rank) continue;
        gdb200_poisson_plan::Peer &peer = p->peer[r];
        if (peer.opened[0]) { cudaIpcCloseMemHandle(peer.planes); peer.opened[0] = false; }      // connecting again replaces the mappings
        if (peer.opened[1]) { cudaIpcCloseMemHandle(peer.mail); peer.opened[1] = false; }
        peer.planes = nullptr; peer.mail = nullptr;
        peer.y0 = h.y0; peer.y1 = h.y1;
        if (h.pid == (long long)getpid()) {
            // same process (one host thread per GPU): plain peer access
            if (h.device != p->device) {
                int can = 0;
                GDB_CUDA(cudaDeviceCanAccessPeer(&can, p->device, h.device));
                if (!can) return set_error(GDB200_ERR_CUDA, "device %d cannot access the memory of device %d", p->device, h.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return set_error(GDB200_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", h.device, cudaGetErrorString(e));
                cudaGetLastError();
            }
            peer.planes = (float *)(uintptr_t)h.planes_ptr; peer.mail = (Mail *)(uintptr_t)h.mail_ptr;
        } else {
            cudaIpcMemHandle_t m;
            void *ptr = nullptr;
            memcpy(&m, h.planes, 64);
            GDB_CUDA(cudaIpcOpenMemHandle(&ptr, m, cudaIpcMemLazyEnablePeerAccess));
            peer.planes = (float *)ptr; peer.opened[0] = true;
            memcpy(&m, h.mail, 64);
            GDB_CUDA(cudaIpcOpenMemHandle(&ptr, m, cudaIpcMemLazyEnablePeerAccess));
            peer.mail = (Mail *)ptr; peer.opened[1] = true;
        }
    }
    return GDB200_OK;
}

int gdb200_poisson_plan_set_variant(gdb200_poisson_plan *p, int variant)
{
    if (!p) return set_error(GDB200_ERR_ARGUMENT, "plan is NULL");
    if (variant < 0 || variant > 3) return set_error(GDB200_ERR_ARGUMENT, "solver kernel variant %d (expected 0..3)", variant);
    struct Bind { int prev = -1; ~Bind() { if (prev >= 0) cudaSetDevice(prev); } } bind;      // function attributes are per device
    if (cudaGetDevice(&bind.prev) != cudaSuccess) { bind.prev = -1; cudaGetLastError(); }
    if (bind.prev != p->device) GDB_CUDA(cudaSetDevice(p->device)); else bind.prev = -1;
    if (!variant_fits(p, variant))
        return set_error(GDB200_ERR_ARGUMENT, "image of %dx%d does not fit solver kernel variant %d (%lld tiles, %d CTAs x %d resident tiles)",
                         p->w, p->h, variant, p->nTiles, p->grid, res_tiles(variant));
    p->variant = variant;
    return GDB200_OK;
}

int gdb200_poisson_plan_variant(const gdb200_poisson_plan *p) { return p ? p->variant : -1; }

static thread_local gdb200_poisson_plan *g_cachedPlan = nullptr;     // plan of this thread's host-pointer solves

int gdb200_poisson_metrics_device(gdb200_poisson_plan *p, float *d_err, float *out_errL1, float *out_errL2, void *stream);

/* Host-pointer form: Solver::evaluateMetricsMTS after the gdb200_poisson_solve this thread made last. */
int gdb200_poisson_metrics(float *err, float *out_errL1, float *out_errL2)
{
    gdb200_poisson_plan *p = g_cachedPlan;
    if (!p || !p->solved || !p->d_out) return set_error(GDB200_ERR_ARGUMENT, "evaluateMetrics needs a gdb200_poisson_solve on this thread first");
    if (!err) return set_error(GDB200_ERR_ARGUMENT, "err is NULL");
    if (int rc = gdb200_poisson_metrics_device(p, p->d_out, out_errL1, out_errL2, nullptr)) return rc;    // d_out: staging, already copied out
    GDB_CUDA(cudaMemcpy(err, p->d_out, (size_t)p->w * p->h * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    return GDB200_OK;
}

int gdb200_poisson_metrics_device(gdb200_poisson_plan *p, float *d_err, float *out_errL1, float *out_errL2, void *stream)
{
    if (!p || !d_err || !out_errL1 || !out_errL2) return set_error(GDB200_ERR_ARGUMENT, "plan, err and the two outputs are required");
    if (!p->solved) return set_error(GDB200_ERR_ARGUMENT, "evaluateMetrics needs a solve on this plan first");
    struct Bind { int prev = -1; ~Bind() { if (prev >= 0) cudaSetDevice(prev); } } bind;
    if (cudaGetDevice(&bind.prev) != cudaSuccess) { bind.prev = -1; cudaGetLastError(); }
    if (bind.prev != p->device) GDB_CUDA(cudaSetDevice(p->device)); else bind.prev = -1;
    cudaStream_t s = (cudaStream_t)stream;
    PoissonArgs a = p->last;
    a.aosVec = (p->w % 4 == 0) && (((uintptr_t)d_err & 15) == 0);
    double *sums = p->red;                       // the solve's reduction scratch is free between solves
    GDB_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2, s));
    poisson_metrics_kernel<<<p->grid, kThreads, 0, s>>>(a, d_err, sums);
    GDB_CUDA(cudaGetLastError());
    double h[2];
    GDB_CUDA(cudaMemcpyAsync(h, sums, sizeof(h), cudaMemcpyDeviceToHost, s));
    GDB_CUDA(cudaStreamSynchronize(s));
    const double n3 = 3.0 * (double)p->w * (double)p->h;
    *out_errL1 = (float)(h[0] / n3); *out_errL2 = (float)(h[1] / n3);
    return GDB200_OK;
}

int gdb200_poisson_solve(const float *dx, const float *dy, const float *throughput,
                         const float *direct, int w, int h, float alpha, const char *preset,
                         float *out_final, gdb200_stats *stats)
{
    if (!dx || !dy || !out_final) return set_error(GDB200_ERR_ARGUMENT, "dx, dy and out_final are required");
    gdb200_poisson_config cfg;
    if (int rc = gdb200_poisson_preset(preset, &cfg)) return rc;
    // One cached plan per thread: Mitsuba calls this once per render, benches call it in a loop.
    gdb200_poisson_plan *&cached = g_cachedPlan;
    int dev = -1;
    if (int rc = require_device()) return rc;
    GDB_CUDA(cudaGetDevice(&dev));
    if (cached && (cached->w != w || cached->h != h || cached->device != dev)) {
        gdb200_poisson_plan_destroy(cached);
        cached = nullptr;
    }
    if (!cached) if (int rc = gdb200_poisson_plan_create(w, h, &cached)) return rc;
    gdb200_poisson_plan *p = cached;
    const size_t bytes = (size_t)w * h * 3 * sizeof(float);
    const float *src[4] = {dx, dy, throughput, direct};
    for (int i = 0; i < 4; i++)
        if (src[i] && !p->d_in[i]) GDB_CUDA(cudaMalloc(&p->d_in[i], bytes));
    if (!p->d_out) GDB_CUDA(cudaMalloc(&p->d_out, bytes));

    cudaEvent_t *e = p->evHost;
    for (int i = 0; i < 4; i++) if (!e[i]) GDB_CUDA(cudaEventCreate(&e[i]));
    cudaStream_t s = 0;
    GDB_CUDA(cudaEventRecord(e[0], s));
    for (int i = 0; i < 4; i++)
        if (src[i]) GDB_CUDA(cudaMemcpyAsync(p->d_in[i], src[i], bytes, cudaMemcpyHostToDevice, s));
    GDB_CUDA(cudaEventRecord(e[1], s));
    gdb200_stats local;
    memset(&local, 0, sizeof(local));
    int rc = gdb200_poisson_solve_device(p, p->d_in[0], p->d_in[1], throughput ? p->d_in[2] : nullptr,
                                         direct ? p->d_in[3] : nullptr, alpha, &cfg, p->d_out, s, &local);
    if (rc) return rc;
    GDB_CUDA(cudaEventRecord(e[2], s));
    GDB_CUDA(cudaMemcpyAsync(out_final, p->d_out, bytes, cudaMemcpyDeviceToHost, s));
    GDB_CUDA(cudaEventRecord(e[3], s));
    GDB_CUDA(cudaEventSynchronize(e[3]));
    if (stats) {
        float a = 0.f, b = 0.f;
        GDB_CUDA(cudaEventElapsedTime(&a, e[0], e[1]));
        GDB_CUDA(cudaEventElapsedTime(&b, e[2], e[3]));
        *stats = local;
        stats->h2d_ms = a; stats->d2h_ms = b;
    }
    return GDB200_OK;
}

}  // extern "C"
#endif  // GDB200_EMU
