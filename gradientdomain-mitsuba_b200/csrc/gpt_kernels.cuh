// gdb200 G-PT wavefront tracer for sm_100a (fp64): state layout, film, per-slot device routines and kernels.
// (Included by gpt.cu; tests/emu/ compiles the per-slot routines for the host through a shim to check them
// against the oracle without a GPU — test infrastructure only, the product never runs them on the CPU.)
#pragma once
#include "gpt_device.cuh"

namespace gdb200 {


// ------------------------------------------------------------------ state layout
// fp64 state is stored as 32-byte records: one vector (and one scalar riding in its 4th lane) per record, one DRAM
// sector each.  Layout [slot][kRecPitch records]: everything of a slot sits in one 2 KB block.  (Round 1 used
// [record][slot]: perfect coalescing while a kernel walks consecutive slots, but the wavefront's queues hold every
// third or tenth slot, so each 32-byte sector came from a different DRAM row -- every kernel then saturates at the
// ~2.5 TB/s that HBM3e delivers for row-per-sector access, with half of each 64-byte fetch granule belonging to a slot
// that is not in the queue (profiles/r02_stage_kernels_ncu.txt).  With the slot's records adjacent, the records a
// thread reads back to back share lines, fetch granules and DRAM rows however scattered the queue is.)
// Records of the base path and of one offset path.  Both groups are padded to 12 records = three 128-byte lines, so a group
// starts on a line boundary and records that are used together (THR|RAD, GRAD|P, ...) share DRAM's 64-byte access granule.
// The *_X records belong to the staged wavefront (gpt_stages.cuh).
enum BaseRec { BR_RAYD = 0 /* w: path pdf */, BR_P /* w: eta */, BR_GN /* w: sample x (fused kernels) */, BR_S /* w: sample y (fused kernels) */,
               BR_T, BR_N, BR_WI, BR_THR, BR_RAD, BR_VD, BR_X0, BR_X1, BR_COUNT };
enum OffRec { OR_THR = 0 /* w: path pdf */, OR_RAD, OR_GRAD, OR_P, OR_GN, OR_S, OR_T, OR_N, OR_WI, OR_X0, OR_X1, OR_X2, OR_COUNT };
static_assert(BR_COUNT == 12 && OR_COUNT == 12, "record groups are whole cache lines");
constexpr int kRecords = BR_COUNT + 4 * OR_COUNT;   // 60 records
constexpr int kRecPitchLog2 = 6, kRecPitch = 1 << kRecPitchLog2;   // records per slot block incl. the staged wavefront's (gpt_stages.cuh), padded to 2 KB
enum IntField { IF_STATUS = 0, IF_MAT, IF_EMI, IF_DEPTH, IF_SAMPLE, IF_RNGN, IF_OFLAGS, IF_STREAM, IF_OMAT0, IF_OMAT1, IF_OMAT2, IF_OMAT3,
                IF_BSTYPE, IF_PEND,      // staged wavefront (gpt_stages.cuh): sampled BSDF component, bookkeeping of the offsets awaiting a ray
                IF_COUNT };
// ST_WAIT_*: staged wavefront only -- the slot has rays in flight and continues in the named stage once they are cast
enum SlotStatus { ST_FRESH = 0, ST_LIVE = 1, ST_FINISHED = 2, ST_DONE = 3, ST_WAIT_PRIMARY = 4, ST_WAIT_SHADE = 5, ST_WAIT_RESOLVE = 6 };
enum { RAY_NOT_CONNECTED = 0, RAY_RECENTLY_CONNECTED = 1, RAY_CONNECTED = 2 };
enum { BUF_FINAL = 0, BUF_THROUGHPUT = 1, BUF_DX = 2, BUF_DY = 3, BUF_DIRECT = 4 };

constexpr int kBounceThreads = 128, kGenThreads = 128;
constexpr int kBsdfTypes = GDB200_BSDF_ROUGHDIELECTRIC + 1;
constexpr int kBuckets = 3 * kBsdfTypes;  // one queue per (BSDF type of the base vertex) x (shift stage of the offset paths)

struct GptArgs {
    double *sd;            // [nSlots][kRecPitch][4]
    int *si;               // [nSlots][16]
    int nSlots, width, height, yBegin;
    int nStreams, nPixels, streamsPerPixel, pad0;   // sample streams (pixel x chunk) are dealt to the slots: stream = chunk * nPixels + pixel
    int bandRows, bandCount, bandIndex, pad1;   // interleaved row bands (bandCount > 1) instead of one strip
    int spp, skipPreview;   // skipPreview: the "-final" preview puts (gpt.cpp:1319-1324) are not needed when a reconstruction overwrites that buffer
    uint64_t seed;
    Config cfg;
    double *film;          // [5][H][W][4]
    int *liveList;         // [2][kBuckets][nSlots]
    int *liveCount;        // [2][kBuckets]
    int *genList;          // [2][nSlots]: slots whose path ended (to splat + regenerate)
    int *genCount;         // [2]
    unsigned long long *counters;   // [0] done slots, [1] rays, [2] path vertices, [3] samples, [4] state bytes, [5] path bounces, [6] next stream, [7] ray-queue overflows
    // staged wavefront (gpt_stages.cuh): ray queues + results, stage queues
    double *rays[2];       // [0] nearest-hit rays, [1] any-hit (shadow) rays: [rayCapacity][8] = o.xyz d.xyz mint maxt
    int *rayOwner[2];      // [rayCapacity]: slot * 8 + ray id of the slot
    int *rayCount;         // [2]
    int rayCapacity, pad2;
    int *qKey;             // [nSlots]: the stage queue a slot belongs in next (-1: none), written by the stage that leaves it there
    int *qList;            // [kStageBuckets][nSlots]: slots per stage bucket
    int *qCount;           // [kStageBuckets]
    unsigned long long *mark;   // != NULL: the kernel stores %globaltimer here when it starts (per-phase times without host events)
};

// Phase timing of the staged wavefront: the first thread of a kernel that opens a phase stamps the GPU's nanosecond timer.
// Kernels of one stream run back to back, so the difference of two stamps is the phase between them -- read back once per
// render (80 KB) instead of ~10 000 cudaEventElapsedTime calls (10 ms of every render, as much as 3 % of a strip at 8 GPUs).
GDB_D void stampPhase(const GptArgs &a)
{
#ifndef GDB200_EMU
    if (a.mark && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        *a.mark = t;
    }
#endif
}

GDB_D double *REC(const GptArgs &a, int rec, int slot) { return a.sd + ((((size_t)slot << kRecPitchLog2) + rec) << 2); }
GDB_D double &W(const GptArgs &a, int rec, int slot) { return REC(a, rec, slot)[3]; }
// int fields of a slot share one 64-byte line [slot][16] (fields 0-7 in its first sector), so a kernel pulls one
// sector per slot instead of one per field; the per-pixel sampler key is recomputed, not stored.
GDB_D int &SI(const GptArgs &a, int field, int slot) { return a.si[((size_t)slot << 4) + field]; }
// Fire-and-forget fp64 accumulation into global memory: one RED.E.ADD.F64 (plain atomicAdd on a generic pointer
// compiles to an address-space dispatch with CAS loops for the shared/local cases).
GDB_D void redAdd(double *p, double v)
{
#ifdef GDB200_EMU
    atomicAdd(p, v);
#else
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "d"(v) : "memory");
#endif
}
GDB_D void redAdd(unsigned long long *p, unsigned long long v)
{
#ifdef GDB200_EMU
    atomicAdd(p, v);
#else
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
#endif
}
GDB_D unsigned long long atomAdd(unsigned long long *p, unsigned long long v)
{
#ifdef GDB200_EMU
    return atomicAdd(p, v);
#else
    unsigned long long o;
    asm volatile("atom.global.add.u64 %0, [%1], %2;" : "=l"(o) : "l"(__cvta_generic_to_global(p)), "l"(v) : "memory");
    return o;
#endif
}
GDB_D int atomAdd(int *p, int v)
{
#ifdef GDB200_EMU
    return atomicAdd(p, v);
#else
    int o;
    asm volatile("atom.global.add.s32 %0, [%1], %2;" : "=r"(o) : "l"(__cvta_generic_to_global(p)), "r"(v) : "memory");
    return o;
#endif
}
GDB_D void prefetchL2(const void *p)
{
#ifndef GDB200_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
GDB_D V3 ldv(const GptArgs &a, int rec, int slot)
{
    const double2 *p = reinterpret_cast<const double2 *>(REC(a, rec, slot));
    const double2 lo = p[0], hi = p[1];
    return mk(lo.x, lo.y, hi.x);
}
// Read of a record that was updated with redAdd (performed at L2): bypass L1, which may hold the line from before
// the reductions when the same thread reads it back inside one launch (gpt_tail_kernel).
GDB_D V3 ldvL2(const GptArgs &a, int rec, int slot)
{
#ifdef GDB200_EMU
    return ldv(a, rec, slot);
#else
    const double2 *p = reinterpret_cast<const double2 *>(REC(a, rec, slot));
    const double2 lo = __ldcg(p), hi = __ldcg(p + 1);
    return mk(lo.x, lo.y, hi.x);
#endif
}
GDB_D void ldvw(const GptArgs &a, int rec, int slot, V3 &v, Float &w)
{
    const double2 *p = reinterpret_cast<const double2 *>(REC(a, rec, slot));
    const double2 lo = p[0], hi = p[1];
    v = mk(lo.x, lo.y, hi.x); w = hi.y;
}
GDB_D void stv(const GptArgs &a, int rec, int slot, V3 v)
{
    double *p = REC(a, rec, slot);
    *reinterpret_cast<double2 *>(p) = make_double2(v.x, v.y);
    p[2] = v.z;
}
// Full-sector store (4th lane zeroed): a 24-byte store makes L2 fetch the sector from DRAM first to merge it, a 32-byte one
// does not -- the staged wavefront writes every record whole (the fused kernels keep scalars of their own in some 4th lanes).
GDB_D void stvf(const GptArgs &a, int rec, int slot, V3 v)
{
    double2 *p = reinterpret_cast<double2 *>(REC(a, rec, slot));
    p[0] = make_double2(v.x, v.y); p[1] = make_double2(v.z, 0.0);
}
GDB_D void stvw(const GptArgs &a, int rec, int slot, V3 v, Float w)
{
    double2 *p = reinterpret_cast<double2 *>(REC(a, rec, slot));
    p[0] = make_double2(v.x, v.y); p[1] = make_double2(v.z, w);
}

GDB_D void storeBaseIts(const GptArgs &a, int slot, const Its &its)
{
    stv(a, BR_P, slot, its.p); stv(a, BR_GN, slot, its.geoN); stv(a, BR_S, slot, its.sh.s); stv(a, BR_T, slot, its.sh.t);
    stv(a, BR_N, slot, its.sh.n); stv(a, BR_WI, slot, its.wi);
    SI(a, IF_MAT, slot) = its.material; SI(a, IF_EMI, slot) = its.emitter;
}
GDB_D void loadBaseIts(const GptArgs &a, int slot, Its &its)
{
    its.t = 0; its.p = ldv(a, BR_P, slot); its.geoN = ldv(a, BR_GN, slot); its.sh.s = ldv(a, BR_S, slot); its.sh.t = ldv(a, BR_T, slot);
    its.sh.n = ldv(a, BR_N, slot); its.wi = ldv(a, BR_WI, slot);
    its.material = SI(a, IF_MAT, slot); its.emitter = SI(a, IF_EMI, slot);
}
// the same with whole-sector stores; eta rides in BR_P's 4th lane
GDB_D void storeBaseItsFull(const GptArgs &a, int slot, const Its &its, Float eta)
{
    stvw(a, BR_P, slot, its.p, eta); stvf(a, BR_GN, slot, its.geoN); stvf(a, BR_S, slot, its.sh.s); stvf(a, BR_T, slot, its.sh.t);
    stvf(a, BR_N, slot, its.sh.n); stvf(a, BR_WI, slot, its.wi);
    SI(a, IF_MAT, slot) = its.material; SI(a, IF_EMI, slot) = its.emitter;
}
GDB_D void storeOffItsFull(const GptArgs &a, int slot, int i, const Its &its)
{
    const int o = BR_COUNT + i * OR_COUNT;
    stvf(a, o + OR_P, slot, its.p); stvf(a, o + OR_GN, slot, its.geoN); stvf(a, o + OR_S, slot, its.sh.s); stvf(a, o + OR_T, slot, its.sh.t);
    stvf(a, o + OR_N, slot, its.sh.n); stvf(a, o + OR_WI, slot, its.wi);
    SI(a, IF_OMAT0 + i, slot) = its.material;
}
GDB_D void storeOffIts(const GptArgs &a, int slot, int i, const Its &its)
{
    const int o = BR_COUNT + i * OR_COUNT;
    stv(a, o + OR_P, slot, its.p); stv(a, o + OR_GN, slot, its.geoN); stv(a, o + OR_S, slot, its.sh.s); stv(a, o + OR_T, slot, its.sh.t);
    stv(a, o + OR_N, slot, its.sh.n); stv(a, o + OR_WI, slot, its.wi);
    SI(a, IF_OMAT0 + i, slot) = its.material;
}
GDB_D void loadOffIts(const GptArgs &a, int slot, int i, Its &its)
{
    const int o = BR_COUNT + i * OR_COUNT;
    its.t = 0; its.p = ldv(a, o + OR_P, slot); its.geoN = ldv(a, o + OR_GN, slot); its.sh.s = ldv(a, o + OR_S, slot); its.sh.t = ldv(a, o + OR_T, slot);
    its.sh.n = ldv(a, o + OR_N, slot); its.wi = ldv(a, o + OR_WI, slot);
    its.material = SI(a, IF_OMAT0 + i, slot); its.emitter = -1;
}

// shifted.addRadiance / addGradient (gpt.cpp:147-156).  Most bounces add exact zeros (light sample occluded, no
// emitter hit); x + 0 == x bit for bit, so those skip the read-modify-write of the two accumulator records.
// The accumulators are only ever added to, so the update is a fire-and-forget reduction (RED.ADD.F64: same x + d
// arithmetic as load-add-store, but no load whose latency this thread would have to wait out).
GDB_D void accumulateOffset(const GptArgs &a, int o, int slot, Spec dRad, Spec dGrad)
{
#ifdef GDB_RMW_ACCUM
    if (!(dRad.x == 0 && dRad.y == 0 && dRad.z == 0)) stv(a, o + OR_RAD, slot, ldv(a, o + OR_RAD, slot) + dRad);
    if (!(dGrad.x == 0 && dGrad.y == 0 && dGrad.z == 0)) stv(a, o + OR_GRAD, slot, ldv(a, o + OR_GRAD, slot) + dGrad);
#else
    if (!(dRad.x == 0 && dRad.y == 0 && dRad.z == 0)) { double *p = REC(a, o + OR_RAD, slot); redAdd(p, dRad.x); redAdd(p + 1, dRad.y); redAdd(p + 2, dRad.z); }
    if (!(dGrad.x == 0 && dGrad.y == 0 && dGrad.z == 0)) { double *p = REC(a, o + OR_GRAD, slot); redAdd(p, dGrad.x); redAdd(p + 1, dGrad.y); redAdd(p + 2, dGrad.z); }
#endif
}

// Warp-aggregated append of an ended slot to the next step's regeneration queue.
GDB_D void appendGen(const GptArgs &a, int parity, int slot)
{
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomAdd(&a.genCount[parity], __popc(m));
    base = __shfl_sync(m, base, leader);
    a.genList[(size_t)parity * a.nSlots + base + __popc(m & ((1u << lane) - 1))] = slot;
}

// Image row of an owned pixel: one contiguous strip, or interleaved bands of bandRows rows dealt round-robin to bandCount ranks.
GDB_D int pixelRow(const GptArgs &a, int pixel)
{
    const int lr = pixel / a.width;
    if (a.bandCount <= 1) return a.yBegin + lr;
    return ((lr / a.bandRows) * a.bandCount + a.bandIndex) * a.bandRows + lr % a.bandRows;
}
// A sample stream = (pixel, chunk): chunk c of a pixel holds spp/C samples (+1 for c < spp%C) and draws from the pixel's
// gdb200_counter stream (c == 0) or from an independently re-keyed one (gdb200_gpt_params.streams_per_pixel).
struct StreamInfo { int px, py, chunk, count; uint64_t key; };
GDB_D StreamInfo streamInfo(const GptArgs &a, int stream)
{
    StreamInfo s;
    s.chunk = stream / a.nPixels;
    const int pixel = stream - s.chunk * a.nPixels;
    s.px = pixel % a.width; s.py = pixelRow(a, pixel);
    s.count = a.spp / a.streamsPerPixel + (s.chunk < a.spp % a.streamsPerPixel ? 1 : 0);
    s.key = samplerKey(a.seed, s.px, s.py);
    if (s.chunk > 0) s.key = mix64(s.key ^ ((uint64_t)s.chunk * 0xD1B54A32D192ED03ULL));
    return s;
}

// One atomic per warp for statistics counters.
GDB_D void countWarp(unsigned long long *ctr, unsigned v)
{
    const unsigned m = __activemask();
    const unsigned s = __reduce_add_sync(m, v);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1 && s) redAdd(ctr, (unsigned long long)s);
}

// ------------------------------------------------------------------ film (ImageBlock::put, imageblock.h:150-195)
GDB_D Float evalDiscretized(Float x)      // rfilter.h:76-77, MTS_FILTER_RESOLUTION = 31
{
    const int idx = min((int)fabs(x * c_scene.filterScale), 31);
    if (c_scene.filterIsBox) return idx < 31 ? c_scene.filterTable[0] : 0.0;
    return c_sceneG->filterTable[idx];                                               // per-lane index: global-memory copy
}
GDB_CALL void filmPut(const GptArgs &a, Float sx, Float sy, Spec v, Float weight, int buf, bool allowNegative)
{
    const Float value[4] = {v.x, v.y, v.z, weight};
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (!isfinite(value[i]) || (!allowNegative && value[i] < 0)) return;       // dropped together with its weight
    const Float radius = c_scene.filterRadius, posx = sx - 0.5, posy = sy - 0.5;
    const int W = a.width, H = a.height;
    const int minx = max((int)ceil(posx - radius), 0), miny = max((int)ceil(posy - radius), 0);
    const int maxx = min((int)floor(posx + radius), W - 1), maxy = min((int)floor(posy + radius), H - 1);
    for (int y = miny; y <= maxy; ++y) {
        const Float weightY = evalDiscretized(y - posy);
        for (int x = minx; x <= maxx; ++x) {
            const Float wgt = evalDiscretized(x - posx) * weightY;
            double *dst = a.film + ((((size_t)buf * H + y) * W + x) << 2);
#pragma unroll
            for (int k = 0; k < 4; k++) redAdd(dst + k, wgt * value[k]);
        }
    }
}

// The 15 puts of renderBlock (gpt.cpp:1319-1352).
GDB_D void splatSample(const GptArgs &a, Float spx, Float spy, Spec veryDirect, Spec C, const Spec rad[4], const Spec grad[4])
{
    const int RIGHT = 0, BOTTOM = 1, LEFT = 2, TOP = 3;
    if (!a.skipPreview) {
        filmPut(a, spx, spy, (8 * veryDirect) + (2 * C), 4.0, BUF_FINAL, false);
        filmPut(a, spx - 1, spy, 2 * rad[LEFT], 1.0, BUF_FINAL, false);
        filmPut(a, spx + 1, spy, 2 * rad[RIGHT], 1.0, BUF_FINAL, false);
        filmPut(a, spx, spy - 1, 2 * rad[TOP], 1.0, BUF_FINAL, false);
        filmPut(a, spx, spy + 1, 2 * rad[BOTTOM], 1.0, BUF_FINAL, false);
    }
    filmPut(a, spx, spy, 2 * C, 4.0, BUF_THROUGHPUT, false);
    filmPut(a, spx - 1, spy, 2 * rad[LEFT], 1.0, BUF_THROUGHPUT, false);
    filmPut(a, spx + 1, spy, 2 * rad[RIGHT], 1.0, BUF_THROUGHPUT, false);
    filmPut(a, spx, spy - 1, 2 * rad[TOP], 1.0, BUF_THROUGHPUT, false);
    filmPut(a, spx, spy + 1, 2 * rad[BOTTOM], 1.0, BUF_THROUGHPUT, false);
    filmPut(a, spx - 1, spy, -(2 * grad[LEFT]), 1.0, BUF_DX, true);
    filmPut(a, spx, spy, 2 * grad[RIGHT], 1.0, BUF_DX, true);
    filmPut(a, spx, spy - 1, -(2 * grad[TOP]), 1.0, BUF_DY, true);
    filmPut(a, spx, spy, 2 * grad[BOTTOM], 1.0, BUF_DY, true);
    filmPut(a, spx, spy, veryDirect, 1.0, BUF_DIRECT, false);
}

GDB_D unsigned packFlag(int i, bool alive, int conn) { return ((alive ? 1u : 0u) | ((unsigned)conn << 1)) << (3 * i); }
GDB_D bool flagAlive(unsigned f, int i) { return (f >> (3 * i)) & 1u; }
GDB_D int flagConn(unsigned f, int i) { return (f >> (3 * i + 1)) & 3u; }
GDB_D unsigned setFlag(unsigned f, int i, bool alive, int conn) { return (f & ~(7u << (3 * i))) | packFlag(i, alive, conn); }

// ------------------------------------------------------------------ generate: splat finished paths, start next samples
GDB_D void generateBody(const GptArgs &a, int slot)
{
    if (SI(a, IF_STATUS, slot) == ST_FINISHED) {
        Spec rad[4], grad[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { const int o = BR_COUNT + i * OR_COUNT; rad[i] = ldvL2(a, o + OR_RAD, slot); grad[i] = ldvL2(a, o + OR_GRAD, slot); }
        splatSample(a, W(a, BR_GN, slot), W(a, BR_S, slot), ldv(a, BR_VD, slot), ldv(a, BR_RAD, slot), rad, grad);
    }

    int stream = SI(a, IF_STREAM, slot);
    StreamInfo si = streamInfo(a, stream);
    Sampler smp; smp.key = si.key; smp.n = (uint32_t)SI(a, IF_RNGN, slot);          // Sampler::generate, gpt.cpp:1250-1251
    int j = SI(a, IF_SAMPLE, slot);
    unsigned long long rays = 0, samples = 0;
    int status = ST_DONE;
    for (;;) {
        if (j >= si.count) {                                                         // stream exhausted: take the next one
            stream = (int)atomAdd(&a.counters[6], 1ULL);
            if (stream >= a.nStreams) break;
            si = streamInfo(a, stream); smp.key = si.key; smp.n = 0; j = 0;
            continue;
        }
        j++; samples++;
        const Float u = smp.next1D(), v = smp.next1D();                              // gpt.cpp:1261
        const Float spx = si.px + u, spy = si.py + v;
        Float apx = 0.5, apy = 0.5;                                                  // gpt.cpp:1235
        if (c_scene.apertureRadius > 0) { apx = smp.next1D(); apy = smp.next1D(); }  // needsApertureSample, gpt.cpp:1263-1265
        Ray ray; Its mits;
        sampleCameraRay(spx, spy, apx, apy, ray);                                    // gpt.cpp:402
        const bool mainValid = rayIntersectByValue(ray, mits); rays += 5;                   // gpt.cpp:472
        Spec veryDirect = splat(0);
        unsigned flags = 0;
        bool early = !mainValid;                                                     // gpt.cpp:482-492
        if (!mainValid && c_scene.env.present) veryDirect = veryDirect + splat(1.0) * envEval(ray.d);   // gpt.cpp:486-488 (looked up without ray differentials)
        if (mainValid && mits.emitter >= 0) veryDirect = veryDirect + splat(1.0) * emittedLe(mits, -ray.d);   // gpt.cpp:497-499
        if (mainValid && a.cfg.strictNormals && dot(ray.d, mits.geoN) * mits.wi.z >= 0) early = true;          // gpt.cpp:518-521
        const Float shiftX[4] = {1, 0, -1, 0}, shiftY[4] = {0, 1, 0, -1};            // gpt.cpp:410-415
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            Ray sray; Its sits;
            sampleCameraRay(spx + shiftX[i], spy + shiftY[i], apx, apy, sray);       // gpt.cpp:418
            bool alive = rayIntersectByValue(sray, sits);                            // gpt.cpp:476-480, 508-513
            if (alive && a.cfg.strictNormals && dot(sray.d, sits.geoN) * sits.wi.z >= 0) alive = false;   // gpt.cpp:523-530
            flags |= packFlag(i, alive, RAY_NOT_CONNECTED);
            if (!early) {
                const int o = BR_COUNT + i * OR_COUNT;
                stvw(a, o + OR_THR, slot, splat(1.0), 1.0);
                stv(a, o + OR_RAD, slot, splat(0)); stv(a, o + OR_GRAD, slot, splat(0));
                if (alive) storeOffIts(a, slot, i, sits);
            }
        }
        if (early || !(1 < a.cfg.maxDepth || a.cfg.maxDepth < 0)) {                  // bounce loop never entered (gpt.cpp:537)
            const Spec zero[4] = {splat(0), splat(0), splat(0), splat(0)};
            splatSample(a, spx, spy, veryDirect, splat(0), zero, zero);
            if (!early) redAdd(&a.counters[2], 1ULL);                             // avgPathLength += depth (1), gpt.cpp:1178-1179
            continue;
        }
        storeBaseIts(a, slot, mits);
        stvw(a, BR_RAYD, slot, ray.d, 1.0); W(a, BR_P, slot) = 1.0;          // pdf = 1, eta = 1
        stv(a, BR_THR, slot, splat(1.0));
        stv(a, BR_RAD, slot, splat(0)); stv(a, BR_VD, slot, veryDirect);
        W(a, BR_GN, slot) = spx; W(a, BR_S, slot) = spy;
        SI(a, IF_DEPTH, slot) = 1; SI(a, IF_OFLAGS, slot) = (int)flags;
        status = ST_LIVE;
        break;
    }
    SI(a, IF_STATUS, slot) = status; SI(a, IF_SAMPLE, slot) = j; SI(a, IF_RNGN, slot) = (int)smp.n; SI(a, IF_STREAM, slot) = stream;
    countWarp(&a.counters[0], status == ST_DONE ? 1u : 0u);
    countWarp(&a.counters[1], (unsigned)rays);
    countWarp(&a.counters[3], (unsigned)samples);
}

#ifndef GDB_GEN_MINBLOCKS
#define GDB_GEN_MINBLOCKS 2         // resident CTAs/SM the generate kernel's register allocation is sized for (2 => up to 255 registers)
#endif
__global__ void __launch_bounds__(kGenThreads, GDB_GEN_MINBLOCKS) gpt_generate_kernel(const GptArgs a, int parity)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.genCount[parity]) return;
    generateBody(a, a.genList[(size_t)parity * a.nSlots + g]);
}

// ------------------------------------------------------------------ bounce: one iteration of gpt.cpp:537-1175
// The reference runs two loops over the offset paths per bounce (NEE, then BSDF-sample stage).
// Here the base path's NEE, BSDF sample and extension ray are computed first and ONE loop then
// performs both stages per offset path, so each offset's state crosses HBM once per bounce; the
// base path's radiance is still accumulated in the reference's order (all NEE terms, then all
// BSDF-stage terms).
// PHASE 0 = next-event estimation of the base path and of its four offset paths (gpt.cpp:565-730);
// PHASE 1 = BSDF sample, extension ray, shifts, Russian roulette (gpt.cpp:737-1175).  Two launches per
// step over the same queues: each phase's hot code fits the instruction cache and needs fewer registers.
// QUEUED: the step-synchronous wavefront (ended slots are appended to the regeneration queue);
// !QUEUED: the tail kernel, where a thread runs its slot to completion.
template <int PHASE, bool QUEUED>
GDB_D void bounceBody(const GptArgs &a, int slot, int parity)
{
    constexpr bool kNee = PHASE != 1, kBsdf = PHASE != 0;   // PHASE 2 runs both stages in one pass over the state
    if (PHASE == 1 && SI(a, IF_STATUS, slot) != ST_LIVE) return;      // ended in phase 0 (strictNormals)
    const Config cfg = a.cfg;

#ifdef GDB_PREFETCH
    // The offset paths' records are read one offset at a time further down (each read a full DRAM round trip on
    // the critical path of this thread): request their sectors now so those reads hit L2.
    {
        const unsigned f0 = (unsigned)SI(a, IF_OFLAGS, slot);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (!flagAlive(f0, i)) continue;
            const int o = BR_COUNT + i * OR_COUNT;
            prefetchL2(REC(a, o + OR_THR, slot)); prefetchL2(REC(a, o + OR_RAD, slot)); prefetchL2(REC(a, o + OR_GRAD, slot));
            if (flagConn(f0, i) != RAY_CONNECTED) prefetchL2(REC(a, o + OR_P, slot));
            if (flagConn(f0, i) == RAY_NOT_CONNECTED) {
                prefetchL2(REC(a, o + OR_GN, slot)); prefetchL2(REC(a, o + OR_S, slot)); prefetchL2(REC(a, o + OR_T, slot));
                prefetchL2(REC(a, o + OR_N, slot)); prefetchL2(REC(a, o + OR_WI, slot));
            }
        }
    }
#endif
    Its mits; loadBaseIts(a, slot, mits);
    V3 mrayD; Float mpdf;
    ldvw(a, BR_RAYD, slot, mrayD, mpdf);
    Spec mthr = ldv(a, BR_THR, slot), mrad = ldv(a, BR_RAD, slot);
    Float meta = W(a, BR_P, slot);
    int depth = SI(a, IF_DEPTH, slot);
    unsigned flags = (unsigned)SI(a, IF_OFLAGS, slot);
    Sampler smp; smp.key = streamInfo(a, SI(a, IF_STREAM, slot)).key; smp.n = (uint32_t)SI(a, IF_RNGN, slot);
    unsigned rays = 0;
    bool ended = false;
    if (kNee) {   // algorithmic state traffic of this path-bounce (record sizes of SURVEY.md §8d: base 320 B,
        // unconnected offset 304 B, connected offset 88 B; read + write)
        unsigned bytes = 320;
        for (int i = 0; i < 4; i++) if (flagAlive(flags, i)) bytes += flagConn(flags, i) == RAY_CONNECTED ? 88 : 304;
        countWarp(&a.counters[4], 2 * bytes);
        countWarp(&a.counters[5], 1u);
    }

    if (kNee && cfg.strictNormals) {                                           // gpt.cpp:541-555
        if (dot(mrayD, mits.geoN) * mits.wi.z >= 0) ended = true;
        else
            for (int i = 0; i < 4; i++) {       // an unconnected offset's ray direction is -toWorld(wi) of its stored vertex
                if (!flagAlive(flags, i) || flagConn(flags, i) != RAY_NOT_CONNECTED) continue;
                Its sits; loadOffIts(a, slot, i, sits);
                const V3 sd = -toWorld(sits.sh, sits.wi);
                if (dot(sd, sits.geoN) * sits.wi.z >= 0) flags = setFlag(flags, i, false, flagConn(flags, i));
            }
    }

    if (!ended) {
        const bool lastSegment = (depth + 1 == cfg.maxDepth);                        // gpt.cpp:558
        const DMaterial &mainBSDF = c_sceneG->materials[mits.material];
        const Frame prevSh = mits.sh; const V3 prevP = mits.p, prevWi = mits.wi;     // the vertex both stages shade (previousMainIts, gpt.cpp:753)

        // ---------------- base path: next event estimation, gpt.cpp:565-607
        bool neeActive = false, neeVisible = false, atPointLight = false;
        Float lsx = 0, lsy = 0, neeBsdfPdf = 0, neeDistSq = 0, neeOppCos = 0, neeWNum = 0, neeWDen = 0, neeLightPdf = 0;
        V3 neeWoLocal = mk(0, 0, 0), neeLightP = mk(0, 0, 0), neeLightN = mk(0, 0, 0);
        Spec neeBsdfValue = splat(0), neeEmitterRadiance = splat(0), neeContributionAll = splat(0);
        if (kNee && (mainBSDF.flags & ESmooth) && depth + 1 >= cfg.minDepth) {       // gpt.cpp:568
            DRec dRec; initDRec(mits, dRec);
            lsx = smp.next1D(); lsy = smp.next1D();                                  // gpt.cpp:572
            const Spec value = sampleEmitterDirectVisible(dRec, lsx, lsy, neeVisible); rays++;
            neeEmitterRadiance = value * dRec.pdf;                                   // gpt.cpp:575
            neeWoLocal = toLocal(mits.sh, dRec.d);
            bsdfEvalPdf(mainBSDF, mits.wi, neeWoLocal, ESolidAngle, neeBsdfValue, neeBsdfPdf);   // gpt.cpp:588
            atPointLight = c_sceneG->emitters[dRec.emitter].kind == EM_POINT || c_sceneG->emitters[dRec.emitter].kind == EM_SPOT;        // dRec.measure == EDiscrete; such an emitter is not "on a surface"
            if (!neeVisible || atPointLight) neeBsdfPdf = 0;                         // gpt.cpp:592
            neeDistSq = len2(mits.p - dRec.p);                                       // gpt.cpp:595-596
            neeOppCos = dot(dRec.n, (mits.p - dRec.p)) / sqrt(neeDistSq);
            neeWNum = mpdf * dRec.pdf;                                               // gpt.cpp:599-600
            neeWDen = (mpdf * mpdf) * ((dRec.pdf * dRec.pdf) + (neeBsdfPdf * neeBsdfPdf));
            neeLightP = dRec.p; neeLightN = dRec.n; neeLightPdf = dRec.pdf;
            neeActive = !cfg.strictNormals || dot(mits.geoN, dRec.d) * neeWoLocal.z > 0;   // gpt.cpp:607
            neeContributionAll = mthr * (neeBsdfValue * neeEmitterRadiance);
        }

        // ---------------- base path: BSDF sample + extension, gpt.cpp:737-820
        bool bsdfStage = false, mainHitEmitter = false, escaped = false;
        BSDFSample bs;
        bs.weight = splat(0); bs.pdf = 0; bs.eta = 1.0; bs.sampledType = 0; bs.wo = mk(0, 0, 0);
        bs.extraDraws = 0;
        if (kBsdf) {                                                                 // gpt.cpp:456-457
            const Float sx = smp.next1D(), sy = smp.next1D();
            const Float s3 = mainBSDF.type == GDB200_BSDF_ROUGHDIELECTRIC ? smp.peek1D() : 0.0;   // EUsesSampler: drawn inside BSDF::sample
            bsdfSample(mainBSDF, mits.wi, sx, sy, s3, bs);
            smp.n += (uint32_t)bs.extraDraws;
        }
        Spec mainEmitterRadiance = splat(0), mainContributionAll = splat(0);
        DRec mainDRec; initDRec(mits, mainDRec);                                     // gpt.cpp:759
        int mainVertexType = 0, mainNextVertexType = 0;
        Float mainLumPdf = 0, mainWeightNumerator = 0, mainWeightDenominator = 0;
        if (!kBsdf) { }
        else if (bs.pdf <= 0.0) ended = true;                                        // gpt.cpp:739
        else {
            const V3 mainWo = toWorld(mits.sh, bs.wo);
            if (cfg.strictNormals && dot(mits.geoN, mainWo) * bs.wo.z <= 0) ended = true;   // gpt.cpp:748
            else {
                mainVertexType = vertexType(mainBSDF, bs.sampledType);               // gpt.cpp:764
                Ray mray; mray.o = mits.p; mray.d = mainWo; mray.mint = kEpsilon; mray.maxt = CUDART_INF;   // gpt.cpp:767
                rays++;
                if (rayIntersect(mray, mits)) {
                    bsdfStage = true;
                    if (mits.emitter >= 0) {                                         // gpt.cpp:771-776
                        mainEmitterRadiance = emittedLe(mits, -mainWo);
                        mainDRec.p = mits.p; mainDRec.n = mits.sh.n; mainDRec.d = mainWo; mainDRec.dist = mits.t; mainDRec.emitter = mits.emitter;
                        mainHitEmitter = true;
                    }
                    mainNextVertexType = vertexType(c_sceneG->materials[mits.material], bs.sampledType);   // gpt.cpp:784
                } else if (c_scene.env.present) {                                    // gpt.cpp:786-799: the base path left the scene
                    mainEmitterRadiance = envEval(mainWo);
                    if (envFillDRec(mainDRec, mray.o, mainWo)) {
                        bsdfStage = true; escaped = true; mainHitEmitter = true;
                        mainNextVertexType = VERTEX_TYPE_DIFFUSE;                    // "environment connection as diffuse"
                    } else ended = true;
                } else ended = true;                                                 // gpt.cpp:800-803
                if (bsdfStage) {
                    mrayD = mainWo;
                    const Float mainPreviousPdf = mpdf;                              // gpt.cpp:807-812
                    mthr = mthr * (bs.weight * bs.pdf);
                    mpdf *= bs.pdf;
                    meta *= bs.eta;
                    mainLumPdf = (mainHitEmitter && depth + 1 >= cfg.minDepth && !(bs.sampledType & EDelta)) ? pdfEmitterDirect(mainDRec) : 0;   // gpt.cpp:815-816
                    mainWeightNumerator = mainPreviousPdf * bs.pdf;                  // gpt.cpp:819-820
                    mainWeightDenominator = (mainPreviousPdf * mainPreviousPdf) * ((mainLumPdf * mainLumPdf) + (bs.pdf * bs.pdf));
                    mainContributionAll = mthr * mainEmitterRadiance;
                }
            }
        }
        const Float mainBsdfPdf = bs.pdf;
        const bool addBsdfStage = bsdfStage && depth + 1 >= cfg.minDepth;            // gpt.cpp:1140

        // ---------------- the four offset paths: gpt.cpp:609-727 and 830-1151 in one pass
        Float bw0 = 0, bw1 = 0, bw2 = 0, bw3 = 0; unsigned bHas = 0;                 // BSDF-stage weights of the base contribution
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            const int o = BR_COUNT + i * OR_COUNT;
            bool alive = flagAlive(flags, i);
            int conn = flagConn(flags, i);
            Spec sthr = splat(0); Float spdf = 0;
            if (alive) ldvw(a, o + OR_THR, slot, sthr, spdf);
            Its sits;
            if (alive && conn == RAY_NOT_CONNECTED) loadOffIts(a, slot, i, sits);
            V3 recentWiL = mk(0, 0, 0);
            if (alive && conn == RAY_RECENTLY_CONNECTED) recentWiL = toLocal(prevSh, normalize(ldv(a, o + OR_P, slot) - prevP));   // gpt.cpp:640, 864

            if (kNee && neeActive) {                                           // ---- NEE stage, gpt.cpp:609-727
                Spec mainContribution = splat(0), shiftedContribution = splat(0);
                Float weight = 0;
                bool shiftSuccessful = alive;
                if (shiftSuccessful) {
                    if (conn == RAY_CONNECTED) {                                     // gpt.cpp:622-637
                        const Float jacobian = 1;
                        const Float den = (jacobian * spdf) * (jacobian * spdf) * ((neeLightPdf * neeLightPdf) + (neeBsdfPdf * neeBsdfPdf));
                        weight = neeWNum / (kDEps + den + neeWDen);
                        mainContribution = neeContributionAll;
                        shiftedContribution = jacobian * sthr * (neeBsdfValue * neeEmitterRadiance);
                    } else if (conn == RAY_RECENTLY_CONNECTED) {                     // gpt.cpp:638-658
                        Spec shiftedBsdfValue; Float shiftedBsdfPdf;
                        bsdfEvalPdf(mainBSDF, recentWiL, neeWoLocal, ESolidAngle, shiftedBsdfValue, shiftedBsdfPdf);
                        if (!neeVisible || atPointLight) shiftedBsdfPdf = 0;
                        const Float jacobian = 1;
                        const Float den = (jacobian * spdf) * (jacobian * spdf) * ((neeLightPdf * neeLightPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                        weight = neeWNum / (kDEps + den + neeWDen);
                        mainContribution = neeContributionAll;
                        shiftedContribution = jacobian * sthr * (shiftedBsdfValue * neeEmitterRadiance);
                    } else {                                                         // gpt.cpp:659-705
                        const DMaterial &shiftedBSDF = c_sceneG->materials[sits.material];
                        if (atPointLight || (vertexType(mainBSDF, ESmooth) == VERTEX_TYPE_DIFFUSE && vertexType(shiftedBSDF, ESmooth) == VERTEX_TYPE_DIFFUSE)) {   // gpt.cpp:668-672
                            DRec sRec; initDRec(sits, sRec);
                            bool shiftedEmitterVisible;
                            const Spec sv = sampleEmitterDirectVisible(sRec, lsx, lsy, shiftedEmitterVisible); rays++;
                            const Spec shiftedEmitterRadiance = sv * sRec.pdf;
                            const Float shiftedDRecPdf = sRec.pdf;
                            const Float shiftedDistanceSquared = len2(neeLightP - sits.p);
                            const V3 emitterDirection = (neeLightP - sits.p) / sqrt(shiftedDistanceSquared);
                            const Float shiftedOpposingCosine = -dot(neeLightN, emitterDirection);
                            const V3 woL = toLocal(sits.sh, emitterDirection);
                            if (cfg.strictNormals && dot(sits.geoN, emitterDirection) * woL.z < 0) {
                                shiftSuccessful = false;
                            } else {
                                Spec shiftedBsdfValue; Float shiftedBsdfPdf;
                                bsdfEvalPdf(shiftedBSDF, sits.wi, woL, ESolidAngle, shiftedBsdfValue, shiftedBsdfPdf);
                                if (!shiftedEmitterVisible || atPointLight) shiftedBsdfPdf = 0;
                                const Float jacobian = fabs(shiftedOpposingCosine * neeDistSq) / (kEpsilon + fabs(neeOppCos * shiftedDistanceSquared));   // gpt.cpp:695
                                const Float den = (jacobian * spdf) * (jacobian * spdf) * ((shiftedDRecPdf * shiftedDRecPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                                weight = neeWNum / (kDEps + den + neeWDen);
                                mainContribution = neeContributionAll;
                                shiftedContribution = jacobian * sthr * (shiftedBsdfValue * shiftedEmitterRadiance);
                            }
                        }   // else: weight and both contributions stay 0 (gpt.cpp:613-615)
                    }
                }
                if (!shiftSuccessful) {                                              // gpt.cpp:708-717
                    weight = neeWNum / (kDEps + neeWDen);
                    mainContribution = neeContributionAll;
                    shiftedContribution = splat(0);
                }
                mrad = mrad + mainContribution * weight;                             // gpt.cpp:723-726
                accumulateOffset(a, o, slot, shiftedContribution * weight, (shiftedContribution - mainContribution) * weight);
            }

            if (kBsdf && bsdfStage) {                                           // ---- BSDF-sample stage, gpt.cpp:830-1151
                Spec mainContribution = splat(0), shiftedContribution = splat(0);
                Float weight = 0;
                bool postponedShiftEnd = false;
                if (alive) {
                    const Float shiftedPreviousPdf = spdf;
                    if (conn == RAY_CONNECTED) {                                     // gpt.cpp:844-861
                        sthr = sthr * (bs.weight * bs.pdf);
                        spdf *= mainBsdfPdf;
                        const Float den = (shiftedPreviousPdf * shiftedPreviousPdf) * ((mainLumPdf * mainLumPdf) + (mainBsdfPdf * mainBsdfPdf));
                        weight = mainWeightNumerator / (kDEps + den + mainWeightDenominator);
                        mainContribution = mainContributionAll;
                        shiftedContribution = sthr * mainEmitterRadiance;
                    } else if (conn == RAY_RECENTLY_CONNECTED) {                     // gpt.cpp:862-888
                        const V3 woL = toLocal(prevSh, mrayD);
                        const int measure = (bs.sampledType & EDelta) ? EDiscrete : ESolidAngle;
                        Spec shiftedBsdfValue; Float shiftedBsdfPdf;
                        bsdfEvalPdf(mainBSDF, recentWiL, woL, measure, shiftedBsdfValue, shiftedBsdfPdf);
                        sthr = sthr * shiftedBsdfValue;
                        spdf *= shiftedBsdfPdf;
                        conn = RAY_CONNECTED;
                        const Float den = (shiftedPreviousPdf * shiftedPreviousPdf) * ((mainLumPdf * mainLumPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                        weight = mainWeightNumerator / (kDEps + den + mainWeightDenominator);
                        mainContribution = mainContributionAll;
                        shiftedContribution = sthr * mainEmitterRadiance;
                    } else {                                                         // gpt.cpp:889-1126
                        const DMaterial &shiftedBSDF = c_sceneG->materials[sits.material];
                        const int shiftedVertexType = vertexType(shiftedBSDF, bs.sampledType);
                        if (mainVertexType == VERTEX_TYPE_DIFFUSE && mainNextVertexType == VERTEX_TYPE_DIFFUSE && shiftedVertexType == VERTEX_TYPE_DIFFUSE) {
                            if (!lastSegment || mainHitEmitter) {                    // gpt.cpp:901
                                const ShiftResult sr = escaped ? environmentShift(mrayD, sits.p)                   // gpt.cpp:908-915
                                                               : reconnectShift(prevP, mits.p, sits.p, mits.geoN); // gpt.cpp:907
                                rays++;
                                if (!sr.success) alive = false;
                                else {
                                    const V3 outgoingDirection = sr.wo;
                                    const V3 woL = toLocal(sits.sh, outgoingDirection);
                                    if (cfg.strictNormals && dot(outgoingDirection, sits.geoN) * woL.z <= 0) alive = false;
                                    else {
                                        Spec shiftedBsdfValue; Float shiftedBsdfPdf;
                                        bsdfEvalPdf(shiftedBSDF, sits.wi, woL, ESolidAngle, shiftedBsdfValue, shiftedBsdfPdf);   // gpt.cpp:935-936
                                        sthr = sthr * (shiftedBsdfValue * sr.jacobian);
                                        spdf *= shiftedBsdfPdf * sr.jacobian;
                                        conn = RAY_RECENTLY_CONNECTED;
                                        if (mainHitEmitter) {                        // gpt.cpp:944-985
                                            Spec shiftedEmitterRadiance; Float shiftedLumPdf;
                                            if (!escaped) {
                                                shiftedEmitterRadiance = emittedLe(mits, -outgoingDirection);
                                                DRec sd;                             // gpt.cpp:957-964 (measure: solid angle)
                                                sd.p = mainDRec.p; sd.n = mainDRec.n;
                                                sd.dist = len(mainDRec.p - sits.p);
                                                sd.d = (mainDRec.p - sits.p) / sd.dist;
                                                sd.ref = mainDRec.ref; sd.refN = sits.sh.n; sd.emitter = mainDRec.emitter;
                                                shiftedLumPdf = pdfEmitterDirect(sd);
                                                if (cfg.refUninitMeasure && c_sceneG->emitters[sd.emitter].kind != EM_ENV) shiftedLumPdf = 0;   // gpt_host.h setupArgs
                                            } else { shiftedEmitterRadiance = mainEmitterRadiance; shiftedLumPdf = mainLumPdf; }   // gpt.cpp:973-977
                                            const Float den = (shiftedPreviousPdf * shiftedPreviousPdf) * ((shiftedLumPdf * shiftedLumPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                                            weight = mainWeightNumerator / (kDEps + den + mainWeightDenominator);
                                            mainContribution = mainContributionAll;
                                            shiftedContribution = sthr * shiftedEmitterRadiance;
                                        }   // else weight and contributions stay 0 (gpt.cpp:833-836)
                                    }
                                }
                            }
                        } else {                                                     // half-vector shift, gpt.cpp:987-1126
                            Spec shiftedEmitterRadiance = splat(0);
                            const bool bothDelta = (bs.sampledType & EDelta) && (shiftedBSDF.flags & EDelta);
                            const bool bothSmooth = (bs.sampledType & ESmooth) && (shiftedBSDF.flags & ESmooth);
                            bool ok = bothDelta || bothSmooth;
                            if (ok) {
                                ShiftResult sr = halfVectorShift(prevWi, bs.wo, sits.wi, mainBSDF.bsdfEta, shiftedBSDF.bsdfEta);   // gpt.cpp:1006
                                if (bs.sampledType & EDelta) sr.jacobian = 1;        // gpt.cpp:1008-1011
                                ok = sr.success;
                                if (ok) {
                                    sthr = sthr * sr.jacobian;
                                    spdf *= sr.jacobian;
                                    const V3 tangentSpaceOutgoingDirection = sr.wo;
                                    const V3 outgoingDirection = toWorld(sits.sh, tangentSpaceOutgoingDirection);
                                    const int measure = (bs.sampledType & EDelta) ? EDiscrete : ESolidAngle;
                                    Spec ev; Float pv;
                                    bsdfEvalPdf(shiftedBSDF, sits.wi, tangentSpaceOutgoingDirection, measure, ev, pv);   // gpt.cpp:1030-1031
                                    sthr = sthr * ev;
                                    spdf *= pv;
                                    if (spdf == 0) ok = false;                       // gpt.cpp:1033-1037
                                    if (ok && cfg.strictNormals && dot(outgoingDirection, sits.geoN) * tangentSpaceOutgoingDirection.z <= 0) ok = false;
                                    if (ok) {
                                        const int shiftedVertexType2 = vertexType(shiftedBSDF, bs.sampledType);   // gpt.cpp:1047
                                        Ray sray; sray.o = sits.p; sray.d = outgoingDirection; sray.mint = kEpsilon; sray.maxt = CUDART_INF;   // gpt.cpp:1050
                                        rays++;
                                        if (!rayIntersect(sray, sits)) {             // gpt.cpp:1052-1074
                                            if (!c_scene.env.present || !escaped) ok = false;                      // no env, or env vs non-env
                                            else if (mainVertexType == VERTEX_TYPE_DIFFUSE && shiftedVertexType2 == VERTEX_TYPE_DIFFUSE) ok = false;
                                            else { shiftedEmitterRadiance = envEval(sray.d); postponedShiftEnd = true; }
                                        } else if (escaped) ok = false;               // gpt.cpp:1078-1082
                                        else {
                                            const int shiftedNextVertexType = vertexType(c_sceneG->materials[sits.material], bs.sampledType);
                                            if (mainVertexType == VERTEX_TYPE_DIFFUSE && shiftedVertexType2 == VERTEX_TYPE_DIFFUSE && shiftedNextVertexType == VERTEX_TYPE_DIFFUSE) ok = false;   // gpt.cpp:1089-1093
                                            else {
                                                if (sits.emitter >= 0) shiftedEmitterRadiance = emittedLe(sits, -sray.d);   // gpt.cpp:1095-1098
                                                storeOffIts(a, slot, i, sits);
                                            }
                                        }
                                    }
                                }
                            }
                            if (ok) {                                                // gpt.cpp:1107-1112
                                weight = mpdf / (spdf * spdf + mpdf * mpdf);
                                mainContribution = mainContributionAll;
                                shiftedContribution = sthr * shiftedEmitterRadiance;
                            } else {                                                 // gpt.cpp:1113-1125
                                weight = (Float)1 / mpdf;
                                mainContribution = mainContributionAll;
                                shiftedContribution = splat(0);
                                postponedShiftEnd = true;
                            }
                        }
                    }
                }
                if (!alive) {                                                        // gpt.cpp:1131-1136
                    weight = mainWeightNumerator / (kDEps + mainWeightDenominator);
                    mainContribution = mainContributionAll;
                    shiftedContribution = splat(0);
                }
                if (addBsdfStage) {                                                  // gpt.cpp:1140-1146
                    const bool has = !(mainContribution.x == 0 && mainContribution.y == 0 && mainContribution.z == 0 && weight == 0);
                    if (has) {
                        bHas |= 1u << i;
                        if (i == 0) bw0 = weight; else if (i == 1) bw1 = weight; else if (i == 2) bw2 = weight; else bw3 = weight;
                    }
                    accumulateOffset(a, o, slot, shiftedContribution * weight, (shiftedContribution - mainContribution) * weight);
                }
                if (postponedShiftEnd) alive = false;                                // gpt.cpp:1148-1150
                flags = setFlag(flags, i, alive, conn);
            }
            if (kBsdf && (flagAlive(flags, i) || alive)) stvw(a, o + OR_THR, slot, sthr, spdf);
        }
        // base radiance: BSDF-stage terms after all NEE terms, in offset order (gpt.cpp:1142)
        if (bHas & 1u) mrad = mrad + mainContributionAll * bw0;
        if (bHas & 2u) mrad = mrad + mainContributionAll * bw1;
        if (bHas & 4u) mrad = mrad + mainContributionAll * bw2;
        if (bHas & 8u) mrad = mrad + mainContributionAll * bw3;

        if (kBsdf && escaped) ended = true;                                          // gpt.cpp:1153-1157
        if (kBsdf && !ended) {
            if (depth++ >= cfg.rrDepth) {                                            // gpt.cpp:1159-1174
                const Float q = fmin(maxComp(mthr / mpdf) * meta * meta, (Float)0.95f);
                if (smp.next1D() >= q) ended = true;
                else {
                    mpdf *= q;
                    for (int i = 0; i < 4; ++i) W(a, BR_COUNT + i * OR_COUNT + OR_THR, slot) *= q;
                }
            }
            if (!ended && !(depth < cfg.maxDepth || cfg.maxDepth < 0)) ended = true; // gpt.cpp:537
        }
    }

    stv(a, BR_RAD, slot, mrad);
    SI(a, IF_RNGN, slot) = (int)smp.n;
    countWarp(&a.counters[1], rays);
    countWarp(&a.counters[2], ended ? (unsigned)depth : 0u);                         // gpt.cpp:1178-1179
    if (ended) {
        SI(a, IF_STATUS, slot) = ST_FINISHED;
        if (QUEUED) appendGen(a, parity ^ 1, slot);
    } else if (!kBsdf) {
        if (cfg.strictNormals) SI(a, IF_OFLAGS, slot) = (int)flags;
    } else {
        storeBaseIts(a, slot, mits);
        stvw(a, BR_RAYD, slot, mrayD, mpdf); W(a, BR_P, slot) = meta;
        stv(a, BR_THR, slot, mthr);
        SI(a, IF_DEPTH, slot) = depth; SI(a, IF_OFLAGS, slot) = (int)flags;
    }
}

#ifndef GDB_BOUNCE_MINBLOCKS
#define GDB_BOUNCE_MINBLOCKS 2      // resident CTAs/SM the register allocation is sized for (2 => 255 registers)
#endif
template <int PHASE>
__global__ void __launch_bounds__(kBounceThreads, GDB_BOUNCE_MINBLOCKS) gpt_bounce_kernel(const GptArgs a, int parity)
{
    // thread -> (BSDF-type bucket, index).  Buckets are padded to whole warps so a warp shades one BSDF
    // type; inside a bucket the slots are in ascending pixel order (gpt_compact_kernel), so the
    // struct-of-arrays state rows are still read as (near-)contiguous sectors.
    __shared__ int s_begin[kBuckets + 1], s_count[kBuckets];
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < kBuckets; b++) { const int c = a.liveCount[parity * kBuckets + b]; s_begin[b] = acc; s_count[b] = c; acc += (c + 31) & ~31; }
        s_begin[kBuckets] = acc;
        if (PHASE != 0 && blockIdx.x == 0) for (int b = 0; b < kBuckets; b++) a.liveCount[(parity ^ 1) * kBuckets + b] = 0;   // for the next step's compaction
        if (PHASE != 1 && blockIdx.x == 0) a.genCount[parity] = 0;          // this step's regeneration queue has been consumed
    }
    __syncthreads();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= s_begin[kBuckets]) return;
    int b = 0;
    while (g >= s_begin[b + 1]) b++;
    const int idx = g - s_begin[b];
    if (idx >= s_count[b]) return;
    const int slot = a.liveList[((size_t)parity * kBuckets + b) * a.nSlots + idx];
    bounceBody<PHASE, true>(a, slot, parity);
}

// Tail of the render: once few pixel streams are still running, stepping the whole wavefront costs four
// launches per bounce for a handful of warps.  Here every remaining slot is simply run to completion by
// one thread (generate -> NEE phase -> BSDF phase -> ... until its pixel's samples are exhausted).
GDB_D void runSlotToCompletion(const GptArgs &a, int slot)
{
    for (;;) {
        const int st = SI(a, IF_STATUS, slot);
        if (st == ST_DONE) break;
        if (st != ST_LIVE) { generateBody(a, slot); continue; }
        bounceBody<0, false>(a, slot, 0);
        if (SI(a, IF_STATUS, slot) == ST_LIVE) bounceBody<1, false>(a, slot, 0);
    }
}
__global__ void __launch_bounds__(kBounceThreads) gpt_tail_kernel(const GptArgs a)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.nSlots) return;
    runSlotToCompletion(a, slot);
}

// Stream compaction of the live slots, bucketed by the BSDF type of the base vertex.  Order-preserving
// inside each 256-slot chunk (warp ballots + popc ranks, a shared-memory prefix over the 8 warps, one
// atomic per chunk and bucket), so bucket lists stay sorted by pixel up to chunk granularity.
__global__ void __launch_bounds__(256) gpt_compact_kernel(const GptArgs a, int parity)
{
    __shared__ int s_warp[8][kBuckets], s_base[kBuckets];
    const int slot = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int bucket = -1;
    if (slot < a.nSlots && SI(a, IF_STATUS, slot) == ST_LIVE) {
        // stage 0: some offset path is still unconnected (shadow + reconnection / half-vector rays ahead),
        // stage 1: some offset was connected on the previous bounce (extra BSDF evaluations), stage 2: all
        // offsets ride along with the base path or are dead.  Lanes of one warp then run the same branches.
        const unsigned f = (unsigned)SI(a, IF_OFLAGS, slot);
        int stage = 2;
        for (int i = 0; i < 4; i++) {
            if (!flagAlive(f, i)) continue;
            const int c = flagConn(f, i);
            if (c == RAY_NOT_CONNECTED) stage = 0; else if (c == RAY_RECENTLY_CONNECTED && stage == 2) stage = 1;
        }
        bucket = c_sceneG->materials[SI(a, IF_MAT, slot)].type * 3 + stage;
    }
    int rank = 0;
#pragma unroll
    for (int b = 0; b < kBuckets; b++) {
        const unsigned m = __ballot_sync(0xffffffffu, bucket == b);
        if (bucket == b) rank = __popc(m & ((1u << lane) - 1));
        if (lane == 0) s_warp[warp][b] = __popc(m);
    }
    __syncthreads();
    if (threadIdx.x < kBuckets) {
        int tot = 0;
        for (int w = 0; w < 8; w++) { const int c = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = tot; tot += c; }
        s_base[threadIdx.x] = tot ? atomAdd(&a.liveCount[parity * kBuckets + threadIdx.x], tot) : 0;
    }
    __syncthreads();
    if (bucket >= 0) a.liveList[((size_t)parity * kBuckets + bucket) * a.nSlots + s_base[bucket] + s_warp[warp][bucket] + rank] = slot;
}

__global__ void gpt_init_kernel(const GptArgs a)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < 2 * kBuckets) a.liveCount[slot] = 0;
    if (slot == 0) { a.genCount[0] = a.nSlots; a.genCount[1] = 0; a.counters[6] = (unsigned long long)a.nSlots; }
    if (slot >= a.nSlots) return;
    a.genList[slot] = slot;
    SI(a, IF_STATUS, slot) = ST_FRESH; SI(a, IF_SAMPLE, slot) = 0; SI(a, IF_RNGN, slot) = 0; SI(a, IF_STREAM, slot) = slot;   // slot s starts on stream s
}

// Self-check of the candidate selection in closestPrimitive: random nearest-hit and shadow rays (from
// surface points, between surface points, from free space) must give bit-identical answers with and
// without the bounds pass.
__global__ void gpt_check_culling_kernel(unsigned long long seed, int nRays, unsigned long long *mismatch)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nRays) return;
    Sampler smp; smp.key = samplerKey(seed, g, 12345); smp.n = 0;
    auto surfacePoint = [&]() {
        const int nR = c_scene.nRects, nT = c_scene.nTris, nS = c_scene.nSpheres, nB = c_scene.nBvhTris;
        const int pick = min(nR + nT + nS + nB - 1, (int)(smp.next1D() * (nR + nT + nS + nB)));
        const Float u = smp.next1D(), v = smp.next1D();
        if (pick < nR) return xfAffine(c_sceneG->rects[pick].toWorld, mk(2 * u - 1, 2 * v - 1, 0));
        if (pick < nR + nT) { const DTri &T = c_sceneG->tris[pick - nR]; const Float a = sqrt(u); return T.p0 * (1 - a) + T.p1 * (a * (1 - v)) + T.p2 * (a * v); }
        if (pick >= nR + nT + nS) { const DTri &T = c_scene.bvhTris[pick - nR - nT - nS]; const Float a = sqrt(u); return T.p0 * (1 - a) + T.p1 * (a * (1 - v)) + T.p2 * (a * v); }
        const DSphere &sp = c_sceneG->spheres[pick - nR - nT];
        const Float z = 1 - 2 * u, r = sqrt(fmax(0.0, 1 - z * z)), phi = 2 * kPi * v;
        return sp.center + mk(r * cos(phi), r * sin(phi), z) * sp.radius;
    };
    Ray ray;
    const int flavour = g % 3;
    if (flavour == 0) {            // extension ray from a surface point
        ray.o = surfacePoint();
        const Float z = 1 - 2 * smp.next1D(), r = sqrt(fmax(0.0, 1 - z * z)), phi = 2 * kPi * smp.next1D();
        ray.d = mk(r * cos(phi), r * sin(phi), z); ray.mint = kEpsilon; ray.maxt = CUDART_INF;
    } else if (flavour == 1) {     // visibility segment between two surface points (gpt.cpp:84-93)
        ray.o = surfacePoint(); ray.d = surfacePoint() - ray.o; ray.mint = kEpsilon; ray.maxt = 1.0 - kShadowEpsilon;
    } else {                       // camera-like ray from free space
        ray.o = mk((2 * smp.next1D() - 1) * 2, (2 * smp.next1D() - 1) * 2, (2 * smp.next1D() - 1) * 5);
        ray.d = normalize(surfacePoint() - ray.o); ray.mint = 1e-2; ray.maxt = 1e4;
    }
    Float rayMinT = ray.mint;
    if (rayMinT == kEpsilon) rayMinT *= fmax(maxAbs3(ray.o), kEpsilon);
    Float t1 = 0, t2 = 0, u1 = 0, v1 = 0, u2 = 0, v2 = 0; int k1 = -1, i1 = -1, k2 = -1, i2 = -1;
    const bool h1 = closestPrimitive<false>(ray, rayMinT, ray.maxt, t1, k1, i1, u1, v1);
    const bool h2 = closestPrimitiveExhaustive<false>(ray, rayMinT, ray.maxt, t2, k2, i2, u2, v2);
    Float tt = 0, uu = 0, vv = 0; int kk = -1, ii = -1;
    const bool a1 = closestPrimitive<true>(ray, rayMinT, ray.maxt, tt, kk, ii, uu, vv);
    const bool a2 = closestPrimitiveExhaustive<true>(ray, rayMinT, ray.maxt, tt, kk, ii, uu, vv);
    bool bad = (h1 != h2) || (a1 != a2) || (h1 != a1);
    if (h1 && h2) bad = bad || t1 != t2 || k1 != k2 || i1 != i2 || (k1 >= 2 && (u1 != u2 || v1 != v2));
    if (bad) redAdd(mismatch, 1ULL);
    if (h1) redAdd(mismatch + 1, 1ULL);
}

// MultiFilm::developMulti (multifilm.cpp:366-416, fmtconv.cpp:1036-1045): value * (1/weight), plus the
// Float -> float conversion of gpt.cpp:1439-1442 for the solver inputs.
__global__ void gpt_develop_kernel(const double *film, int n, double *dev64 /*[5][n][3]*/, float *dev32 /*[5][n][3]*/)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 5 * n) return;
    const double *p = film + (size_t)i * 4;
    const double wgt = p[3], inv = (wgt != 0) ? 1 / wgt : wgt;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double v = p[c] * inv;
        dev64[(size_t)i * 3 + c] = v;
        dev32[(size_t)i * 3 + c] = (float)v;
    }
}

}  // namespace gdb200
