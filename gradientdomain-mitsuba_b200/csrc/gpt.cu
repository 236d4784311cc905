// gdb200 G-PT wavefront tracer for sm_100a (fp64).
//
// Replaces GradientPathIntegrator::render's block scheduler + renderBlock + evaluatePoint +
// evaluate (reference src/integrators/gpt/gpt.cpp:397-436, 468-1180, 1220-1355) with a
// wavefront over persistent per-pixel path slots:
//
//   * one slot per base pixel owns that pixel's sample stream (the gdb200_counter sampler is
//     re-keyed per pixel exactly where the reference calls Sampler::generate, gpt.cpp:1250, and
//     the spp samples of a pixel consume it sequentially, so results do not depend on scheduling);
//   * each wavefront step launches
//       gpt_generate_kernel : the slots of the regeneration queue splat the 15 film contributions of
//                             their finished path (gpt.cpp:1319-1352) and start their next sample:
//                             5 camera rays + primary hits, very-direct emission;
//       gpt_compact_kernel  : ballot/popc stream compaction of the live slots into 12 queues =
//                             BSDF type of the base vertex x shift stage of the offset paths
//                             (order-preserving per 256-slot chunk), so a warp shades one BSDF type with
//                             its offset paths in the same connection state;
//       gpt_bounce_kernel   : one iteration of the reference's bounce loop for every queued slot: NEE
//                             with the 4-strategy MIS, BSDF sample, extension ray, reconnection /
//                             half-vector shift of the 4 offset paths, Russian roulette;
//     the last few pixel streams are run to completion by gpt_tail_kernel in one launch;
//   * state lives in HBM as 32-byte records [record][slot][4 fp64] (one DRAM sector each), so queues
//     of scattered slots never over-fetch;
//   * film accumulators are fp64 value+weight planes updated with red.global.add.f64.
//
// Arithmetic follows the reference's operation order (compiled with -fmad=false) so that the
// fp64 buffers match the CPU oracle to rounding of the transcendental functions.
#include "common.h"
#include "gpt_device.cuh"
#include <math_constants.h>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

namespace gdb200 {


// ------------------------------------------------------------------ state layout
// fp64 state is stored as 32-byte records [record][slot][4]: one vector (and one scalar riding in its 4th
// lane) per record.  A record is exactly one DRAM sector, so a lane always consumes every byte it
// fetches, however scattered the slots of a material/stage queue are.
enum BaseRec { BR_RAYD = 0 /* w: path pdf */, BR_P /* w: eta */, BR_GN /* w: sample x */, BR_S /* w: sample y */, BR_T, BR_N, BR_WI,
               BR_THR, BR_RAD, BR_VD, BR_COUNT };
enum OffRec { OR_THR = 0 /* w: path pdf */, OR_RAD, OR_GRAD, OR_P, OR_GN, OR_S, OR_T, OR_N, OR_WI, OR_COUNT };
constexpr int kRecords = BR_COUNT + 4 * OR_COUNT;   // 46 records = 1472 B per slot
enum IntField { IF_STATUS = 0, IF_MAT, IF_EMI, IF_DEPTH, IF_SAMPLE, IF_RNGN, IF_OFLAGS, IF_PAD, IF_OMAT0, IF_OMAT1, IF_OMAT2, IF_OMAT3,
                IF_COUNT };
enum SlotStatus { ST_FRESH = 0, ST_LIVE = 1, ST_FINISHED = 2, ST_DONE = 3 };
enum { RAY_NOT_CONNECTED = 0, RAY_RECENTLY_CONNECTED = 1, RAY_CONNECTED = 2 };
enum { BUF_FINAL = 0, BUF_THROUGHPUT = 1, BUF_DX = 2, BUF_DY = 3, BUF_DIRECT = 4 };

constexpr int kBounceThreads = 128, kGenThreads = 128;
constexpr int kBuckets = 12;  // one queue per (BSDF type of the base vertex) x (shift stage of the offset paths)

struct GptArgs {
    double *sd;            // [kRecords][nSlots][4]
    int *si;               // [nSlots][16]
    int nSlots, width, height, yBegin;
    int bandRows, bandCount, bandIndex, pad1;   // interleaved row bands (bandCount > 1) instead of one strip
    int spp, skipPreview;   // skipPreview: the "-final" preview puts (gpt.cpp:1319-1324) are not needed when a reconstruction overwrites that buffer
    uint64_t seed;
    Config cfg;
    double *film;          // [5][H][W][4]
    int *liveList;         // [2][kBuckets][nSlots]
    int *liveCount;        // [2][kBuckets]
    int *genList;          // [2][nSlots]: slots whose path ended (to splat + regenerate)
    int *genCount;         // [2]
    unsigned long long *counters;   // [0] done slots, [1] rays, [2] path vertices, [3] samples, [4] state bytes, [5] path bounces
};

GDB_D double *REC(const GptArgs &a, int rec, int slot) { return a.sd + (((size_t)rec * a.nSlots + slot) << 2); }
GDB_D double &W(const GptArgs &a, int rec, int slot) { return REC(a, rec, slot)[3]; }
// int fields of a slot share one 64-byte line [slot][16] (fields 0-7 in its first sector), so a kernel pulls one
// sector per slot instead of one per field; the per-pixel sampler key is recomputed, not stored.
GDB_D int &SI(const GptArgs &a, int field, int slot) { return a.si[((size_t)slot << 4) + field]; }
GDB_D V3 ldv(const GptArgs &a, int rec, int slot)
{
    const double2 *p = reinterpret_cast<const double2 *>(REC(a, rec, slot));
    const double2 lo = p[0], hi = p[1];
    return mk(lo.x, lo.y, hi.x);
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
GDB_D void accumulateOffset(const GptArgs &a, int o, int slot, Spec dRad, Spec dGrad)
{
    if (!(dRad.x == 0 && dRad.y == 0 && dRad.z == 0)) stv(a, o + OR_RAD, slot, ldv(a, o + OR_RAD, slot) + dRad);
    if (!(dGrad.x == 0 && dGrad.y == 0 && dGrad.z == 0)) stv(a, o + OR_GRAD, slot, ldv(a, o + OR_GRAD, slot) + dGrad);
}

// Warp-aggregated append of an ended slot to the next step's regeneration queue.
GDB_D void appendGen(const GptArgs &a, int parity, int slot)
{
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&a.genCount[parity], __popc(m));
    base = __shfl_sync(m, base, leader);
    a.genList[(size_t)parity * a.nSlots + base + __popc(m & ((1u << lane) - 1))] = slot;
}

// Image row of a slot: one contiguous strip, or interleaved bands of bandRows rows dealt round-robin to bandCount ranks.
GDB_D int slotRow(const GptArgs &a, int slot)
{
    const int lr = slot / a.width;
    if (a.bandCount <= 1) return a.yBegin + lr;
    return ((lr / a.bandRows) * a.bandCount + a.bandIndex) * a.bandRows + lr % a.bandRows;
}

// One atomic per warp for statistics counters.
GDB_D void countWarp(unsigned long long *ctr, unsigned v)
{
    const unsigned m = __activemask();
    const unsigned s = __reduce_add_sync(m, v);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1 && s) atomicAdd(ctr, (unsigned long long)s);
}

// ------------------------------------------------------------------ film (ImageBlock::put, imageblock.h:150-195)
GDB_D Float evalDiscretized(Float x)      // rfilter.h:76-77, MTS_FILTER_RESOLUTION = 31; box taps = 1/(2r) (rfilter.cpp:37-55)
{
    const int idx = min((int)fabs(x * c_scene.filterScale), 31);
    return idx < 31 ? c_scene.filterTap : 0.0;
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
            for (int k = 0; k < 4; k++) atomicAdd(dst + k, wgt * value[k]);
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
    const int px = slot % a.width, py = slotRow(a, slot);

    if (SI(a, IF_STATUS, slot) == ST_FINISHED) {
        Spec rad[4], grad[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { const int o = BR_COUNT + i * OR_COUNT; rad[i] = ldv(a, o + OR_RAD, slot); grad[i] = ldv(a, o + OR_GRAD, slot); }
        splatSample(a, W(a, BR_GN, slot), W(a, BR_S, slot), ldv(a, BR_VD, slot), ldv(a, BR_RAD, slot), rad, grad);
    }

    Sampler smp; smp.key = samplerKey(a.seed, px, py); smp.n = (uint32_t)SI(a, IF_RNGN, slot);    // Sampler::generate, gpt.cpp:1250-1251
    int j = SI(a, IF_SAMPLE, slot);
    unsigned long long rays = 0, samples = 0;
    int status = ST_DONE;
    while (j < a.spp) {
        j++; samples++;
        const Float u = smp.next1D(), v = smp.next1D();                              // gpt.cpp:1261
        const Float spx = px + u, spy = py + v;
        Ray ray; Its mits;
        sampleCameraRay(spx, spy, ray);                                              // gpt.cpp:402
        const bool mainValid = rayIntersect(ray, mits); rays += 5;                   // gpt.cpp:472
        Spec veryDirect = splat(0);
        unsigned flags = 0;
        bool early = !mainValid;                                                     // gpt.cpp:482-492 (no environment emitter)
        if (mainValid && mits.emitter >= 0) veryDirect = veryDirect + splat(1.0) * emittedLe(mits, -ray.d);   // gpt.cpp:497-499
        if (mainValid && a.cfg.strictNormals && dot(ray.d, mits.geoN) * mits.wi.z >= 0) early = true;          // gpt.cpp:518-521
        const Float shiftX[4] = {1, 0, -1, 0}, shiftY[4] = {0, 1, 0, -1};            // gpt.cpp:410-415
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            Ray sray; Its sits;
            sampleCameraRay(spx + shiftX[i], spy + shiftY[i], sray);                 // gpt.cpp:418
            bool alive = rayIntersect(sray, sits);                                   // gpt.cpp:476-480, 508-513
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
            if (!early) atomicAdd(&a.counters[2], 1ULL);                             // avgPathLength += depth (1), gpt.cpp:1178-1179
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
    SI(a, IF_STATUS, slot) = status; SI(a, IF_SAMPLE, slot) = j; SI(a, IF_RNGN, slot) = (int)smp.n;
    countWarp(&a.counters[0], status == ST_DONE ? 1u : 0u);
    countWarp(&a.counters[1], (unsigned)rays);
    countWarp(&a.counters[3], (unsigned)samples);
}

__global__ void __launch_bounds__(kGenThreads) gpt_generate_kernel(const GptArgs a, int parity)
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

    Its mits; loadBaseIts(a, slot, mits);
    V3 mrayD; Float mpdf;
    ldvw(a, BR_RAYD, slot, mrayD, mpdf);
    Spec mthr = ldv(a, BR_THR, slot), mrad = ldv(a, BR_RAD, slot);
    Float meta = W(a, BR_P, slot);
    int depth = SI(a, IF_DEPTH, slot);
    unsigned flags = (unsigned)SI(a, IF_OFLAGS, slot);
    Sampler smp; smp.key = samplerKey(a.seed, slot % a.width, slotRow(a, slot)); smp.n = (uint32_t)SI(a, IF_RNGN, slot);
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
        bool neeActive = false, neeVisible = false;
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
            if (!neeVisible) neeBsdfPdf = 0;                                         // gpt.cpp:592
            neeDistSq = len2(mits.p - dRec.p);                                       // gpt.cpp:595-596
            neeOppCos = dot(dRec.n, (mits.p - dRec.p)) / sqrt(neeDistSq);
            neeWNum = mpdf * dRec.pdf;                                               // gpt.cpp:599-600
            neeWDen = (mpdf * mpdf) * ((dRec.pdf * dRec.pdf) + (neeBsdfPdf * neeBsdfPdf));
            neeLightP = dRec.p; neeLightN = dRec.n; neeLightPdf = dRec.pdf;
            neeActive = !cfg.strictNormals || dot(mits.geoN, dRec.d) * neeWoLocal.z > 0;   // gpt.cpp:607
            neeContributionAll = mthr * (neeBsdfValue * neeEmitterRadiance);
        }

        // ---------------- base path: BSDF sample + extension, gpt.cpp:737-820
        bool bsdfStage = false, mainHitEmitter = false;
        BSDFSample bs;
        bs.weight = splat(0); bs.pdf = 0; bs.eta = 1.0; bs.sampledType = 0; bs.wo = mk(0, 0, 0);
        if (kBsdf) { const Float sx = smp.next1D(), sy = smp.next1D(); bsdfSample(mainBSDF, mits.wi, sx, sy, bs); }   // gpt.cpp:456-457
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
                if (!rayIntersect(mray, mits)) ended = true;                         // gpt.cpp:800-803 (no environment emitter)
                else {
                    bsdfStage = true;
                    mrayD = mainWo;
                    if (mits.emitter >= 0) {                                         // gpt.cpp:771-776
                        mainEmitterRadiance = emittedLe(mits, -mainWo);
                        mainDRec.p = mits.p; mainDRec.n = mits.sh.n; mainDRec.d = mainWo; mainDRec.dist = mits.t; mainDRec.emitter = mits.emitter;
                        mainHitEmitter = true;
                    }
                    mainNextVertexType = vertexType(c_sceneG->materials[mits.material], bs.sampledType);   // gpt.cpp:784
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
                        if (!neeVisible) shiftedBsdfPdf = 0;
                        const Float jacobian = 1;
                        const Float den = (jacobian * spdf) * (jacobian * spdf) * ((neeLightPdf * neeLightPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                        weight = neeWNum / (kDEps + den + neeWDen);
                        mainContribution = neeContributionAll;
                        shiftedContribution = jacobian * sthr * (shiftedBsdfValue * neeEmitterRadiance);
                    } else {                                                         // gpt.cpp:659-705
                        const DMaterial &shiftedBSDF = c_sceneG->materials[sits.material];
                        if (vertexType(mainBSDF, ESmooth) == VERTEX_TYPE_DIFFUSE && vertexType(shiftedBSDF, ESmooth) == VERTEX_TYPE_DIFFUSE) {   // gpt.cpp:672
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
                                if (!shiftedEmitterVisible) shiftedBsdfPdf = 0;
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
                                const ShiftResult sr = reconnectShift(prevP, mits.p, sits.p, mits.geoN); rays++;   // gpt.cpp:907
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
                                            const Spec shiftedEmitterRadiance = emittedLe(mits, -outgoingDirection);
                                            DRec sd;                                 // gpt.cpp:957-964 (measure: solid angle)
                                            sd.p = mainDRec.p; sd.n = mainDRec.n;
                                            sd.dist = len(mainDRec.p - sits.p);
                                            sd.d = (mainDRec.p - sits.p) / sd.dist;
                                            sd.ref = mainDRec.ref; sd.refN = sits.sh.n; sd.emitter = mainDRec.emitter;
                                            const Float shiftedLumPdf = pdfEmitterDirect(sd);
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
                                        if (!rayIntersect(sray, sits)) ok = false;   // gpt.cpp:1052-1058 (no environment emitter)
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

template <int PHASE>
__global__ void __launch_bounds__(kBounceThreads) gpt_bounce_kernel(const GptArgs a, int parity)
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
__global__ void __launch_bounds__(kBounceThreads) gpt_tail_kernel(const GptArgs a)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.nSlots) return;
    for (;;) {
        const int st = SI(a, IF_STATUS, slot);
        if (st == ST_DONE) break;
        if (st != ST_LIVE) { generateBody(a, slot); continue; }
        bounceBody<0, false>(a, slot, 0);
        if (SI(a, IF_STATUS, slot) == ST_LIVE) bounceBody<1, false>(a, slot, 0);
    }
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
        s_base[threadIdx.x] = tot ? atomicAdd(&a.liveCount[parity * kBuckets + threadIdx.x], tot) : 0;
    }
    __syncthreads();
    if (bucket >= 0) a.liveList[((size_t)parity * kBuckets + bucket) * a.nSlots + s_base[bucket] + s_warp[warp][bucket] + rank] = slot;
}

__global__ void gpt_init_kernel(const GptArgs a)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < 2 * kBuckets) a.liveCount[slot] = 0;
    if (slot == 0) { a.genCount[0] = a.nSlots; a.genCount[1] = 0; }
    if (slot >= a.nSlots) return;
    a.genList[slot] = slot;
    SI(a, IF_STATUS, slot) = ST_FRESH; SI(a, IF_SAMPLE, slot) = 0; SI(a, IF_RNGN, slot) = 0;
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
        const int nR = c_scene.nRects, nT = c_scene.nTris, nS = c_scene.nSpheres;
        const int pick = min(nR + nT + nS - 1, (int)(smp.next1D() * (nR + nT + nS)));
        const Float u = smp.next1D(), v = smp.next1D();
        if (pick < nR) return xfAffine(c_sceneG->rects[pick].toWorld, mk(2 * u - 1, 2 * v - 1, 0));
        if (pick < nR + nT) { const DTri &T = c_sceneG->tris[pick - nR]; const Float a = sqrt(u); return T.p0 * (1 - a) + T.p1 * (a * (1 - v)) + T.p2 * (a * v); }
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
    if (h1 && h2) bad = bad || t1 != t2 || k1 != k2 || i1 != i2 || (k1 == 2 && (u1 != u2 || v1 != v2));
    if (bad) atomicAdd(mismatch, 1ULL);
    if (h1) atomicAdd(mismatch + 1, 1ULL);
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

// ======================================================================================= host ==

using namespace gdb200;

struct gdb200_scene {
    int device = 0;
    DScene host;                       // flattened tables (vertex classification filled per render)
    DBounds bounds[kMaxPrims];         // padded per-primitive bounds (candidate selection)
    DScene *dScene = nullptr;          // global-memory copy of `host` for per-lane indexed reads
    std::vector<gdb200_material> mats;
    int width = 0, height = 0;
    // device buffers
    double *film = nullptr, *dev64 = nullptr; float *dev32 = nullptr;
    double *sd = nullptr; int *si = nullptr;
    int *liveList = nullptr, *liveCount = nullptr, *genList = nullptr, *genCount = nullptr;
    unsigned long long *counters = nullptr;
    int slotCapacity = 0;
    volatile int cancel = 0;
};

static std::mutex g_constMutex;    // c_scene is one per device: serialise renders that share a device

namespace {

int flattenScene(const gdb200_scene_desc *d, gdb200_scene *s)
{
    DScene &h = s->host;
    memset(&h, 0, sizeof(h));
    const gdb200_camera &c = d->camera;
    if (c.width <= 0 || c.height <= 0) return set_error(GDB200_ERR_ARGUMENT, "invalid film size %dx%d", c.width, c.height);
    if (d->n_emitters < 1) return set_error(GDB200_ERR_ARGUMENT, "scene has no emitter");
    if (d->n_materials > kMaxMaterials || d->n_emitters > kMaxEmitters)
        return set_error(GDB200_ERR_ARGUMENT, "too many materials/emitters (%d/%d, limits %d/%d)", d->n_materials, d->n_emitters, kMaxMaterials, kMaxEmitters);
    memcpy(h.sampleToCamera, c.sample_to_camera, sizeof(h.sampleToCamera));
    memcpy(h.cameraToWorld, c.camera_to_world, sizeof(h.cameraToWorld));
    h.nearClip = c.near_clip; h.farClip = c.far_clip; h.width = c.width; h.height = c.height;
    h.invResX = 1.0 / c.width; h.invResY = 1.0 / c.height;
    h.filterRadius = d->rfilter_radius; h.filterTap = 1.0 / (2 * d->rfilter_radius); h.filterScale = 31 / d->rfilter_radius;
    s->width = c.width; s->height = c.height;
    s->mats.assign(d->materials, d->materials + d->n_materials);
    h.nMaterials = d->n_materials;
    std::vector<int> rectOfShape(d->n_shapes, -1);
    for (int i = 0; i < d->n_shapes; i++) {
        const gdb200_shape &sh = d->shapes[i];
        if (sh.material < 0 || sh.material >= d->n_materials) return set_error(GDB200_ERR_ARGUMENT, "shape %d: bad material index", i);
        if (sh.type == GDB200_SHAPE_RECTANGLE) {                                     // rectangle.cpp:100-110
            if (h.nRects >= kMaxRects) return set_error(GDB200_ERR_ARGUMENT, "too many rectangles (limit %d)", kMaxRects);
            if (sh.to_world[12] != 0 || sh.to_world[13] != 0 || sh.to_world[14] != 0 || sh.to_world[15] != 1)
                return set_error(GDB200_ERR_ARGUMENT, "shape %d: toWorld must be affine", i);
            DRect &r = h.rects[h.nRects];
            memcpy(r.toObject, sh.to_object, sizeof(r.toObject)); memcpy(r.toWorld, sh.to_world, sizeof(r.toWorld));
            r.dpdu = xfVector(sh.to_world, mk(2, 0, 0));
            const V3 dpdv = xfVector(sh.to_world, mk(0, 2, 0));
            r.n = normalize(xfNormal(sh.to_object, mk(0, 0, 1)));
            r.invArea = 1.0 / (len(r.dpdu) * len(dpdv));
            r.material = sh.material; r.emitter = sh.emitter;
            rectOfShape[i] = h.nRects++;
        } else if (sh.type == GDB200_SHAPE_SPHERE) {
            if (h.nSpheres >= kMaxSpheres) return set_error(GDB200_ERR_ARGUMENT, "too many spheres (limit %d)", kMaxSpheres);
            if (sh.emitter >= 0) return set_error(GDB200_ERR_ARGUMENT, "shape %d: sphere emitters are not supported yet", i);
            DSphere &sp = h.spheres[h.nSpheres++];
            sp.center = mk(sh.center[0], sh.center[1], sh.center[2]); sp.radius = sh.radius; sp.flip = sh.flip_normals;
            sp.material = sh.material; sp.emitter = -1;
        } else if (sh.type == GDB200_SHAPE_MESH) {
            if (sh.emitter >= 0) return set_error(GDB200_ERR_ARGUMENT, "shape %d: mesh emitters are not supported yet", i);
            if (h.nMeshes >= kMaxMeshes) return set_error(GDB200_ERR_ARGUMENT, "too many meshes (limit %d)", kMaxMeshes);
            DMesh &M = h.meshes[h.nMeshes++];
            M.first = h.nTris; M.count = 0;
            const double big = std::numeric_limits<double>::infinity();
            M.lo = mk(big, big, big); M.hi = mk(-big, -big, -big);
            std::vector<DTri> meshTris;
            for (int t = sh.first_tri; t < sh.first_tri + sh.tri_count; t++) {
                if (h.nTris + (int)meshTris.size() >= kMaxTris) return set_error(GDB200_ERR_ARGUMENT, "too many triangles for the constant-memory scene table (limit %d); the BVH path is not built yet", kMaxTris);
                if (t < 0 || t >= d->n_triangles) return set_error(GDB200_ERR_ARGUMENT, "shape %d: triangle range out of bounds", i);
                const int *ix = d->triangles + 3 * t;
                const double *va = d->vertices + 3 * ix[0], *vb = d->vertices + 3 * ix[1], *vc = d->vertices + 3 * ix[2];
                const V3 A = mk(va[0], va[1], va[2]), B = mk(vb[0], vb[1], vb[2]), C = mk(vc[0], vc[1], vc[2]);
                for (const V3 &P : {A, B, C}) {
                    M.lo = mk(std::min(M.lo.x, P.x), std::min(M.lo.y, P.y), std::min(M.lo.z, P.z));
                    M.hi = mk(std::max(M.hi.x, P.x), std::max(M.hi.y, P.y), std::max(M.hi.z, P.z));
                }
                DTri T;                                                              // TriAccel::load, triaccel.h:61-95
                memset(&T, 0, sizeof(T));
                static const int waldModulo[4] = {1, 2, 0, 1};
                const V3 b = C - A, cc = B - A, N = cross(cc, b);
                const double Nv[3] = {N.x, N.y, N.z}, bv[3] = {b.x, b.y, b.z}, cv[3] = {cc.x, cc.y, cc.z}, Av[3] = {A.x, A.y, A.z};
                int k = 0;
                for (int j = 0; j < 3; j++) if (std::abs(Nv[j]) > std::abs(Nv[k])) k = j;
                const int u = waldModulo[k], v = waldModulo[k + 1];
                const double n_k = Nv[k], denom = bv[u] * cv[v] - bv[v] * cv[u];
                T.p0 = A; T.p1 = B; T.p2 = C; T.material = sh.material; T.emitter = -1;
                if (denom == 0) continue;                                            // degenerate: k = 3, never hit (triaccel.h:75-78)
                T.k = k;
                T.n_u = Nv[u] / n_k; T.n_v = Nv[v] / n_k; T.n_d = dot(A, N) / n_k;
                T.b_nu = bv[u] / denom; T.b_nv = -bv[v] / denom; T.a_u = Av[u]; T.a_v = Av[v];
                T.c_nu = cv[v] / denom; T.c_nv = -cv[u] / denom;
                V3 faceNormal = cross(B - A, C - A);                                 // skdtree.h:367-371
                const double l = len(faceNormal);
                if (!isZero(faceNormal)) faceNormal = faceNormal / l;
                T.faceNormal = faceNormal;
                meshTris.push_back(T);
            }
            for (int k = 0; k < 3; k++) {          // store grouped by projection axis (order inside a group is kept)
                for (const DTri &T : meshTris) if (T.k == k) h.tris[h.nTris++] = T;
                M.kEnd[k] = h.nTris;
            }
            M.count = h.nTris - M.first;
        } else return set_error(GDB200_ERR_ARGUMENT, "shape %d: unknown type %d", i, sh.type);
    }
    for (int mi = 0; mi < h.nMeshes; mi++) {   // enlarge the skip-bounds far beyond any rounding of the slab test
        DMesh &M = h.meshes[mi];
        const V3 ext = M.hi - M.lo;
        const double pad = 1e-6 * std::max(1.0, std::max(ext.x, std::max(ext.y, ext.z))) + 1e-9 * std::max(maxComp(M.hi), -std::min(M.lo.x, std::min(M.lo.y, M.lo.z)));
        M.lo = M.lo - splat(pad); M.hi = M.hi + splat(pad);
    }
    // padded bounds of every primitive for the candidate pass of closestPrimitive
    {
        double scale = 0;
        for (int k = 0; k < 3; k++) scale = std::max(scale, std::abs(c.camera_to_world[4 * k + 3]));
        int np = 0;
        auto grow = [&](DBounds &B, V3 P) {
            const double v[3] = {P.x, P.y, P.z};
            for (int k = 0; k < 3; k++) { B.lo[k] = std::min(B.lo[k], (float)v[k]); B.hi[k] = std::max(B.hi[k], (float)v[k]); scale = std::max(scale, std::abs(v[k])); }
        };
        auto reset = [](DBounds &B) { for (int k = 0; k < 3; k++) { B.lo[k] = std::numeric_limits<float>::infinity(); B.hi[k] = -B.lo[k]; } };
        for (int i = 0; i < h.nRects; i++) {
            DBounds &B = s->bounds[np++]; reset(B);
            for (int sx = -1; sx <= 1; sx += 2) for (int sy = -1; sy <= 1; sy += 2) grow(B, xfAffine(h.rects[i].toWorld, mk(sx, sy, 0)));
        }
        for (int i = 0; i < h.nSpheres; i++) {
            DBounds &B = s->bounds[np++]; reset(B);
            grow(B, h.spheres[i].center - splat(h.spheres[i].radius)); grow(B, h.spheres[i].center + splat(h.spheres[i].radius));
        }
        for (int i = 0; i < h.nTris; i++) {
            DBounds &B = s->bounds[np++]; reset(B);
            grow(B, h.tris[i].p0); grow(B, h.tris[i].p1); grow(B, h.tris[i].p2);
        }
        for (int i = 0; i < np; i++)        // pad by 1e-4 of the scene scale: >100x the fp32 rounding of the slab test (errors and
            for (int k = 0; k < 3; k++) {   // padding both scale with |1/d| per axis, so the margin holds for any ray direction)
                const float pad = (float)(1e-4 * (scale + (s->bounds[i].hi[k] - s->bounds[i].lo[k])) + 1e-6);
                s->bounds[i].lo[k] -= pad; s->bounds[i].hi[k] += pad;
            }
    }
    // emitters: DiscreteDistribution over samplingWeight (scene.cpp:357-380, pmf.h:100-114)
    h.nEmitters = d->n_emitters;
    h.emCdf[0] = 0.0;
    for (int i = 0; i < d->n_emitters; i++) h.emCdf[i + 1] = h.emCdf[i] + d->emitters[i].sampling_weight;
    const double sum = h.emCdf[d->n_emitters], norm = sum > 0 ? 1.0 / sum : 0.0;
    if (sum > 0) { for (int i = 1; i <= d->n_emitters; i++) h.emCdf[i] *= norm; h.emCdf[d->n_emitters] = 1.0; }
    for (int i = 0; i < d->n_emitters; i++) {
        const gdb200_emitter &e = d->emitters[i];
        if (e.shape < 0 || e.shape >= d->n_shapes || rectOfShape[e.shape] < 0)
            return set_error(GDB200_ERR_ARGUMENT, "emitter %d: only rectangle area emitters are supported", i);
        h.emitters[i].rect = rectOfShape[e.shape];
        h.emitters[i].radiance = mk(e.radiance[0], e.radiance[1], e.radiance[2]);
        h.emitters[i].pdfDiscrete = e.sampling_weight * norm;
    }
    return GDB200_OK;
}

// Per-material facts incl. the vertex classification of gpt.cpp:176-226 for this shiftThreshold.
void classifyMaterials(gdb200_scene *s, double shiftThreshold)
{
    for (size_t i = 0; i < s->mats.size(); i++) {
        const gdb200_material &m = s->mats[i];
        DMaterial &o = s->host.materials[i];
        o.type = m.type; o.distribution = m.distribution;
        o.reflectance = mk(m.reflectance[0], m.reflectance[1], m.reflectance[2]);
        o.specR = mk(m.specular_reflectance[0], m.specular_reflectance[1], m.specular_reflectance[2]);
        o.specT = mk(m.specular_transmittance[0], m.specular_transmittance[1], m.specular_transmittance[2]);
        o.eta = mk(m.eta[0], m.eta[1], m.eta[2]); o.k = mk(m.k[0], m.k[1], m.k[2]);
        o.alpha = std::max(m.alpha, (double)1e-4f);                                  // microfacet.h:67-71
        o.iorRatio = m.ior_ratio;
        o.bsdfEta = m.type == GDB200_BSDF_DIELECTRIC ? m.ior_ratio : 1.0;            // bsdf.cpp:62-64, dielectric.cpp:389
        int nComp = 1; double rough[2] = {0, 0};
        const double inf = std::numeric_limits<double>::infinity();
        switch (m.type) {
            case GDB200_BSDF_DIFFUSE:                                                // diffuse.cpp:97-101,167-169
                o.flags = std::max(m.reflectance[0], std::max(m.reflectance[1], m.reflectance[2])) > 0 ? (EDiffuseReflection | EFrontSide) : 0;
                nComp = o.flags ? 1 : 0; rough[0] = inf; break;
            case GDB200_BSDF_ROUGHCONDUCTOR: o.flags = EGlossyReflection | EFrontSide; rough[0] = 0.5 * (m.alpha + m.alpha); break;   // roughconductor.cpp:437-440
            case GDB200_BSDF_CONDUCTOR: o.flags = EDeltaReflection | EFrontSide; rough[0] = 0; break;
            default: o.flags = EDeltaReflection | EDeltaTransmission | EFrontSide | EBackSide; nComp = 2; break;
        }
        o.refNFromShading = (o.flags & (ETransmissionBits | EBackSide)) == 0;        // records.inl:160-165
        for (int deltaQuery = 0; deltaQuery < 2; deltaQuery++) {                     // gpt.cpp:194-226
            double lowest = inf; bool found_smooth = false, found_dirac = false;
            for (int c = 0; c < nComp; c++) {
                const double r = rough[c];
                if (r == 0) { found_dirac = true; if (!deltaQuery) continue; } else found_smooth = true;
                if (r < lowest) lowest = r;
            }
            if (!found_smooth && found_dirac && !deltaQuery) lowest = 0;
            (deltaQuery ? o.vtDelta : o.vtSmooth) = lowest <= shiftThreshold ? VERTEX_TYPE_GLOSSY : VERTEX_TYPE_DIFFUSE;
        }
    }
}

void freeSceneBuffers(gdb200_scene *s)
{
    cudaFree(s->film); cudaFree(s->dev64); cudaFree(s->dev32); cudaFree(s->sd); cudaFree(s->si);
    cudaFree(s->liveList); cudaFree(s->liveCount); cudaFree(s->genList); cudaFree(s->genCount); cudaFree(s->counters); cudaFree(s->dScene); s->dScene = nullptr;
    s->film = s->dev64 = s->sd = nullptr; s->dev32 = nullptr; s->si = nullptr;
    s->liveList = s->liveCount = s->genList = s->genCount = nullptr; s->counters = nullptr; s->slotCapacity = 0;
}

int uploadScene(gdb200_scene *s)
{
    GDB_CUDA(cudaMemcpyToSymbol(c_scene, &s->host, sizeof(DScene)));
    GDB_CUDA(cudaMemcpyToSymbol(c_bounds, s->bounds, sizeof(s->bounds)));
    if (!s->dScene) GDB_CUDA(cudaMalloc(&s->dScene, sizeof(DScene)));
    GDB_CUDA(cudaMemcpy(s->dScene, &s->host, sizeof(DScene), cudaMemcpyHostToDevice));
    GDB_CUDA(cudaMemcpyToSymbol(c_sceneG, &s->dScene, sizeof(s->dScene)));
    return GDB200_OK;
}

int developAndCopy(gdb200_scene *s, gdb200_buffers *out)
{
    const int n = s->width * s->height;
    gpt_develop_kernel<<<(5 * n + 255) / 256, 256>>>(s->film, n, s->dev64, s->dev32);
    GDB_CUDA(cudaGetLastError());
    if (out) {
        double *dst[5] = {out->preview_final, out->throughput, out->dx, out->dy, out->direct};
        for (int b = 0; b < 5; b++)
            if (dst[b]) GDB_CUDA(cudaMemcpy(dst[b], s->dev64 + (size_t)b * n * 3, sizeof(double) * n * 3, cudaMemcpyDeviceToHost));
    }
    GDB_CUDA(cudaDeviceSynchronize());
    return GDB200_OK;
}

}  // namespace

extern "C" {

int gdb200_scene_create(const gdb200_scene_desc *desc, gdb200_scene **out)
{
    if (!desc || !out) return set_error(GDB200_ERR_ARGUMENT, "desc/out_scene is NULL");
    *out = nullptr;
    DeviceInfo di;
    if (int rc = device_info(&di)) return rc;
    gdb200_scene *s = new gdb200_scene;
    s->device = di.device;
    if (int rc = flattenScene(desc, s)) { delete s; return rc; }
    const size_t n = (size_t)s->width * s->height;
    cudaError_t e = cudaMalloc(&s->film, sizeof(double) * 5 * n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&s->dev64, sizeof(double) * 5 * n * 3);
    if (e == cudaSuccess) e = cudaMalloc(&s->dev32, sizeof(float) * 5 * n * 3);
    if (e == cudaSuccess) e = cudaMalloc(&s->counters, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess) e = cudaMemset(s->film, 0, sizeof(double) * 5 * n * 4);
    if (e != cudaSuccess) { freeSceneBuffers(s); delete s; return set_error(GDB200_ERR_CUDA, "scene allocation failed: %s", cudaGetErrorString(e)); }
    *out = s;
    return GDB200_OK;
}

void gdb200_scene_destroy(gdb200_scene *s)
{
    if (!s) return;
    freeSceneBuffers(s);
    delete s;
}

void gdb200_cancel(gdb200_scene *s) { if (s) s->cancel = 1; }

int gdb200_gpt_render(gdb200_scene *s, const gdb200_gpt_params *p, gdb200_buffers *out, gdb200_stats *stats)
{
    if (!s || !p) return set_error(GDB200_ERR_ARGUMENT, "scene/params is NULL");
    // Parameter validation of gpt.cpp:1194-1210 / integrator.cpp:190-225.
    if (p->max_depth <= 0 && p->max_depth != -1) return set_error(GDB200_ERR_ARGUMENT, "'maxDepth' must be set to -1 (infinite) or a value greater than zero!");
    if (p->rr_depth <= 0) return set_error(GDB200_ERR_ARGUMENT, "'rrDepth' must be set to a value greater than zero!");
    if (p->spp <= 0) return set_error(GDB200_ERR_ARGUMENT, "sampleCount must be positive");
    const bool banded = p->band_count > 1;
    const bool all = banded || (p->y_begin == 0 && p->y_end == 0);
    const int y0 = all ? 0 : p->y_begin, y1 = all ? s->height : p->y_end;
    if (y0 < 0 || y1 > s->height || y0 >= y1) return set_error(GDB200_ERR_ARGUMENT, "invalid row range [%d,%d)", y0, y1);
    int ownedRows = y1 - y0;
    if (banded) {
        if (p->band_rows <= 0 || p->band_index < 0 || p->band_index >= p->band_count)
            return set_error(GDB200_ERR_ARGUMENT, "invalid band sharding (%d rows, index %d of %d)", p->band_rows, p->band_index, p->band_count);
        ownedRows = 0;
        for (int y = 0; y < s->height; y++) ownedRows += ((y / p->band_rows) % p->band_count) == p->band_index;
        if (ownedRows == 0) return set_error(GDB200_ERR_ARGUMENT, "band sharding leaves rank %d without rows", p->band_index);
    }
    GDB_CUDA(cudaSetDevice(s->device));
    const int nSlots = s->width * ownedRows;
    if (nSlots > s->slotCapacity) {
        cudaFree(s->sd); cudaFree(s->si); cudaFree(s->liveList); cudaFree(s->liveCount); cudaFree(s->genList); cudaFree(s->genCount);
        s->sd = nullptr; s->si = nullptr; s->liveList = s->liveCount = s->genList = s->genCount = nullptr;
        GDB_CUDA(cudaMalloc(&s->sd, sizeof(double) * 4 * kRecords * (size_t)nSlots));
        static_assert(IF_COUNT <= 16, "int fields must fit the 16-int slot line");
        GDB_CUDA(cudaMalloc(&s->si, sizeof(int) * 16 * (size_t)nSlots));
        GDB_CUDA(cudaMalloc(&s->liveList, sizeof(int) * 2 * (size_t)kBuckets * nSlots));
        GDB_CUDA(cudaMalloc(&s->liveCount, sizeof(int) * 2 * kBuckets));
        GDB_CUDA(cudaMalloc(&s->genList, sizeof(int) * 2 * (size_t)nSlots));
        GDB_CUDA(cudaMalloc(&s->genCount, sizeof(int) * 2));
        s->slotCapacity = nSlots;
    }
    classifyMaterials(s, p->shift_threshold);

    std::lock_guard<std::mutex> lock(g_constMutex);
    if (int rc = uploadScene(s)) return rc;
    GDB_CUDA(cudaMemset(s->film, 0, sizeof(double) * 5 * (size_t)s->width * s->height * 4));
    GDB_CUDA(cudaMemset(s->counters, 0, sizeof(unsigned long long) * 8));

    GptArgs a;
    memset(&a, 0, sizeof(a));
    a.sd = s->sd; a.si = s->si; a.nSlots = nSlots; a.width = s->width; a.height = s->height; a.yBegin = y0;
    a.spp = p->spp; a.seed = p->seed; a.skipPreview = p->skip_preview != 0;
    a.bandRows = banded ? p->band_rows : 0; a.bandCount = banded ? p->band_count : 0; a.bandIndex = banded ? p->band_index : 0;
    a.cfg.maxDepth = p->max_depth; a.cfg.minDepth = 1; a.cfg.rrDepth = p->rr_depth;         // gpt.cpp:1368-1371
    a.cfg.strictNormals = p->strict_normals; a.cfg.shiftThreshold = p->shift_threshold;
    a.film = s->film; a.liveList = s->liveList; a.liveCount = s->liveCount; a.genList = s->genList; a.genCount = s->genCount;
    a.counters = s->counters;

    cudaEvent_t e0, e1;
    GDB_CUDA(cudaEventCreate(&e0)); GDB_CUDA(cudaEventCreate(&e1));
    GDB_CUDA(cudaEventRecord(e0));
    s->cancel = 0;
    gpt_init_kernel<<<(nSlots + 255) / 256, 256>>>(a);
    int launches = 1;
    const int genBlocks = (nSlots + kGenThreads - 1) / kGenThreads;
    const int bounceBlocks = (nSlots + 32 * kBuckets + kBounceThreads - 1) / kBounceThreads;
    unsigned long long hostCounters[6] = {0, 0, 0, 0, 0, 0};
    int parity = 0;
    std::vector<cudaEvent_t> marks;   // per-kernel timing (only when the caller asked for stats)
    unsigned long long tailThreshold = (unsigned long long)std::max(nSlots / 16, std::min(nSlots, 16384));
    // Default: both stages of a bounce in ONE pass over the state (least HBM traffic, fewest launches).
    // GDB200_SPLIT_PHASES=1 runs them as two kernels (smaller hot code per kernel) for A/B measurements.
    const bool fused = getenv("GDB200_SPLIT_PHASES") == nullptr;
    const long long maxSteps = (long long)p->spp * 4096 + 65536;     // safety net: never spin forever
    for (long long step = 0;; step++) {
        if (step > maxSteps) { cudaEventDestroy(e0); cudaEventDestroy(e1); return set_error(GDB200_ERR_CUDA, "wavefront did not drain after %lld steps", step); }
        auto mark = [&]() { if (stats) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev); marks.push_back(ev); } };
        mark();
        gpt_generate_kernel<<<genBlocks, kGenThreads>>>(a, parity);
        mark();
        gpt_compact_kernel<<<(nSlots + 255) / 256, 256>>>(a, parity);
        mark();
        if (fused) { gpt_bounce_kernel<2><<<bounceBlocks, kBounceThreads>>>(a, parity); launches += 3; }
        else {
            gpt_bounce_kernel<0><<<bounceBlocks, kBounceThreads>>>(a, parity);
            gpt_bounce_kernel<1><<<bounceBlocks, kBounceThreads>>>(a, parity);
            launches += 4;
        }
        mark();
        parity ^= 1;
        if ((step & 15) == 15) {
            GDB_CUDA(cudaMemcpy(hostCounters, s->counters, sizeof(hostCounters), cudaMemcpyDeviceToHost));
            if (hostCounters[0] >= (unsigned long long)nSlots) break;
            if (s->cancel) { cudaEventDestroy(e0); cudaEventDestroy(e1); return set_error(GDB200_ERR_CANCELLED, "render cancelled"); }
            // tail: few pixel streams left => finish them in one launch instead of 4 launches per bounce
            const unsigned long long remaining = (unsigned long long)nSlots - hostCounters[0];
            if (remaining <= tailThreshold) {
                gpt_tail_kernel<<<(nSlots + kBounceThreads - 1) / kBounceThreads, kBounceThreads>>>(a);
                launches++;
                break;
            }
        }
    }
    GDB_CUDA(cudaEventRecord(e1));
    GDB_CUDA(cudaEventSynchronize(e1));
    GDB_CUDA(cudaGetLastError());
    float ms = 0.f;
    GDB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    GDB_CUDA(cudaMemcpy(hostCounters, s->counters, sizeof(hostCounters), cudaMemcpyDeviceToHost));
    if (int rc = developAndCopy(s, out)) return rc;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->device_ms = ms; stats->launches = launches + 1;
        stats->samples = (double)hostCounters[3]; stats->rays = (double)hostCounters[1]; stats->path_vertices = (double)hostCounters[2];
        stats->state_bytes = (double)hostCounters[4]; stats->path_bounces = (double)hostCounters[5];
        for (size_t i = 0; i + 3 < marks.size(); i += 4) {
            float g = 0, c = 0, b = 0;
            cudaEventElapsedTime(&g, marks[i], marks[i + 1]); cudaEventElapsedTime(&c, marks[i + 1], marks[i + 2]); cudaEventElapsedTime(&b, marks[i + 2], marks[i + 3]);
            stats->generate_ms += g; stats->compact_ms += c; stats->bounce_ms += b; stats->bounce_launches++;
        }
    }
    for (cudaEvent_t ev : marks) cudaEventDestroy(ev);
    return GDB200_OK;
}

int gdb200_debug_check_culling(gdb200_scene *s, int n_rays, unsigned long long seed, unsigned long long *out_mismatches, unsigned long long *out_hits)
{
    if (!s || !out_mismatches || n_rays <= 0) return set_error(GDB200_ERR_ARGUMENT, "scene/out is NULL");
    GDB_CUDA(cudaSetDevice(s->device));
    classifyMaterials(s, 0.001);
    std::lock_guard<std::mutex> lock(g_constMutex);
    if (int rc = uploadScene(s)) return rc;
    GDB_CUDA(cudaMemset(s->counters, 0, sizeof(unsigned long long) * 8));
    gpt_check_culling_kernel<<<(n_rays + 127) / 128, 128>>>(seed, n_rays, s->counters);
    GDB_CUDA(cudaGetLastError());
    unsigned long long h[2];
    GDB_CUDA(cudaMemcpy(h, s->counters, sizeof(h), cudaMemcpyDeviceToHost));
    *out_mismatches = h[0];
    if (out_hits) *out_hits = h[1];
    return GDB200_OK;
}

int gdb200_gpt_solver_inputs(gdb200_scene *s, const float **d_dx, const float **d_dy, const float **d_thr, const float **d_direct)
{
    if (!s) return set_error(GDB200_ERR_ARGUMENT, "scene is NULL");
    const size_t n3 = (size_t)s->width * s->height * 3;
    if (d_thr) *d_thr = s->dev32 + BUF_THROUGHPUT * n3;
    if (d_dx) *d_dx = s->dev32 + BUF_DX * n3;
    if (d_dy) *d_dy = s->dev32 + BUF_DY * n3;
    if (d_direct) *d_direct = s->dev32 + BUF_DIRECT * n3;
    return GDB200_OK;
}

int gdb200_gpt_accumulators(gdb200_scene *s, double **d_accum, size_t *bytes)
{
    if (!s || !d_accum) return set_error(GDB200_ERR_ARGUMENT, "scene/d_accum is NULL");
    *d_accum = s->film;
    if (bytes) *bytes = sizeof(double) * 5 * (size_t)s->width * s->height * 4;
    return GDB200_OK;
}

int gdb200_gpt_develop(gdb200_scene *s, gdb200_buffers *out)
{
    if (!s) return set_error(GDB200_ERR_ARGUMENT, "scene is NULL");
    GDB_CUDA(cudaSetDevice(s->device));
    return developAndCopy(s, out);
}

}  // extern "C"
