// gdb200 G-PT tracer — the staged wavefront: shading stages that never trace, and one small kernel that only traces.
//
// gpt_bounce_kernel (gpt_kernels.cuh) runs a whole bounce of gpt.cpp:537-1175 in one thread: ~10 ray casts inlined
// between the BSDF / MIS arithmetic, 255 registers, a 1.5 KB stack frame, 8 warps per SM, 21 of 32 lanes busy.  Here the
// bounce is cut at its ray casts.  Every ray a path needs is appended to a dense queue (64 B: origin, direction, interval)
// and answered by gpt_cast_kernel — the intersection search alone, at several times the occupancy and with every lane of a
// warp doing the same thing — and the path continues in the next stage with the answers:
//
//   generate  splat the ended path (gpt.cpp:1319-1352), start the next sample: 5 camera rays          -> WAIT_PRIMARY
//   primary   primary hits of the base path and its 4 offsets, very-direct emission (gpt.cpp:468-534) -> LIVE | FINISHED
//   prepare   light sample of the base vertex and of every unconnected offset vertex (shadow rays),
//             BSDF sample of the base path (extension ray)                    (gpt.cpp:565-575,737-767) -> WAIT_SHADE
//   shade     everything else of the bounce.  An offset path whose shift needs one more ray (reconnection:
//             visibility; half-vector: its own extension ray) is computed as far as the ray allows
//             and parked                                                                               -> LIVE | FINISHED | WAIT_RESOLVE
//   resolve   parked offsets: pick the outcome the ray decided, finish the BSDF-stage accumulation     -> LIVE | FINISHED
//
// One tick = compact(A) -> primary, shade, resolve -> compact(B) -> prepare, generate -> cast; a slot advances one stage
// per tick, slots of all stages share the casts.  Per slot the arithmetic and its order are those of bounceBody — light and
// BSDF samples are re-derived from the path's sampler position where two stages need them — so the film is the same.
#pragma once
#include "gpt_kernels.cuh"

namespace gdb200 {

// The staged stages write records whole (stvf / stvw).  The sample's film position rides in BR_VD.w (x) and BR_RAD.w (y).
enum StagedRec { XR_HIT0 = BR_X0 /* answer to the slot's nearest-hit ray 0: t u v primitive */, XR_OCCLUDED = BR_X1 /* answers to its any-hit rays 0..4, as ints */,
                 XR_PD_BW = BR_COUNT + OR_X2 /* BSDF-stage weights of offsets 0..3 (spare record of offset 0's group) */,
                 XR_BS_WO = kRecords /* w: pdf */, XR_BS_WEIGHT /* w: eta */,
                 XR_PD_MAIN /* base contribution of the BSDF stage */, XR_PD_W /* x: weight if a reconnection fails, y: if a half-vector shift fails */,
                 kRecordsStaged };
GDB_D constexpr int hitRec(int id) { return id == 0 ? (int)XR_HIT0 : BR_COUNT + (id - 1) * OR_COUNT + OR_X1; }   // answer to nearest-hit ray id (1..4: in offset id-1's group)
GDB_D constexpr int pendRec(int i) { return BR_COUNT + i * OR_COUNT + OR_X0; }   // parked offset i: xyz + weight of the successful outcome
static_assert(kRecordsStaged <= kRecPitch, "the slot block holds every record");

// Stage queues.  A: built after the casts (continuations); B: built after those ran (slots that need new rays).
constexpr int QA_PRIMARY = 0, QA_SHADE0 = 1, QA_RESOLVE = QA_SHADE0 + kBuckets, kQA = QA_RESOLVE + 1;
constexpr int QB_PREPARE0 = kQA, QB_GEN = QB_PREPARE0 + kBsdfTypes, kStageBuckets = QB_GEN + 1;
constexpr int kStageThreads = 128;
enum { PEND_NONE = 0, PEND_RECONNECT = 1, PEND_HALFVECTOR = 2 };
// IF_PEND: bits 0-7 pending kind of offsets 0..3 (2 bits each), 8-11 offsets whose BSDF-stage weight is set, 12 the stage is
// accumulated (minDepth), 13 the base path left the scene, 14 the base path ended, 15 base vertex type, 16-19 offset vertex types
GDB_D int pendKind(unsigned p, int i) { return (p >> (2 * i)) & 3u; }

// Where a slot goes next.  The compaction kernels read only this dense array (4 B per slot) instead of the slots' state.
GDB_D int shiftStage(unsigned flags)
{
    // stage 0: some offset path is still unconnected (its own light sample, a reconnection or half-vector ray ahead),
    // stage 1: some offset was connected on the previous bounce (extra BSDF evaluations), stage 2: all offsets ride along
    // with the base path or are dead.  Lanes of one warp then run the same branches.
    int stage = 2;
    for (int i = 0; i < 4; i++) {
        if (!flagAlive(flags, i)) continue;
        const int c = flagConn(flags, i);
        if (c == RAY_NOT_CONNECTED) stage = 0; else if (c == RAY_RECENTLY_CONNECTED && stage == 2) stage = 1;
    }
    return stage;
}
GDB_D void setStatus(const GptArgs &a, int slot, int status, int queue) { SI(a, IF_STATUS, slot) = status; a.qKey[slot] = queue; }
GDB_D int prepareQueue(int material) { return QB_PREPARE0 + c_sceneG->materials[material].type; }

// Statistics of one thread over its persistent loop; flushed with one warp-aggregated atomic per counter when the kernel ends.
struct Tally { unsigned done = 0, rays = 0, vertices = 0, samples = 0, bytes = 0, bounces = 0; };
GDB_D void flushTally(const GptArgs &a, const Tally &t)
{
    countWarp(&a.counters[0], t.done); countWarp(&a.counters[1], t.rays); countWarp(&a.counters[2], t.vertices);
    countWarp(&a.counters[3], t.samples); countWarp(&a.counters[4], t.bytes); countWarp(&a.counters[5], t.bounces);
}

// Ray queues.  A stage reserves the entries it may need for a slot with ONE warp-aggregated atomic (an atomic with a return
// value is a ~1 us round trip: five or six of them in a row per slot were a visible part of prepare and generate), writes the
// rays it really casts and turns the rest into holes, which the cast kernels skip.  A slot's rays are adjacent in the queue.
template <int WHICH>
GDB_D int reserveRays(const GptArgs &a, int count)       // count <= 7 per lane
{
#ifdef GDB200_EMU
    return count ? atomAdd(&a.rayCount[WHICH], count) : 0;
#else
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    const unsigned lt = (1u << lane) - 1;
    int prefix = 0, total = 0;
#pragma unroll
    for (int b = 0; b < 3; b++) { const unsigned v = __ballot_sync(m, (count >> b) & 1); prefix += __popc(v & lt) << b; total += __popc(v) << b; }
    int base = 0;
    if (lane == leader && total) base = atomAdd(&a.rayCount[WHICH], total);
    return __shfl_sync(m, base, leader) + prefix;
#endif
}
template <int WHICH>
GDB_D void putRay(const GptArgs &a, int idx, int slot, int id, const Ray &ray)
{
    if (idx >= a.rayCapacity) { redAdd(&a.counters[7], 1ULL); return; }          // reported by the host as an error
    double2 *p = reinterpret_cast<double2 *>(a.rays[WHICH] + ((size_t)idx << 3));
    p[0] = make_double2(ray.o.x, ray.o.y); p[1] = make_double2(ray.o.z, ray.d.x);
    p[2] = make_double2(ray.d.y, ray.d.z); p[3] = make_double2(ray.mint, ray.maxt);
    a.rayOwner[WHICH][idx] = slot * 8 + id;
}
template <int WHICH>
GDB_D void putHoles(const GptArgs &a, int from, int to) { for (int i = from; i < to && i < a.rayCapacity; i++) a.rayOwner[WHICH][i] = -1; }
GDB_D Hit loadHit(const GptArgs &a, int id, int slot)
{
    const double2 *p = reinterpret_cast<const double2 *>(REC(a, hitRec(id), slot));
    const double2 lo = p[0], hi = p[1];
    Hit h; h.t = lo.x; h.u = lo.y; h.v = hi.x;
    const int prim = (int)hi.y; h.kind = prim >> 28; h.index = prim & 0x0fffffff;
    return h;
}
GDB_D bool loadOccluded(const GptArgs &a, int id, int slot) { return reinterpret_cast<const int *>(REC(a, XR_OCCLUDED, slot))[id] != 0; }

// ------------------------------------------------------------------ cast: the only kernels that intersect
// The ray queue is dense, so a CTA's batch of rays is one contiguous block of memory: it is brought into shared memory by the
// TMA unit as a 1-D bulk copy (cp.async.bulk, completion on an mbarrier), double buffered -- thread 0 issues the copy of the
// CTA's next batch before the threads start on the current one, so the loads of batch k+1 run under the intersection
// arithmetic of batch k and cost the warps no load instructions and no registers.
constexpr int kCastThreads = 128;
#ifndef GDB200_EMU
GDB_D unsigned smemAddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
GDB_D void mbarInit(unsigned long long *bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory"); }
GDB_D void mbarExpectTx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
GDB_D void mbarWait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n .reg .pred p;\n WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra WAIT;\n DONE:\n}" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
GDB_D void bulkLoad(void *dstShared, const void *srcGlobal, unsigned bytes, unsigned long long *bar)       // bytes: multiple of 16
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smemAddr(dstShared)), "l"(__cvta_generic_to_global(srcGlobal)), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
#endif

// One queue entry: the search, and its answer into the owner slot's block.
template <bool Any>
GDB_D void castOne(const GptArgs &a, int owner, const double *q)
{
    if (owner < 0) return;                                                           // reserved, not cast
    const int slot = owner >> 3, id = owner & 7;
    Ray ray; ray.o = mk(q[0], q[1], q[2]); ray.d = mk(q[3], q[4], q[5]); ray.mint = q[6]; ray.maxt = q[7];
    if (Any) reinterpret_cast<int *>(REC(a, XR_OCCLUDED, slot))[id] = rayOccludedImpl(ray) ? 1 : 0;
    else {
        Hit h; castClosest(ray, h);
        double2 *o = reinterpret_cast<double2 *>(REC(a, hitRec(id), slot));
        o[0] = make_double2(h.t, h.u); o[1] = make_double2(h.v, (Float)((h.kind << 28) | h.index));
    }
}

template <bool Any>
__global__ void __launch_bounds__(kCastThreads) gpt_cast_kernel(const GptArgs a)
{
    constexpr int Q = Any ? 1 : 0;
    stampPhase(a);
    const int n = min(a.rayCount[Q], a.rayCapacity);
#ifdef GDB200_EMU
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) castOne<Any>(a, a.rayOwner[Q][r], a.rays[Q] + ((size_t)r << 3));
#else
    if (c_scene.nBvhNodes > 0) {
        // Scenes behind the BVH: a ray's traversal length varies by an order of magnitude, and the barrier that hands a
        // staged batch back to the TMA unit makes the CTA's 128 threads wait for its slowest ray (ncu: 3 of 13 stall cycles
        // per issue on the barrier).  Plain per-thread loads here, no barrier.
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) castOne<Any>(a, a.rayOwner[Q][r], a.rays[Q] + ((size_t)r << 3));
        return;
    }
    const int nBatches = (n + kCastThreads - 1) / kCastThreads;
    __shared__ __align__(128) double s_rays[2][kCastThreads * 8];
    __shared__ __align__(16) int s_owner[2][kCastThreads];
    __shared__ __align__(8) unsigned long long s_bar[2];
    if (threadIdx.x == 0) {
        mbarInit(&s_bar[0], 1); mbarInit(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto fetch = [&](int batch, int stage) {      // thread 0: the batch's rays (64 B each) and owners (4 B each, the copy padded to 16 B)
        const int cnt = min(kCastThreads, n - batch * kCastThreads);
        const unsigned rayBytes = (unsigned)cnt * 64u, ownerBytes = ((unsigned)cnt * 4u + 15u) & ~15u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the stage was last read through the generic proxy
        mbarExpectTx(&s_bar[stage], rayBytes + ownerBytes);
        bulkLoad(&s_rays[stage][0], a.rays[Q] + ((size_t)batch * kCastThreads << 3), rayBytes, &s_bar[stage]);
        bulkLoad(&s_owner[stage][0], a.rayOwner[Q] + (size_t)batch * kCastThreads, ownerBytes, &s_bar[stage]);
    };
    unsigned parity[2] = {0u, 0u};
    int stage = 0;
    if (threadIdx.x == 0 && (int)blockIdx.x < nBatches) fetch(blockIdx.x, 0);
    for (int batch = blockIdx.x; batch < nBatches; batch += gridDim.x, stage ^= 1) {
        if (threadIdx.x == 0 && batch + (int)gridDim.x < nBatches) fetch(batch + gridDim.x, stage ^ 1);
        mbarWait(&s_bar[stage], parity[stage]); parity[stage] ^= 1u;
        const int r = batch * kCastThreads + threadIdx.x;
        castOne<Any>(a, r < n ? s_owner[stage][threadIdx.x] : -1, &s_rays[stage][threadIdx.x << 3]);
        __syncthreads();                          // every thread is done with the stage before the TMA unit refills it
    }
#endif
}

// ------------------------------------------------------------------ generate
GDB_D void stagedGenerateBody(const GptArgs &a, int slot, Tally &tally)
{
    if (SI(a, IF_STATUS, slot) == ST_FINISHED) {
        Spec rad[4], grad[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { const int o = BR_COUNT + i * OR_COUNT; rad[i] = ldvL2(a, o + OR_RAD, slot); grad[i] = ldvL2(a, o + OR_GRAD, slot); }
        Spec vd, C; Float spx, spy;
        ldvw(a, BR_VD, slot, vd, spx); ldvw(a, BR_RAD, slot, C, spy);
        splatSample(a, spx, spy, vd, C, rad, grad);
    }
    int stream = SI(a, IF_STREAM, slot);
    StreamInfo si = streamInfo(a, stream);
    Sampler smp; smp.key = si.key; smp.n = (uint32_t)SI(a, IF_RNGN, slot);          // Sampler::generate, gpt.cpp:1250-1251
    int j = SI(a, IF_SAMPLE, slot);
    unsigned samples = 0;
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
        if (c_scene.apertureRadius > 0) { apx = smp.next1D(); apy = smp.next1D(); }  // gpt.cpp:1263-1265
        const int q = reserveRays<0>(a, 5);
        Ray ray;
        sampleCameraRay(spx, spy, apx, apy, ray);                                    // gpt.cpp:402
        putRay<0>(a, q, slot, 0, ray);
        const Float shiftX[4] = {1, 0, -1, 0}, shiftY[4] = {0, 1, 0, -1};            // gpt.cpp:410-415
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            sampleCameraRay(spx + shiftX[i], spy + shiftY[i], apx, apy, ray);        // gpt.cpp:418
            putRay<0>(a, q + 1 + i, slot, 1 + i, ray);
        }
        stvw(a, BR_VD, slot, splat(0), spx); stvw(a, BR_RAD, slot, splat(0), spy);
        status = ST_WAIT_PRIMARY;
        break;
    }
    setStatus(a, slot, status, status == ST_DONE ? -1 : QA_PRIMARY);
    SI(a, IF_SAMPLE, slot) = j; SI(a, IF_RNGN, slot) = (int)smp.n; SI(a, IF_STREAM, slot) = stream;
    tally.done += status == ST_DONE ? 1u : 0u; tally.rays += 5u * samples; tally.samples += samples;
}

// ------------------------------------------------------------------ primary: gpt.cpp:468-534 with the camera rays' answers
GDB_D void stagedPrimaryBody(const GptArgs &a, int slot, Tally &tally)
{
    const Float spx = W(a, BR_VD, slot), spy = W(a, BR_RAD, slot);
    Float apx = 0.5, apy = 0.5;
    if (c_scene.apertureRadius > 0) {                                                // the two values drawn last by generate
        Sampler smp; smp.key = streamInfo(a, SI(a, IF_STREAM, slot)).key; smp.n = (uint32_t)SI(a, IF_RNGN, slot) - 2u;
        apx = smp.next1D(); apy = smp.next1D();
    }
    Ray ray; Its mits;
    sampleCameraRay(spx, spy, apx, apy, ray);
    const Hit mh = loadHit(a, 0, slot);
    const bool mainValid = mh.t != CUDART_INF;                                       // gpt.cpp:472
    if (mainValid) fillIts(ray, mh, mits);
    Spec veryDirect = splat(0);
    unsigned flags = 0;
    bool early = !mainValid;                                                         // gpt.cpp:482-492
    if (!mainValid && c_scene.env.present) veryDirect = veryDirect + splat(1.0) * envEval(ray.d);
    if (mainValid && mits.emitter >= 0) veryDirect = veryDirect + splat(1.0) * emittedLe(mits, -ray.d);
    if (mainValid && a.cfg.strictNormals && dot(ray.d, mits.geoN) * mits.wi.z >= 0) early = true;
    const Float shiftX[4] = {1, 0, -1, 0}, shiftY[4] = {0, 1, 0, -1};
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        Ray sray; Its sits;
        sampleCameraRay(spx + shiftX[i], spy + shiftY[i], apx, apy, sray);
        const Hit sh = loadHit(a, 1 + i, slot);
        bool alive = sh.t != CUDART_INF;                                             // gpt.cpp:476-480, 508-513
        if (alive) fillIts(sray, sh, sits);
        if (alive && a.cfg.strictNormals && dot(sray.d, sits.geoN) * sits.wi.z >= 0) alive = false;
        flags |= packFlag(i, alive, RAY_NOT_CONNECTED);
        const int o = BR_COUNT + i * OR_COUNT;
        stvw(a, o + OR_THR, slot, splat(1.0), 1.0);
        stvf(a, o + OR_RAD, slot, splat(0)); stvf(a, o + OR_GRAD, slot, splat(0));
        if (!early && alive) storeOffItsFull(a, slot, i, sits);
    }
    stvw(a, BR_RAD, slot, splat(0), spy); stvw(a, BR_VD, slot, veryDirect, spx);
    if (early || !(1 < a.cfg.maxDepth || a.cfg.maxDepth < 0)) {                      // bounce loop never entered (gpt.cpp:537): the
        if (!early) tally.vertices += 1u;                                            // sample is its very-direct term; generate splats it
        setStatus(a, slot, ST_FINISHED, QB_GEN);
        return;
    }
    storeBaseItsFull(a, slot, mits, 1.0);                                            // eta = 1
    stvw(a, BR_RAYD, slot, ray.d, 1.0);                                              // pdf = 1
    stvf(a, BR_THR, slot, splat(1.0));
    SI(a, IF_DEPTH, slot) = 1; SI(a, IF_OFLAGS, slot) = (int)flags;
    setStatus(a, slot, ST_LIVE, prepareQueue(mits.material));
}

// ------------------------------------------------------------------ prepare: the rays a bounce starts with
// strictNormals pre-pass of the bounce (gpt.cpp:541-555); returns false when the base path ends here.
GDB_D bool strictNormalsPrepass(const GptArgs &a, int slot, const Its &mits, V3 mrayD, unsigned &flags)
{
    if (dot(mrayD, mits.geoN) * mits.wi.z >= 0) return false;
    for (int i = 0; i < 4; i++) {       // an unconnected offset's ray direction is -toWorld(wi) of its stored vertex
        if (!flagAlive(flags, i) || flagConn(flags, i) != RAY_NOT_CONNECTED) continue;
        Its sits; loadOffIts(a, slot, i, sits);
        const V3 sd = -toWorld(sits.sh, sits.wi);
        if (dot(sd, sits.geoN) * sits.wi.z >= 0) flags = setFlag(flags, i, false, flagConn(flags, i));
    }
    return true;
}
GDB_D bool emitterIsDirac(int emitter) { const int k = c_sceneG->emitters[emitter].kind; return k == EM_POINT || k == EM_SPOT; }
// an unconnected offset path draws its own light sample iff (gpt.cpp:668-672)
GDB_D bool offsetSamplesLight(const DMaterial &mainBSDF, const DMaterial &shiftedBSDF, bool atPointLight)
{
    return atPointLight || (vertexType(mainBSDF, ESmooth) == VERTEX_TYPE_DIFFUSE && vertexType(shiftedBSDF, ESmooth) == VERTEX_TYPE_DIFFUSE);
}

GDB_D void stagedPrepareBody(const GptArgs &a, int slot)
{
    const Config cfg = a.cfg;
    Its mits; loadBaseIts(a, slot, mits);
    const int depth = SI(a, IF_DEPTH, slot);
    unsigned flags = (unsigned)SI(a, IF_OFLAGS, slot);
    const int shadeQueue = QA_SHADE0 + shiftStage(flags) * kBsdfTypes + c_sceneG->materials[mits.material].type;
    Sampler smp; smp.key = streamInfo(a, SI(a, IF_STREAM, slot)).key; smp.n = (uint32_t)SI(a, IF_RNGN, slot);
    bool ended = false;
    if (cfg.strictNormals) ended = !strictNormalsPrepass(a, slot, mits, ldv(a, BR_RAYD, slot), flags);
    if (!ended) {
        const DMaterial &mainBSDF = c_sceneG->materials[mits.material];
        if ((mainBSDF.flags & ESmooth) && depth + 1 >= cfg.minDepth) {               // gpt.cpp:568
            DRec dRec; initDRec(mits, dRec);
            const Float lsx = smp.next1D(), lsy = smp.next1D();                      // gpt.cpp:572
            bool needsRay; Ray sray;
            sampleEmitterDirect(dRec, lsx, lsy, needsRay, sray);
            const V3 woL = toLocal(mits.sh, dRec.d);
            const bool atPointLight = emitterIsDirac(dRec.emitter);
            const bool neeActive = !cfg.strictNormals || dot(mits.geoN, dRec.d) * woL.z > 0;   // gpt.cpp:607
            unsigned own = 0;                                                        // offsets that draw their own light sample, gpt.cpp:659-676
            if (neeActive)
                for (int i = 0; i < 4; i++)
                    if (flagAlive(flags, i) && flagConn(flags, i) == RAY_NOT_CONNECTED &&
                        offsetSamplesLight(mainBSDF, c_sceneG->materials[SI(a, IF_OMAT0 + i, slot)], atPointLight)) own |= 1u << i;
            int q = reserveRays<1>(a, 1 + __popc(own));
            const int qEnd = q + 1 + __popc(own);
            if (needsRay) putRay<1>(a, q++, slot, 0, sray);
#pragma unroll 1
            for (int i = 0; i < 4; i++) {
                if (!(own & (1u << i))) continue;
                const int smat = SI(a, IF_OMAT0 + i, slot);
                const int o = BR_COUNT + i * OR_COUNT;
                DRec sRec; sRec.ref = ldv(a, o + OR_P, slot);
                sRec.refN = c_sceneG->materials[smat].refNFromShading ? ldv(a, o + OR_N, slot) : mk(0, 0, 0);
                sampleEmitterDirect(sRec, lsx, lsy, needsRay, sray);
                if (needsRay) putRay<1>(a, q++, slot, 1 + i, sray);
            }
            putHoles<1>(a, q, qEnd);
        }
        const Float sx = smp.next1D(), sy = smp.next1D();                            // gpt.cpp:456-457
        const Float s3 = mainBSDF.type == GDB200_BSDF_ROUGHDIELECTRIC ? smp.peek1D() : 0.0;
        BSDFSample bs;
        bsdfSample(mainBSDF, mits.wi, sx, sy, s3, bs);
        stvw(a, XR_BS_WO, slot, bs.wo, bs.pdf); stvw(a, XR_BS_WEIGHT, slot, bs.weight, bs.eta);
        SI(a, IF_BSTYPE, slot) = (int)(bs.sampledType | ((unsigned)bs.extraDraws << 8));
        const int q = reserveRays<0>(a, 1);
        bool cast = false;
        if (!(bs.pdf <= 0.0)) {                                                      // gpt.cpp:739
            const V3 mainWo = toWorld(mits.sh, bs.wo);
            if (!(cfg.strictNormals && dot(mits.geoN, mainWo) * bs.wo.z <= 0)) {     // gpt.cpp:748
                Ray mray; mray.o = mits.p; mray.d = mainWo; mray.mint = kEpsilon; mray.maxt = CUDART_INF;   // gpt.cpp:767
                putRay<0>(a, q, slot, 0, mray); cast = true;
            }
        }
        if (!cast) putHoles<0>(a, q, q + 1);
    }
    setStatus(a, slot, ST_WAIT_SHADE, shadeQueue);
}

// ------------------------------------------------------------------ shade: bounceBody without a single ray cast
// STAGE = the shift stage of the slot's queue (gpt_stage_compact_kernel): 0 some offset path is unconnected, 1 none is but
// some was connected on the previous bounce, 2 every live offset rides along with the base path.  The branches a stage
// cannot reach are compiled out, so the later (and most frequent) stages run in a fraction of stage 0's registers.
//
// The bounce is evaluated as the reference orders it — next-event estimation of the base path and of its four offsets
// (gpt.cpp:565-730), then the BSDF-sample stage of all five (gpt.cpp:737-1151) — in two passes over the offset records, so
// that nothing of the first stage but the base path's radiance sum is live during the second (the light sample, its MIS
// terms and the offset vertices of pass one would otherwise sit in registers next to everything pass two needs).

// Pass one for offset i.  Everything it needs from the base path's light sample:
struct NeeBase {
    Float lsx, lsy, bsdfPdf, distSq, oppCos, wNum, wDen, lightPdf;
    V3 woLocal, lightP, lightN;
    Spec bsdfValue, emitterRadiance, contributionAll;
    bool visible, atPointLight;
};
template <int STAGE>
GDB_D void shadeOffsetNee(const GptArgs &a, int slot, int i, unsigned flags, const Config &cfg, const DMaterial &mainBSDF, const Frame &prevSh, V3 prevP,
                          const NeeBase &n, Spec &mrad)
{
    constexpr bool kUnconnected = STAGE == 0, kRecent = STAGE <= 1;
    const int o = BR_COUNT + i * OR_COUNT;
    const bool alive = flagAlive(flags, i);
    const int conn = flagConn(flags, i);
    Spec sthr = splat(0); Float spdf = 0;
    if (alive) ldvw(a, o + OR_THR, slot, sthr, spdf);
    Spec mainContribution = splat(0), shiftedContribution = splat(0);
    Float weight = 0;
    bool shiftSuccessful = alive;
    if (shiftSuccessful) {
        if (conn == RAY_CONNECTED || !kRecent) {                                     // gpt.cpp:622-637
            const Float jacobian = 1;
            const Float den = (jacobian * spdf) * (jacobian * spdf) * ((n.lightPdf * n.lightPdf) + (n.bsdfPdf * n.bsdfPdf));
            weight = n.wNum / (kDEps + den + n.wDen);
            mainContribution = n.contributionAll;
            shiftedContribution = jacobian * sthr * (n.bsdfValue * n.emitterRadiance);
        } else if (conn == RAY_RECENTLY_CONNECTED || !kUnconnected) {                // gpt.cpp:638-658
            const V3 recentWiL = toLocal(prevSh, normalize(ldv(a, o + OR_P, slot) - prevP));   // gpt.cpp:640
            Spec shiftedBsdfValue; Float shiftedBsdfPdf;
            bsdfEvalPdf(mainBSDF, recentWiL, n.woLocal, ESolidAngle, shiftedBsdfValue, shiftedBsdfPdf);
            if (!n.visible || n.atPointLight) shiftedBsdfPdf = 0;
            const Float jacobian = 1;
            const Float den = (jacobian * spdf) * (jacobian * spdf) * ((n.lightPdf * n.lightPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
            weight = n.wNum / (kDEps + den + n.wDen);
            mainContribution = n.contributionAll;
            shiftedContribution = jacobian * sthr * (shiftedBsdfValue * n.emitterRadiance);
        } else {                                                                     // gpt.cpp:659-705
            Its sits; loadOffIts(a, slot, i, sits);
            const DMaterial &shiftedBSDF = c_sceneG->materials[sits.material];
            if (offsetSamplesLight(mainBSDF, shiftedBSDF, n.atPointLight)) {          // gpt.cpp:668-672
                DRec sRec; initDRec(sits, sRec);
                bool needsRay; Ray sray;
                Spec sv = sampleEmitterDirect(sRec, n.lsx, n.lsy, needsRay, sray);
                bool shiftedEmitterVisible = true;
                if (needsRay && loadOccluded(a, 1 + i, slot)) { shiftedEmitterVisible = false; sv = splat(0); }
                const Spec shiftedEmitterRadiance = sv * sRec.pdf;
                const Float shiftedDRecPdf = sRec.pdf;
                const Float shiftedDistanceSquared = len2(n.lightP - sits.p);
                const V3 emitterDirection = (n.lightP - sits.p) / sqrt(shiftedDistanceSquared);
                const Float shiftedOpposingCosine = -dot(n.lightN, emitterDirection);
                const V3 woL = toLocal(sits.sh, emitterDirection);
                if (cfg.strictNormals && dot(sits.geoN, emitterDirection) * woL.z < 0) {
                    shiftSuccessful = false;
                } else {
                    Spec shiftedBsdfValue; Float shiftedBsdfPdf;
                    bsdfEvalPdf(shiftedBSDF, sits.wi, woL, ESolidAngle, shiftedBsdfValue, shiftedBsdfPdf);
                    if (!shiftedEmitterVisible || n.atPointLight) shiftedBsdfPdf = 0;
                    const Float jacobian = fabs(shiftedOpposingCosine * n.distSq) / (kEpsilon + fabs(n.oppCos * shiftedDistanceSquared));   // gpt.cpp:695
                    const Float den = (jacobian * spdf) * (jacobian * spdf) * ((shiftedDRecPdf * shiftedDRecPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                    weight = n.wNum / (kDEps + den + n.wDen);
                    mainContribution = n.contributionAll;
                    shiftedContribution = jacobian * sthr * (shiftedBsdfValue * shiftedEmitterRadiance);
                }
            }   // else: weight and both contributions stay 0 (gpt.cpp:613-615)
        }
    }
    if (!shiftSuccessful) {                                                          // gpt.cpp:708-717
        weight = n.wNum / (kDEps + n.wDen);
        mainContribution = n.contributionAll;
        shiftedContribution = splat(0);
    }
    mrad = mrad + mainContribution * weight;                                         // gpt.cpp:723-726
    accumulateOffset(a, o, slot, shiftedContribution * weight, (shiftedContribution - mainContribution) * weight);
}

template <int STAGE>
GDB_D void stagedShadeBody(const GptArgs &a, int slot, Tally &tally)
{
    constexpr bool kUnconnected = STAGE == 0, kRecent = STAGE <= 1;
    const Config cfg = a.cfg;
    Its mits; loadBaseIts(a, slot, mits);
    V3 mrayD; Float mpdf;
    ldvw(a, BR_RAYD, slot, mrayD, mpdf);
    Spec mrad; Float spy;
    ldvw(a, BR_RAD, slot, mrad, spy);
    int depth = SI(a, IF_DEPTH, slot);
    unsigned flags = (unsigned)SI(a, IF_OFLAGS, slot);
    Sampler smp; smp.key = streamInfo(a, SI(a, IF_STREAM, slot)).key; smp.n = (uint32_t)SI(a, IF_RNGN, slot);
    unsigned rays = 0, pend = 0;
    bool ended = false;
    {   // algorithmic state traffic of this path-bounce (record sizes of SURVEY.md §8d: base 320 B, unconnected offset
        // 304 B, connected offset 88 B; read + write)
        unsigned bytes = 320;
        for (int i = 0; i < 4; i++) if (flagAlive(flags, i)) bytes += flagConn(flags, i) == RAY_CONNECTED ? 88 : 304;
        tally.bytes += 2 * bytes; tally.bounces += 1u;
    }
    if (cfg.strictNormals) ended = !strictNormalsPrepass(a, slot, mits, mrayD, flags);            // gpt.cpp:541-555

    Spec mainContributionAll = splat(0);
    Float bw0 = 0, bw1 = 0, bw2 = 0, bw3 = 0; unsigned bHas = 0;                     // BSDF-stage weights of the base contribution
    Float failReconnect = 0, failHalfVector = 0;
    bool addBsdfStage = false, escaped = false;
    int mainVertexType = 0;
    unsigned sampledType = 0, offVertexTypes = 0;
    if (!ended) {
        const bool lastSegment = (depth + 1 == cfg.maxDepth);                        // gpt.cpp:558
        const DMaterial &mainBSDF = c_sceneG->materials[mits.material];

        // ================ pass one: next event estimation, gpt.cpp:565-730
        if ((mainBSDF.flags & ESmooth) && depth + 1 >= cfg.minDepth) {               // gpt.cpp:568
            NeeBase n;
            const Spec mthr = ldv(a, BR_THR, slot);
            DRec dRec; initDRec(mits, dRec);
            n.lsx = smp.next1D(); n.lsy = smp.next1D();                              // gpt.cpp:572
            bool needsRay; Ray sray;
            Spec value = sampleEmitterDirect(dRec, n.lsx, n.lsy, needsRay, sray); rays++;
            n.visible = true;
            if (needsRay && loadOccluded(a, 0, slot)) { n.visible = false; value = splat(0); }   // scene.cpp:869-876
            n.emitterRadiance = value * dRec.pdf;                                    // gpt.cpp:575
            n.woLocal = toLocal(mits.sh, dRec.d);
            bsdfEvalPdf(mainBSDF, mits.wi, n.woLocal, ESolidAngle, n.bsdfValue, n.bsdfPdf);   // gpt.cpp:588
            n.atPointLight = emitterIsDirac(dRec.emitter);                           // dRec.measure == EDiscrete
            if (!n.visible || n.atPointLight) n.bsdfPdf = 0;                         // gpt.cpp:592
            n.distSq = len2(mits.p - dRec.p);                                        // gpt.cpp:595-596
            n.oppCos = dot(dRec.n, (mits.p - dRec.p)) / sqrt(n.distSq);
            n.wNum = mpdf * dRec.pdf;                                                // gpt.cpp:599-600
            n.wDen = (mpdf * mpdf) * ((dRec.pdf * dRec.pdf) + (n.bsdfPdf * n.bsdfPdf));
            n.lightP = dRec.p; n.lightN = dRec.n; n.lightPdf = dRec.pdf;
            const bool neeActive = !cfg.strictNormals || dot(mits.geoN, dRec.d) * n.woLocal.z > 0;   // gpt.cpp:607
            n.contributionAll = mthr * (n.bsdfValue * n.emitterRadiance);
            if (neeActive) {
                // rays the reference casts for the offsets' own light samples (statistics only)
                if (kUnconnected)
                    for (int i = 0; i < 4; i++)
                        if (flagAlive(flags, i) && flagConn(flags, i) == RAY_NOT_CONNECTED &&
                            offsetSamplesLight(mainBSDF, c_sceneG->materials[SI(a, IF_OMAT0 + i, slot)], n.atPointLight)) rays++;
#pragma unroll 1
                for (int i = 0; i < 4; ++i) shadeOffsetNee<STAGE>(a, slot, i, flags, cfg, mainBSDF, mits.sh, mits.p, n, mrad);
            }
        }

        // ================ pass two: BSDF sample (drawn by prepare) + the extension ray's answer, gpt.cpp:737-1151
        const Frame prevSh = mits.sh; const V3 prevP = mits.p, prevWi = mits.wi;     // the vertex both stages shade (previousMainIts, gpt.cpp:753)
        Spec mthr = ldv(a, BR_THR, slot);
        Float meta = W(a, BR_P, slot);
        bool bsdfStage = false, mainHitEmitter = false;
        BSDFSample bs;
        smp.n += 2;                                                                  // the two values prepare drew, gpt.cpp:456-457
        {
            V3 w; Float f;
            ldvw(a, XR_BS_WO, slot, w, f); bs.wo = w; bs.pdf = f;
            ldvw(a, XR_BS_WEIGHT, slot, w, f); bs.weight = w; bs.eta = f;
            const unsigned t = (unsigned)SI(a, IF_BSTYPE, slot);
            bs.sampledType = t & 0xffu; bs.extraDraws = (int)(t >> 8);
            smp.n += (uint32_t)bs.extraDraws;
        }
        sampledType = bs.sampledType;
        Spec mainEmitterRadiance = splat(0);
        int mainNextVertexType = 0;
        Float mainLumPdf = 0, mainWeightNumerator = 0, mainWeightDenominator = 0;
        if (bs.pdf <= 0.0) ended = true;                                             // gpt.cpp:739
        else {
            const V3 mainWo = toWorld(mits.sh, bs.wo);
            if (cfg.strictNormals && dot(mits.geoN, mainWo) * bs.wo.z <= 0) ended = true;   // gpt.cpp:748
            else {
                DRec mainDRec; initDRec(mits, mainDRec);                             // gpt.cpp:759
                mainVertexType = vertexType(mainBSDF, bs.sampledType);               // gpt.cpp:764
                Ray mray; mray.o = mits.p; mray.d = mainWo; mray.mint = kEpsilon; mray.maxt = CUDART_INF;   // gpt.cpp:767
                rays++;
                const Hit mh = loadHit(a, 0, slot);
                if (mh.t != CUDART_INF) {
                    fillIts(mray, mh, mits);
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
        addBsdfStage = bsdfStage && depth + 1 >= cfg.minDepth;                       // gpt.cpp:1140
        failReconnect = mainWeightNumerator / (kDEps + mainWeightDenominator);       // gpt.cpp:1131-1136
        failHalfVector = (Float)1 / mpdf;                                            // gpt.cpp:1113-1125
        // The new base vertex is final from here on: it goes to the state now, so that only what the offsets need of it
        // (position, normals, emitter) stays in registers through their loop.  (A path that ends below never reads it.)
        if (bsdfStage && !escaped) storeBaseItsFull(a, slot, mits, meta);
        // Ray-queue entries for the offsets that may need a ray of their own this bounce: a visibility ray for a reconnection
        // (any-hit queue), an extension ray for a half-vector shift (nearest-hit queue).  Reserved at once, see reserveRays.
        int qAny = 0, qAnyEnd = 0, qNear = 0, qNearEnd = 0;
        if (kUnconnected && bsdfStage) {
            int nAny = 0, nNear = 0;
            for (int i = 0; i < 4; i++) {
                if (!flagAlive(flags, i) || flagConn(flags, i) != RAY_NOT_CONNECTED) continue;
                const DMaterial &sb = c_sceneG->materials[SI(a, IF_OMAT0 + i, slot)];
                if (mainVertexType == VERTEX_TYPE_DIFFUSE && mainNextVertexType == VERTEX_TYPE_DIFFUSE && vertexType(sb, bs.sampledType) == VERTEX_TYPE_DIFFUSE) {
                    if (!lastSegment || mainHitEmitter) nAny++;
                } else if (((bs.sampledType & EDelta) && (sb.flags & EDelta)) || ((bs.sampledType & ESmooth) && (sb.flags & ESmooth))) nNear++;
            }
            qAny = reserveRays<1>(a, nAny); qAnyEnd = qAny + nAny;
            qNear = reserveRays<0>(a, nNear); qNearEnd = qNear + nNear;
        }

        if (bsdfStage)
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {                                           // ---- BSDF-sample stage of offset i, gpt.cpp:830-1151
            const int o = BR_COUNT + i * OR_COUNT;
            bool alive = flagAlive(flags, i);
            int conn = flagConn(flags, i);
            Spec sthr = splat(0); Float spdf = 0;
            if (alive) ldvw(a, o + OR_THR, slot, sthr, spdf);
            Spec mainContribution = splat(0), shiftedContribution = splat(0);
            Float weight = 0;
            bool postponedShiftEnd = false;
            int parked = PEND_NONE;
            if (alive) {
                const Float shiftedPreviousPdf = spdf;
                if (conn == RAY_CONNECTED || !kRecent) {                             // gpt.cpp:844-861
                    sthr = sthr * (bs.weight * bs.pdf);
                    spdf *= mainBsdfPdf;
                    const Float den = (shiftedPreviousPdf * shiftedPreviousPdf) * ((mainLumPdf * mainLumPdf) + (mainBsdfPdf * mainBsdfPdf));
                    weight = mainWeightNumerator / (kDEps + den + mainWeightDenominator);
                    mainContribution = mainContributionAll;
                    shiftedContribution = sthr * mainEmitterRadiance;
                } else if (conn == RAY_RECENTLY_CONNECTED || !kUnconnected) {        // gpt.cpp:862-888
                    const V3 recentWiL = toLocal(prevSh, normalize(ldv(a, o + OR_P, slot) - prevP));   // gpt.cpp:864
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
                } else {                                                             // gpt.cpp:889-1126
                    Its sits; loadOffIts(a, slot, i, sits);
                    const DMaterial &shiftedBSDF = c_sceneG->materials[sits.material];
                    const int shiftedVertexType = vertexType(shiftedBSDF, bs.sampledType);
                    if (shiftedVertexType == VERTEX_TYPE_DIFFUSE) offVertexTypes |= 1u << i;
                    if (mainVertexType == VERTEX_TYPE_DIFFUSE && mainNextVertexType == VERTEX_TYPE_DIFFUSE && shiftedVertexType == VERTEX_TYPE_DIFFUSE) {
                        if (!lastSegment || mainHitEmitter) {                        // gpt.cpp:901
                            // The reconnection computed as if its visibility ray were free; resolve replaces the outcome by
                            // the failed one (offset dead, base contribution with the base-only weight) when it is not.
                            Ray vray;
                            const ShiftResult sr = escaped ? environmentShiftUnoccluded(mrayD, sits.p, vray)                   // gpt.cpp:908-915
                                                           : reconnectShiftUnoccluded(prevP, mits.p, sits.p, mits.geoN, vray); // gpt.cpp:907
                            rays++;
                            const V3 outgoingDirection = sr.wo;
                            const V3 woL = toLocal(sits.sh, outgoingDirection);
                            if (cfg.strictNormals && dot(outgoingDirection, sits.geoN) * woL.z <= 0) alive = false;   // fails whatever the ray says
                            else {
                                putRay<1>(a, qAny++, slot, 1 + i, vray);
                                parked = PEND_RECONNECT;
                                Spec shiftedBsdfValue; Float shiftedBsdfPdf;
                                bsdfEvalPdf(shiftedBSDF, sits.wi, woL, ESolidAngle, shiftedBsdfValue, shiftedBsdfPdf);   // gpt.cpp:935-936
                                sthr = sthr * (shiftedBsdfValue * sr.jacobian);
                                spdf *= shiftedBsdfPdf * sr.jacobian;
                                conn = RAY_RECENTLY_CONNECTED;
                                if (mainHitEmitter) {                                // gpt.cpp:944-985
                                    Spec shiftedEmitterRadiance; Float shiftedLumPdf;
                                    if (!escaped) {
                                        shiftedEmitterRadiance = emittedLe(mits, -outgoingDirection);
                                        DRec sd;                                     // gpt.cpp:957-964 (measure: solid angle); the base path's
                                        sd.p = mits.p; sd.n = mits.sh.n;             // record (gpt.cpp:771-776) holds its new vertex, seen from the old one
                                        sd.dist = len(mits.p - sits.p);
                                        sd.d = (mits.p - sits.p) / sd.dist;
                                        sd.ref = prevP; sd.refN = sits.sh.n; sd.emitter = mits.emitter;
                                        shiftedLumPdf = pdfEmitterDirect(sd);
                                        if (cfg.refUninitMeasure && c_sceneG->emitters[sd.emitter].kind != EM_ENV) shiftedLumPdf = 0;   // gpt_host.h setupArgs
                                    } else { shiftedEmitterRadiance = mainEmitterRadiance; shiftedLumPdf = mainLumPdf; }   // gpt.cpp:973-977
                                    const Float den = (shiftedPreviousPdf * shiftedPreviousPdf) * ((shiftedLumPdf * shiftedLumPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                                    weight = mainWeightNumerator / (kDEps + den + mainWeightDenominator);
                                    mainContribution = mainContributionAll;
                                    shiftedContribution = sthr * shiftedEmitterRadiance;
                                }   // else weight and contributions stay 0 (gpt.cpp:833-836)
                                stvw(a, pendRec(i), slot, shiftedContribution, weight);
                            }
                        }
                    } else {                                                         // half-vector shift, gpt.cpp:987-1126
                        const bool bothDelta = (bs.sampledType & EDelta) && (shiftedBSDF.flags & EDelta);
                        const bool bothSmooth = (bs.sampledType & ESmooth) && (shiftedBSDF.flags & ESmooth);
                        bool ok = bothDelta || bothSmooth;
                        if (ok) {
                            ShiftResult sr = halfVectorShift(prevWi, bs.wo, sits.wi, mainBSDF.bsdfEta, shiftedBSDF.bsdfEta);   // gpt.cpp:1006
                            if (bs.sampledType & EDelta) sr.jacobian = 1;            // gpt.cpp:1008-1011
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
                                if (spdf == 0) ok = false;                           // gpt.cpp:1033-1037
                                if (ok && cfg.strictNormals && dot(outgoingDirection, sits.geoN) * tangentSpaceOutgoingDirection.z <= 0) ok = false;
                                if (ok) {                                            // the offset's own extension ray: resolve continues at gpt.cpp:1052
                                    Ray sray; sray.o = sits.p; sray.d = outgoingDirection; sray.mint = kEpsilon; sray.maxt = CUDART_INF;   // gpt.cpp:1050
                                    putRay<0>(a, qNear++, slot, 1 + i, sray); rays++;
                                    parked = PEND_HALFVECTOR;
                                    stvw(a, pendRec(i), slot, outgoingDirection, mpdf / (spdf * spdf + mpdf * mpdf));   // weight of gpt.cpp:1107-1112
                                }
                            }
                        }
                        if (!ok) {                                                   // gpt.cpp:1113-1125
                            weight = failHalfVector;
                            mainContribution = mainContributionAll;
                            shiftedContribution = splat(0);
                            postponedShiftEnd = true;
                        }
                    }
                }
            }
            if (parked != PEND_NONE) pend |= (unsigned)parked << (2 * i);
            else {
                if (!alive) {                                                        // gpt.cpp:1131-1136
                    weight = failReconnect;
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
            }
            flags = setFlag(flags, i, alive, conn);
            if (alive) stvw(a, o + OR_THR, slot, sthr, spdf);
        }
        if (kUnconnected) { putHoles<1>(a, qAny, qAnyEnd); putHoles<0>(a, qNear, qNearEnd); }
        if (!pend) {    // base radiance: BSDF-stage terms after all NEE terms, in offset order (gpt.cpp:1142)
            if (bHas & 1u) mrad = mrad + mainContributionAll * bw0;
            if (bHas & 2u) mrad = mrad + mainContributionAll * bw1;
            if (bHas & 4u) mrad = mrad + mainContributionAll * bw2;
            if (bHas & 8u) mrad = mrad + mainContributionAll * bw3;
        }

        if (escaped) ended = true;                                                   // gpt.cpp:1153-1157
        if (!ended) {
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
        if (!ended) {       // the vertex itself was stored above
            stvw(a, BR_RAYD, slot, mrayD, mpdf);
            stvf(a, BR_THR, slot, mthr);
            SI(a, IF_DEPTH, slot) = depth;
        }
    }

    stvw(a, BR_RAD, slot, mrad, spy);
    SI(a, IF_RNGN, slot) = (int)smp.n;
    tally.rays += rays; tally.vertices += ended ? (unsigned)depth : 0u;              // gpt.cpp:1178-1179
    SI(a, IF_OFLAGS, slot) = (int)flags;
    if (pend) {
        stvf(a, XR_PD_MAIN, slot, mainContributionAll);
        double2 *w = reinterpret_cast<double2 *>(REC(a, XR_PD_W, slot));
        w[0] = make_double2(failReconnect, failHalfVector);
        double2 *b = reinterpret_cast<double2 *>(REC(a, XR_PD_BW, slot));
        b[0] = make_double2(bw0, bw1); b[1] = make_double2(bw2, bw3);
        pend |= bHas << 8;
        if (addBsdfStage) pend |= 1u << 12;
        if (escaped) pend |= 1u << 13;
        if (ended) pend |= 1u << 14;
        if (mainVertexType == VERTEX_TYPE_DIFFUSE) pend |= 1u << 15;
        pend |= offVertexTypes << 16;
        SI(a, IF_PEND, slot) = (int)pend;
        SI(a, IF_BSTYPE, slot) = (int)sampledType;
        setStatus(a, slot, ST_WAIT_RESOLVE, QA_RESOLVE);
    } else if (ended) setStatus(a, slot, ST_FINISHED, QB_GEN);
    else setStatus(a, slot, ST_LIVE, prepareQueue(mits.material));
}

// ------------------------------------------------------------------ resolve: the parked offsets of gpt.cpp:889-1126
GDB_D void stagedResolveBody(const GptArgs &a, int slot)
{
    const unsigned pend = (unsigned)SI(a, IF_PEND, slot);
    unsigned flags = (unsigned)SI(a, IF_OFLAGS, slot), bHas = (pend >> 8) & 15u;
    const bool addBsdfStage = (pend >> 12) & 1u, escaped = (pend >> 13) & 1u, ended = (pend >> 14) & 1u;
    const int mainVertexType = ((pend >> 15) & 1u) ? VERTEX_TYPE_DIFFUSE : VERTEX_TYPE_GLOSSY;
    const unsigned sampledType = (unsigned)SI(a, IF_BSTYPE, slot) & 0xffu;
    const Spec mainContributionAll = ldv(a, XR_PD_MAIN, slot);
    const double2 fw = *reinterpret_cast<const double2 *>(REC(a, XR_PD_W, slot));
    const double2 b01 = reinterpret_cast<const double2 *>(REC(a, XR_PD_BW, slot))[0], b23 = reinterpret_cast<const double2 *>(REC(a, XR_PD_BW, slot))[1];
    Float bw0 = b01.x, bw1 = b01.y, bw2 = b23.x, bw3 = b23.y;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        const int kind = pendKind(pend, i);
        if (kind == PEND_NONE) continue;
        const int o = BR_COUNT + i * OR_COUNT;
        V3 v; Float weight;
        ldvw(a, pendRec(i), slot, v, weight);
        Spec shiftedContribution = splat(0);
        bool alive = true;
        if (kind == PEND_RECONNECT) {
            if (loadOccluded(a, 1 + i, slot)) { alive = false; weight = fw.x; }     // testVisibility failed: gpt.cpp:1131-1136
            else shiftedContribution = v;
        } else {                                                                     // gpt.cpp:1052-1125
            const V3 d = v;
            const int shiftedVertexType2 = ((pend >> (16 + i)) & 1u) ? VERTEX_TYPE_DIFFUSE : VERTEX_TYPE_GLOSSY;
            const Hit h = loadHit(a, 1 + i, slot);
            Spec shiftedEmitterRadiance = splat(0);
            bool ok = true, postponedShiftEnd = false;
            if (h.t == CUDART_INF) {                                                 // gpt.cpp:1052-1074
                if (!c_scene.env.present || !escaped) ok = false;                    // no env, or env vs non-env
                else if (mainVertexType == VERTEX_TYPE_DIFFUSE && shiftedVertexType2 == VERTEX_TYPE_DIFFUSE) ok = false;
                else { shiftedEmitterRadiance = envEval(d); postponedShiftEnd = true; }
            } else if (escaped) ok = false;                                          // gpt.cpp:1078-1082
            else {
                Ray sray; sray.o = ldv(a, o + OR_P, slot); sray.d = d; sray.mint = kEpsilon; sray.maxt = CUDART_INF;
                Its sits; fillIts(sray, h, sits);
                const int shiftedNextVertexType = vertexType(c_sceneG->materials[sits.material], sampledType);
                if (mainVertexType == VERTEX_TYPE_DIFFUSE && shiftedVertexType2 == VERTEX_TYPE_DIFFUSE && shiftedNextVertexType == VERTEX_TYPE_DIFFUSE) ok = false;   // gpt.cpp:1089-1093
                else {
                    if (sits.emitter >= 0) shiftedEmitterRadiance = emittedLe(sits, -d);   // gpt.cpp:1095-1098
                    storeOffItsFull(a, slot, i, sits);
                }
            }
            if (ok) shiftedContribution = ldv(a, o + OR_THR, slot) * shiftedEmitterRadiance;   // gpt.cpp:1107-1112 (weight: from shade)
            else { weight = fw.y; postponedShiftEnd = true; }                        // gpt.cpp:1113-1125
            if (postponedShiftEnd) alive = false;                                    // gpt.cpp:1148-1150
        }
        if (addBsdfStage) {                                                          // gpt.cpp:1140-1146
            const bool has = !(mainContributionAll.x == 0 && mainContributionAll.y == 0 && mainContributionAll.z == 0 && weight == 0);
            if (has) {
                bHas |= 1u << i;
                if (i == 0) bw0 = weight; else if (i == 1) bw1 = weight; else if (i == 2) bw2 = weight; else bw3 = weight;
            }
            accumulateOffset(a, o, slot, shiftedContribution * weight, (shiftedContribution - mainContributionAll) * weight);
        }
        if (!alive) flags = setFlag(flags, i, false, RAY_NOT_CONNECTED);
    }
    Spec mrad; Float spy;
    ldvw(a, BR_RAD, slot, mrad, spy);
    if (bHas & 1u) mrad = mrad + mainContributionAll * bw0;
    if (bHas & 2u) mrad = mrad + mainContributionAll * bw1;
    if (bHas & 4u) mrad = mrad + mainContributionAll * bw2;
    if (bHas & 8u) mrad = mrad + mainContributionAll * bw3;
    stvw(a, BR_RAD, slot, mrad, spy);
    SI(a, IF_OFLAGS, slot) = (int)flags;
    if (ended) setStatus(a, slot, ST_FINISHED, QB_GEN);
    else setStatus(a, slot, ST_LIVE, prepareQueue(SI(a, IF_MAT, slot)));
}

// ------------------------------------------------------------------ stage kernels: persistent CTAs over the stage's queues
enum StageKind { SK_PRIMARY = 0, SK_SHADE0, SK_SHADE1, SK_SHADE2, SK_RESOLVE, SK_PREPARE, SK_GENERATE };
template <int KIND> struct StageQueues;
// minBlocks: resident CTAs per SM the register allocation is sized for (4 => 128 registers, 3 => 168, 2 => 255).  Measured
// (profiles/r02_tracer_history.md): a stage gains nothing from more resident warps but loses ~8 % to a few hundred bytes of
// spills, so each gets the largest occupancy at which it does not spill.
#ifndef GDB_MB_PRIMARY
#define GDB_MB_PRIMARY 3
#endif
#ifndef GDB_MB_SHADE0
#define GDB_MB_SHADE0 2
#endif
#ifndef GDB_MB_SHADE1
#define GDB_MB_SHADE1 2
#endif
#ifndef GDB_MB_SHADE2
#define GDB_MB_SHADE2 2
#endif
#ifndef GDB_MB_RESOLVE
#define GDB_MB_RESOLVE 4
#endif
#ifndef GDB_MB_PREPARE
#define GDB_MB_PREPARE 3
#endif
#ifndef GDB_MB_GENERATE
#define GDB_MB_GENERATE 3
#endif
template <> struct StageQueues<SK_PRIMARY>  { static constexpr int first = QA_PRIMARY, count = 1, minBlocks = GDB_MB_PRIMARY; };
template <> struct StageQueues<SK_SHADE0>   { static constexpr int first = QA_SHADE0, count = kBsdfTypes, minBlocks = GDB_MB_SHADE0; };
template <> struct StageQueues<SK_SHADE1>   { static constexpr int first = QA_SHADE0 + kBsdfTypes, count = kBsdfTypes, minBlocks = GDB_MB_SHADE1; };
template <> struct StageQueues<SK_SHADE2>   { static constexpr int first = QA_SHADE0 + 2 * kBsdfTypes, count = kBsdfTypes, minBlocks = GDB_MB_SHADE2; };
template <> struct StageQueues<SK_RESOLVE>  { static constexpr int first = QA_RESOLVE, count = 1, minBlocks = GDB_MB_RESOLVE; };
template <> struct StageQueues<SK_PREPARE>  { static constexpr int first = QB_PREPARE0, count = kBsdfTypes, minBlocks = GDB_MB_PREPARE; };
template <> struct StageQueues<SK_GENERATE> { static constexpr int first = QB_GEN, count = 1, minBlocks = GDB_MB_GENERATE; };

// What a stage reads first of a slot, requested into L2 one loop iteration ahead.  Every stage is bound by the latency of
// its state loads at the occupancy its registers allow (profiles/r02_stage_kernels_ncu.txt: long_scoreboard is the top
// stall of all of them, L2 hit rates 20-30 %); the queue tells a thread which slot it will process next, so those DRAM
// round trips are started while the current slot is being shaded.
template <int KIND>
GDB_D void prefetchSlot(const GptArgs &a, int slot)
{
#if defined(GDB_STAGE_PREFETCH) || defined(GDB_STAGE_PREFETCH_SELF)     // measured: -8 % (profiles/r02_tracer_history.md)
    auto rec = [&](int r) { prefetchL2(REC(a, r, slot)); };
    auto hit = [&](int id) { prefetchL2(REC(a, hitRec(id), slot)); };
    auto occ = [&](int id) { if (id == 0) prefetchL2(REC(a, XR_OCCLUDED, slot)); };
    prefetchL2(&SI(a, 0, slot)); prefetchL2(&SI(a, 8, slot));
    if (KIND == SK_PRIMARY) {
        for (int id = 0; id < 5; id++) hit(id);
        rec(BR_VD); rec(BR_RAD);
    } else if (KIND == SK_SHADE0 || KIND == SK_SHADE1 || KIND == SK_SHADE2) {
        for (int r = BR_RAYD; r <= BR_RAD; r++) rec(r);
        rec(XR_BS_WO); rec(XR_BS_WEIGHT); hit(0); occ(0);
        for (int i = 0; i < 4; i++) {
            const int o = BR_COUNT + i * OR_COUNT;
            rec(o + OR_THR);
            if (KIND != SK_SHADE2) rec(o + OR_P);
            if (KIND == SK_SHADE0) { occ(1 + i); rec(o + OR_GN); rec(o + OR_S); rec(o + OR_T); rec(o + OR_N); rec(o + OR_WI); }
        }
    } else if (KIND == SK_RESOLVE) {
        rec(XR_PD_MAIN); rec(XR_PD_W); rec(XR_PD_BW); rec(BR_RAD);
        for (int i = 0; i < 4; i++) { rec(pendRec(i)); hit(1 + i); occ(1 + i); }
    } else if (KIND == SK_PREPARE) {
        for (int r = BR_RAYD; r <= BR_WI; r++) rec(r);
    } else {
        for (int i = 0; i < 4; i++) { const int o = BR_COUNT + i * OR_COUNT; rec(o + OR_RAD); rec(o + OR_GRAD); }
        rec(BR_VD); rec(BR_RAD);
    }
#endif
}

// thread -> (queue, index): queues are padded to whole warps so that a warp runs one queue (one BSDF type / shift stage);
// inside a queue the slots are in ascending order up to the compaction's chunk size, so state rows are read near-contiguously.
// The loop is pipelined by two iterations: the slot index of iteration k+2 is loaded and the state of slot k+1 prefetched
// while slot k is processed.
template <int KIND>
__global__ void __launch_bounds__(kStageThreads, StageQueues<KIND>::minBlocks) gpt_stage_kernel(const GptArgs a)
{
    constexpr int first = StageQueues<KIND>::first, nq = StageQueues<KIND>::count;
    __shared__ int s_begin[nq + 1], s_count[nq];
    stampPhase(a);
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < nq; b++) { const int c = a.qCount[first + b]; s_begin[b] = acc; s_count[b] = c; acc += (c + 31) & ~31; }
        s_begin[nq] = acc;
    }
    __syncthreads();
    const int total = s_begin[nq], stride = gridDim.x * blockDim.x;
    auto slotAt = [&](int g) {          // -1: padding of a queue, or past the end
        if (g >= total) return -1;
        int b = 0;
        while (g >= s_begin[b + 1]) b++;
        const int idx = g - s_begin[b];
        return idx < s_count[b] ? a.qList[(size_t)(first + b) * a.nSlots + idx] : -1;
    };
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    int slot = slotAt(g), slot1 = slotAt(g + stride);
    Tally tally;
    for (; g < total; g += stride) {
        const int slot2 = slotAt(g + 2 * stride);
#ifdef GDB_STAGE_PREFETCH
        if (slot1 >= 0) prefetchSlot<KIND>(a, slot1);
#endif
#ifdef GDB_STAGE_PREFETCH_SELF
        if (slot >= 0) prefetchSlot<KIND>(a, slot);
#endif
        if (slot >= 0) {
            if (KIND == SK_PRIMARY) stagedPrimaryBody(a, slot, tally);
            else if (KIND == SK_SHADE0) stagedShadeBody<0>(a, slot, tally);
            else if (KIND == SK_SHADE1) stagedShadeBody<1>(a, slot, tally);
            else if (KIND == SK_SHADE2) stagedShadeBody<2>(a, slot, tally);
            else if (KIND == SK_RESOLVE) stagedResolveBody(a, slot);
            else if (KIND == SK_PREPARE) stagedPrepareBody(a, slot);
            else stagedGenerateBody(a, slot, tally);
        }
        slot = slot1; slot1 = slot2;
    }
    if (KIND != SK_RESOLVE && KIND != SK_PREPARE) flushTally(a, tally);
}

// Ordered stream compaction into the stage queues (ballot / popc ranks per warp, shared-memory prefix over the CTA's
// warps, one atomic per CTA and queue).  PHASE 0 (after the casts): continuations; PHASE 1: slots that need new rays.
template <int PHASE>
__global__ void __launch_bounds__(256) gpt_stage_compact_kernel(const GptArgs a)
{
    constexpr int first = PHASE == 0 ? 0 : kQA, nq = PHASE == 0 ? kQA : kStageBuckets - kQA;
    __shared__ int s_warp[8][nq], s_base[nq];
    stampPhase(a);
    if (blockIdx.x == 0 && threadIdx.x < kStageBuckets) {      // the other phase's queues have been consumed: empty them for its next pass
        const bool mine = (int)threadIdx.x >= first && (int)threadIdx.x < first + nq;
        if (!mine) a.qCount[threadIdx.x] = 0;
    }
    if (PHASE == 0 && blockIdx.x == 0 && threadIdx.x == 0) { a.rayCount[0] = 0; a.rayCount[1] = 0; }   // every queued ray has been cast
    const int slot = blockIdx.x * 256 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int bucket = -1;
    if (slot < a.nSlots) {
        const int key = a.qKey[slot] - first;
        if (key >= 0 && key < nq) bucket = key;
    }
    int rank = 0;
#pragma unroll
    for (int b = 0; b < nq; b++) {
        const unsigned m = __ballot_sync(0xffffffffu, bucket == b);
        if (bucket == b) rank = __popc(m & ((1u << lane) - 1));
        if (lane == 0) s_warp[warp][b] = __popc(m);
    }
    __syncthreads();
    if (threadIdx.x < nq) {
        int tot = 0;
        for (int w = 0; w < 8; w++) { const int c = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = tot; tot += c; }
        s_base[threadIdx.x] = tot ? atomAdd(&a.qCount[first + threadIdx.x], tot) : 0;
    }
    __syncthreads();
    if (bucket >= 0) a.qList[(size_t)(first + bucket) * a.nSlots + s_base[bucket] + s_warp[warp][bucket] + rank] = slot;
}

__global__ void gpt_stamp_kernel(const GptArgs a) { stampPhase(a); }

__global__ void gpt_stage_init_kernel(const GptArgs a)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < kStageBuckets) a.qCount[slot] = 0;
    if (slot == 0) { a.rayCount[0] = 0; a.rayCount[1] = 0; a.counters[6] = (unsigned long long)a.nSlots; }
    if (slot >= a.nSlots) return;
    setStatus(a, slot, ST_FRESH, QB_GEN);
    SI(a, IF_SAMPLE, slot) = 0; SI(a, IF_RNGN, slot) = 0; SI(a, IF_STREAM, slot) = slot;   // slot s starts on stream s
}

}  // namespace gdb200
