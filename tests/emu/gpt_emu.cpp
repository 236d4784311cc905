// TEST INFRASTRUCTURE — runs the tracer's per-slot device routines (generateBody / bounceBody of
// gradientdomain-mitsuba_b200/csrc/gpt_kernels.cuh, with the intersection, BSDF, emitter and shift code of
// gpt_device.cuh and the scene flattening of gpt_host.h) serially on the CPU through tests/emu/cuda_emu.h.
// Purpose: check new device code against the oracle in the `-m "not gpu"` suite, where no GPU exists.
// The wavefront scheduling itself (queues, compaction, launches) is NOT emulated: every slot is run to
// completion like gpt_tail_kernel does.  Only tests/ may build or load this; the product never does.
#include "cuda_emu.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../include/gdb200.h"

namespace gdb200 {
static std::string g_emuError;
int set_error(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_emuError = buf;
    return code;
}
}
#include "../../gradientdomain-mitsuba_b200/csrc/gpt_kernels.cuh"
#include "../../gradientdomain-mitsuba_b200/csrc/gpt_stages.cuh"
#include "../../gradientdomain-mitsuba_b200/csrc/gpt_host.h"

using namespace gdb200;

extern "C" const char *gdb200_emu_last_error(void) { return g_emuError.c_str(); }

// Same contract as gdb200_gpt_render (include/gdb200.h); counters (optional, 6 doubles): done slots, rays,
// path vertices, samples, state bytes, path bounces.
extern "C" int gdb200_emu_gpt_render(const gdb200_scene_desc *desc, const gdb200_gpt_params *p, gdb200_buffers *out, double *counters)
{
    static HostScene hs;     // c_scene & co are file-scope statics: one render at a time
    if (int rc = flattenScene(desc, &hs)) return rc;
    classifyMaterials(&hs, p->shift_threshold);
    GptArgs a;
    std::vector<double> sd, film; std::vector<int> si; std::vector<unsigned long long> ctr(8, 0);
    const char *capEnv = getenv("GDB200_MAX_SLOTS");
    if (int rc = setupArgs(hs, p, a, capEnv ? atoi(capEnv) : (1 << 23))) return rc;
    hs.host.env.texels = hs.envTexels.data(); hs.host.env.rowWeights = hs.envRowWeights.data(); hs.host.emTriCdf = hs.emTriCdf.data();
    hs.host.env.cdfRows = hs.envCdfRows.data(); hs.host.env.cdfCols = hs.envCdfCols.data(); hs.host.emTris = hs.emTris.data();
    hs.host.bvh = hs.bvh.data(); hs.host.bvhTris = hs.bvhTris.data(); hs.host.triNormals = hs.triNormals.data();
    c_scene = hs.host; static DScene sceneCopy; sceneCopy = hs.host; c_sceneG = &sceneCopy;
    memcpy(c_bounds, hs.bounds, sizeof(hs.bounds));
    const size_t n = (size_t)hs.width * hs.height;
    sd.assign((size_t)4 * kRecPitch * a.nSlots, 0.0); si.assign((size_t)16 * a.nSlots, 0); film.assign(5 * n * 4, 0.0);
    std::vector<int> lists((size_t)2 * a.nSlots + 2 * kBuckets + 2, 0);
    a.sd = sd.data(); a.si = si.data(); a.film = film.data(); a.counters = ctr.data();
    a.genList = lists.data(); a.liveCount = lists.data() + 2 * (size_t)a.nSlots; a.genCount = a.liveCount + 2 * kBuckets; a.liveList = nullptr;
    blockDim.x = 1; threadIdx.x = 0;
    for (int s = 0; s < std::max(a.nSlots, 2 * kBuckets); s++) { blockIdx.x = (unsigned)s; gpt_init_kernel(a); }
    for (int s = 0; s < a.nSlots; s++) runSlotToCompletion(a, s);
    std::vector<double> dev64(5 * n * 3); std::vector<float> dev32(5 * n * 3);
    for (size_t i = 0; i < 5 * n; i++) { blockIdx.x = (unsigned)i; gpt_develop_kernel(film.data(), (int)n, dev64.data(), dev32.data()); }
    if (out) {
        double *dst[5] = {out->preview_final, out->throughput, out->dx, out->dy, out->direct};
        for (int b = 0; b < 5; b++) if (dst[b]) memcpy(dst[b], dev64.data() + (size_t)b * n * 3, sizeof(double) * n * 3);
    }
    if (counters) for (int i = 0; i < 6; i++) counters[i] = (double)ctr[i];
    return 0;
}

// ---- single-plugin entry points of the DEVICE routines for tests/test_chisquare.py (the reference's own consistency
// test of BSDF::sample/eval/pdf and Emitter::sampleDirect/pdfDirect, src/tests/test_chisquare.cpp)
static void emuMaterial(const gdb200_material *m, HostScene &hs)
{
    memset(&hs.host, 0, sizeof(hs.host));
    hs.mats.assign(1, *m);
    classifyMaterials(&hs, 0.001);
}
extern "C" int gdb200_emu_bsdf_sample_batch(const gdb200_material *m, const double *wi, int n, const double *samples,
                                            double *wo, double *weight, double *pdf, int *sampledType)
{
    static HostScene hs; emuMaterial(m, hs);
    for (int i = 0; i < n; i++) {
        BSDFSample bs;
        bsdfSample(hs.host.materials[0], mk(wi[0], wi[1], wi[2]), samples[3 * i], samples[3 * i + 1], samples[3 * i + 2], bs);
        wo[3 * i] = bs.wo.x; wo[3 * i + 1] = bs.wo.y; wo[3 * i + 2] = bs.wo.z;
        weight[3 * i] = bs.weight.x; weight[3 * i + 1] = bs.weight.y; weight[3 * i + 2] = bs.weight.z;
        pdf[i] = bs.pdf; sampledType[i] = (int)bs.sampledType;
    }
    return 0;
}
extern "C" int gdb200_emu_bsdf_eval_batch(const gdb200_material *m, const double *wi, int n, const double *wo, int measure,
                                          double *value, double *pdf)
{
    static HostScene hs; emuMaterial(m, hs);
    for (int i = 0; i < n; i++) {
        Spec f; Float p;
        bsdfEvalPdf(hs.host.materials[0], mk(wi[0], wi[1], wi[2]), mk(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), measure ? EDiscrete : ESolidAngle, f, p);
        value[3 * i] = f.x; value[3 * i + 1] = f.y; value[3 * i + 2] = f.z; pdf[i] = p;
    }
    return 0;
}
static int emuEnvScene(const gdb200_scene_desc *desc, HostScene &hs)
{
    if (int rc = flattenScene(desc, &hs)) return rc;
    if (!hs.host.env.present) return set_error(GDB200_ERR_ARGUMENT, "scene has no environment emitter");
    classifyMaterials(&hs, 0.001);
    // isolate the environment emitter: one-entry selection CDF
    hs.host.emitters[0] = hs.host.emitters[hs.host.env.emitter]; hs.host.emitters[0].pdfDiscrete = 1.0;
    hs.host.nEmitters = 1; hs.host.emCdf[0] = 0; hs.host.emCdf[1] = 1; hs.host.env.emitter = 0;
    hs.host.env.texels = hs.envTexels.data(); hs.host.env.rowWeights = hs.envRowWeights.data();
    hs.host.env.cdfRows = hs.envCdfRows.data(); hs.host.env.cdfCols = hs.envCdfCols.data();
    hs.host.emTris = hs.emTris.data(); hs.host.emTriCdf = hs.emTriCdf.data(); hs.host.triNormals = hs.triNormals.data();
    hs.host.bvh = hs.bvh.data(); hs.host.bvhTris = hs.bvhTris.data();
    c_scene = hs.host; static DScene sceneCopy; sceneCopy = hs.host; c_sceneG = &sceneCopy;
    memcpy(c_bounds, hs.bounds, sizeof(hs.bounds));
    return 0;
}
extern "C" int gdb200_emu_envmap_sample_batch(const gdb200_scene_desc *desc, const double *ref, int n, const double *samples, double *d, double *pdf)
{
    static HostScene hs;
    if (int rc = emuEnvScene(desc, hs)) return rc;
    for (int i = 0; i < n; i++) {
        DRec r; r.ref = mk(ref[0], ref[1], ref[2]); r.refN = mk(0, 0, 0);
        bool vis;
        sampleEmitterDirectVisible(r, samples[2 * i], samples[2 * i + 1], vis);
        d[3 * i] = r.d.x; d[3 * i + 1] = r.d.y; d[3 * i + 2] = r.d.z; pdf[i] = r.pdf;
    }
    return 0;
}
extern "C" int gdb200_emu_envmap_pdf_batch(const gdb200_scene_desc *desc, int n, const double *d, double *pdf)
{
    static HostScene hs;
    if (int rc = emuEnvScene(desc, hs)) return rc;
    for (int i = 0; i < n; i++) pdf[i] = envPdfDirection(xfVector(c_scene.env.toObject, mk(d[3 * i], d[3 * i + 1], d[3 * i + 2])));
    return 0;
}

// ---- block mode: the queued wavefront (gpt_generate_kernel -> gpt_compact_kernel -> gpt_bounce_kernel<2> per step, then
// gpt_tail_kernel) with every CTA run by OS threads, so that the kernels' shared-memory prologues, __syncthreads and
// ballots execute as written.  The step loop below mirrors the host loop of gdb200_gpt_render (csrc/gpt.cu): it is a
// restatement for the test, the kernels are the real source.
#include <functional>
#include <thread>

static void emuLaunch(int blocks, int threads, const std::function<void()> &kernel)
{
    EmuBlock blk; blk.nThreads = threads;
    blockDim.x = (unsigned)threads; gridDim.x = (unsigned)blocks;
    for (int b = 0; b < blocks; b++) {
        pthread_barrier_init(&blk.bar, nullptr, (unsigned)threads);
        emu_block = &blk;
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, b, t]() { blockIdx.x = (unsigned)b; threadIdx.x = (unsigned)t; kernel(); });
        for (std::thread &th : pool) th.join();
        emu_block = nullptr;
        pthread_barrier_destroy(&blk.bar);
    }
    blockDim.x = 1; gridDim.x = 1; blockIdx.x = 0; threadIdx.x = 0;
}

extern "C" int gdb200_emu_gpt_render_wavefront(const gdb200_scene_desc *desc, const gdb200_gpt_params *p, gdb200_buffers *out, double *counters)
{
    static HostScene hs;
    if (int rc = flattenScene(desc, &hs)) return rc;
    classifyMaterials(&hs, p->shift_threshold);
    GptArgs a;
    const char *capEnv = getenv("GDB200_MAX_SLOTS");
    if (int rc = setupArgs(hs, p, a, capEnv ? atoi(capEnv) : (1 << 23))) return rc;
    hs.host.env.texels = hs.envTexels.data(); hs.host.env.rowWeights = hs.envRowWeights.data(); hs.host.emTriCdf = hs.emTriCdf.data();
    hs.host.env.cdfRows = hs.envCdfRows.data(); hs.host.env.cdfCols = hs.envCdfCols.data(); hs.host.emTris = hs.emTris.data();
    hs.host.bvh = hs.bvh.data(); hs.host.bvhTris = hs.bvhTris.data(); hs.host.triNormals = hs.triNormals.data();
    c_scene = hs.host; static DScene sceneCopy; sceneCopy = hs.host; c_sceneG = &sceneCopy;
    memcpy(c_bounds, hs.bounds, sizeof(hs.bounds));
    const size_t n = (size_t)hs.width * hs.height;
    const int nSlots = a.nSlots;
    std::vector<double> sd((size_t)4 * kRecPitch * nSlots, 0.0), film(5 * n * 4, 0.0);
    std::vector<int> si((size_t)16 * nSlots, 0), liveList((size_t)2 * kBuckets * nSlots, -1), liveCount(2 * kBuckets, 0), genList((size_t)2 * nSlots, -1), genCount(2, 0);
    std::vector<unsigned long long> ctr(8, 0);
    a.sd = sd.data(); a.si = si.data(); a.film = film.data(); a.counters = ctr.data();
    a.liveList = liveList.data(); a.liveCount = liveCount.data(); a.genList = genList.data(); a.genCount = genCount.data();

    emuLaunch((nSlots + 255) / 256, 256, [&]() { gpt_init_kernel(a); });
    const int genBlocks = (nSlots + kGenThreads - 1) / kGenThreads;
    const int bounceBlocks = (nSlots + 32 * kBuckets + kBounceThreads - 1) / kBounceThreads;
    unsigned long long tailThreshold = (unsigned long long)std::max(nSlots / 16, std::min(nSlots, 16384));
    if (getenv("GDB200_NO_TAIL")) tailThreshold = 0;
    int parity = 0;
    const long long maxSteps = (long long)p->spp * 4096 + 65536;
    for (long long step = 0;; step++) {
        if (step > maxSteps) return set_error(GDB200_ERR_CUDA, "wavefront did not drain after %lld steps", step);
        emuLaunch(genBlocks, kGenThreads, [&]() { gpt_generate_kernel(a, parity); });
        emuLaunch((nSlots + 255) / 256, 256, [&]() { gpt_compact_kernel(a, parity); });
        for (int b = 0; b < kBuckets; b++)                       // every queue entry must be a valid slot of the right bucket
            for (int i = 0; i < liveCount[parity * kBuckets + b]; i++) {
                const int slot = liveList[((size_t)parity * kBuckets + b) * nSlots + i];
                if (slot < 0 || slot >= nSlots || si[(size_t)slot * 16 + IF_STATUS] != ST_LIVE) return set_error(GDB200_ERR_CUDA, "step %lld: bad entry %d in queue %d", step, slot, b);
            }
        emuLaunch(bounceBlocks, kBounceThreads, [&]() { gpt_bounce_kernel<2>(a, parity); });
        parity ^= 1;
        if ((step & 15) == 15) {
            if (ctr[0] >= (unsigned long long)nSlots) break;
            if ((unsigned long long)nSlots - ctr[0] <= tailThreshold) {
                emuLaunch((nSlots + kBounceThreads - 1) / kBounceThreads, kBounceThreads, [&]() { gpt_tail_kernel(a); });
                break;
            }
        }
    }
    std::vector<double> dev64(5 * n * 3); std::vector<float> dev32(5 * n * 3);
    blockDim.x = 1; threadIdx.x = 0;
    for (size_t i = 0; i < 5 * n; i++) { blockIdx.x = (unsigned)i; gpt_develop_kernel(film.data(), (int)n, dev64.data(), dev32.data()); }
    if (out) {
        double *dst[5] = {out->preview_final, out->throughput, out->dx, out->dy, out->direct};
        for (int b = 0; b < 5; b++) if (dst[b]) memcpy(dst[b], dev64.data() + (size_t)b * n * 3, sizeof(double) * n * 3);
    }
    if (counters) for (int i = 0; i < 6; i++) counters[i] = (double)ctr[i];
    return 0;
}

// ---- gpt_check_culling_kernel on the host: candidate selection by padded bounds (closestPrimitive, incl. the BVH walk)
// against testing every primitive, on random rays.
extern "C" int gdb200_emu_check_culling(const gdb200_scene_desc *desc, int nRays, unsigned long long seed, unsigned long long *out2)
{
    static HostScene hs;
    if (int rc = flattenScene(desc, &hs)) return rc;
    classifyMaterials(&hs, 0.001);
    hs.host.env.texels = hs.envTexels.data(); hs.host.env.rowWeights = hs.envRowWeights.data(); hs.host.emTriCdf = hs.emTriCdf.data();
    hs.host.env.cdfRows = hs.envCdfRows.data(); hs.host.env.cdfCols = hs.envCdfCols.data(); hs.host.emTris = hs.emTris.data();
    hs.host.bvh = hs.bvh.data(); hs.host.bvhTris = hs.bvhTris.data(); hs.host.triNormals = hs.triNormals.data();
    c_scene = hs.host; static DScene sceneCopy; sceneCopy = hs.host; c_sceneG = &sceneCopy;
    memcpy(c_bounds, hs.bounds, sizeof(hs.bounds));
    out2[0] = out2[1] = 0;
    blockDim.x = 1; threadIdx.x = 0;
    for (int g = 0; g < nRays; g++) { blockIdx.x = (unsigned)g; gpt_check_culling_kernel(seed, nRays, out2); }
    blockIdx.x = 0;
    return 0;
}

// ---- block mode of the STAGED wavefront (csrc/gpt_stages.cuh): compact(A) -> primary, shade, resolve -> compact(B) ->
// prepare, generate -> casts, every kernel as written (persistent CTAs looping over their queues, ballots, warp-aggregated
// ray appends), CTAs run by OS threads.  The tick loop mirrors renderStaged() of csrc/gpt.cu.  Every queue entry is checked
// against the status it must have.
extern "C" int gdb200_emu_gpt_render_staged(const gdb200_scene_desc *desc, const gdb200_gpt_params *p, gdb200_buffers *out, double *counters, int grid)
{
    static HostScene hs;
    if (int rc = flattenScene(desc, &hs)) return rc;
    classifyMaterials(&hs, p->shift_threshold);
    GptArgs a;
    const char *capEnv = getenv("GDB200_MAX_SLOTS");
    if (int rc = setupArgs(hs, p, a, capEnv ? atoi(capEnv) : (1 << 23))) return rc;
    hs.host.env.texels = hs.envTexels.data(); hs.host.env.rowWeights = hs.envRowWeights.data(); hs.host.emTriCdf = hs.emTriCdf.data();
    hs.host.env.cdfRows = hs.envCdfRows.data(); hs.host.env.cdfCols = hs.envCdfCols.data(); hs.host.emTris = hs.emTris.data();
    hs.host.bvh = hs.bvh.data(); hs.host.bvhTris = hs.bvhTris.data(); hs.host.triNormals = hs.triNormals.data();
    c_scene = hs.host; static DScene sceneCopy; sceneCopy = hs.host; c_sceneG = &sceneCopy;
    memcpy(c_bounds, hs.bounds, sizeof(hs.bounds));
    const size_t n = (size_t)hs.width * hs.height;
    const int nSlots = a.nSlots;
    const double nan = std::numeric_limits<double>::quiet_NaN();      // a stage that reads a record nobody wrote shows up in the film
    std::vector<double> sd((size_t)4 * kRecPitch * nSlots, nan), film(5 * n * 4, 0.0), rays0((size_t)8 * 5 * nSlots, nan), rays1((size_t)8 * 5 * nSlots, nan);
    std::vector<int> si((size_t)16 * nSlots, 0), owner0((size_t)5 * nSlots, -1), owner1((size_t)5 * nSlots, -1),
        qKey(nSlots, -1), qList((size_t)kStageBuckets * nSlots, -1), qCount(kStageBuckets, 0), rayCount(2, 0);
    std::vector<unsigned long long> ctr(8, 0);
    a.sd = sd.data(); a.si = si.data(); a.film = film.data(); a.counters = ctr.data();
    a.rays[0] = rays0.data(); a.rays[1] = rays1.data(); a.rayOwner[0] = owner0.data(); a.rayOwner[1] = owner1.data();
    a.rayCount = rayCount.data(); a.rayCapacity = 5 * nSlots;
    a.qKey = qKey.data(); a.qList = qList.data(); a.qCount = qCount.data();

    grid = std::max(1, grid);
    emuLaunch((std::max(nSlots, kStageBuckets) + 255) / 256, 256, [&]() { gpt_stage_init_kernel(a); });
    const int compactBlocks = (nSlots + 255) / 256;
    auto checkQueues = [&](int first, int count, long long tick) -> int {
        for (int b = first; b < first + count; b++)
            for (int i = 0; i < qCount[b]; i++) {
                const int slot = qList[(size_t)b * nSlots + i];
                if (slot < 0 || slot >= nSlots) return set_error(GDB200_ERR_CUDA, "tick %lld: bad entry %d in queue %d", tick, slot, b);
                const int st = si[(size_t)slot * 16 + IF_STATUS];
                const bool ok = b == QA_PRIMARY ? st == ST_WAIT_PRIMARY : b == QA_RESOLVE ? st == ST_WAIT_RESOLVE : b < QA_RESOLVE ? st == ST_WAIT_SHADE
                              : b == QB_GEN ? (st == ST_FINISHED || st == ST_FRESH) : st == ST_LIVE;
                if (!ok) return set_error(GDB200_ERR_CUDA, "tick %lld: slot %d with status %d in queue %d", tick, slot, st, b);
            }
        return 0;
    };
    const long long maxTicks = ((long long)p->spp * 8192 + 131072) * std::max(1, a.nStreams / std::max(1, nSlots) + 1);
    for (long long tick = 0;; tick++) {
        if (tick > maxTicks) return set_error(GDB200_ERR_CUDA, "wavefront did not drain after %lld ticks", tick);
        emuLaunch(compactBlocks, 256, [&]() { gpt_stage_compact_kernel<0>(a); });
        if (int rc = checkQueues(0, kQA, tick)) return rc;
        if (rayCount[0] || rayCount[1]) return set_error(GDB200_ERR_CUDA, "tick %lld: ray queues not reset", tick);
        emuLaunch(grid, kStageThreads, [&]() { gpt_stage_kernel<SK_PRIMARY>(a); });
        emuLaunch(grid, kStageThreads, [&]() { gpt_stage_kernel<SK_SHADE0>(a); });
        emuLaunch(grid, kStageThreads, [&]() { gpt_stage_kernel<SK_SHADE1>(a); });
        emuLaunch(grid, kStageThreads, [&]() { gpt_stage_kernel<SK_SHADE2>(a); });
        emuLaunch(grid, kStageThreads, [&]() { gpt_stage_kernel<SK_RESOLVE>(a); });
        emuLaunch(compactBlocks, 256, [&]() { gpt_stage_compact_kernel<1>(a); });
        if (int rc = checkQueues(kQA, kStageBuckets - kQA, tick)) return rc;
        emuLaunch(grid, kStageThreads, [&]() { gpt_stage_kernel<SK_PREPARE>(a); });
        emuLaunch(grid, kStageThreads, [&]() { gpt_stage_kernel<SK_GENERATE>(a); });
        emuLaunch(grid, 128, [&]() { gpt_cast_kernel<false>(a); });
        emuLaunch(grid, 128, [&]() { gpt_cast_kernel<true>(a); });
        if (ctr[7]) return set_error(GDB200_ERR_CUDA, "ray queue overflow");
        if ((tick & 15) == 15 && ctr[0] >= (unsigned long long)nSlots) break;
    }
    std::vector<double> dev64(5 * n * 3); std::vector<float> dev32(5 * n * 3);
    blockDim.x = 1; threadIdx.x = 0;
    for (size_t i = 0; i < 5 * n; i++) { blockIdx.x = (unsigned)i; gpt_develop_kernel(film.data(), (int)n, dev64.data(), dev32.data()); }
    if (out) {
        double *dst[5] = {out->preview_final, out->throughput, out->dx, out->dy, out->direct};
        for (int b = 0; b < 5; b++) if (dst[b]) memcpy(dst[b], dev64.data() + (size_t)b * n * 3, sizeof(double) * n * 3);
    }
    if (counters) for (int i = 0; i < 6; i++) counters[i] = (double)ctr[i];
    return 0;
}
