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
#include "gpt_kernels.cuh"
#include "gpt_host.h"
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

// ======================================================================================= host ==

using namespace gdb200;

struct gdb200_scene : gdb200::HostScene {
    int device = 0;
    DScene *dScene = nullptr;          // global-memory copy of `host` for per-lane indexed reads
    void *dTables = nullptr;           // one allocation holding the variable-size tables (env map + CDFs, emitter triangles, BVH)
    // device buffers
    double *film = nullptr, *dev64 = nullptr; float *dev32 = nullptr;
    unsigned long long *counters = nullptr;
    volatile int cancel = 0;
};

static std::mutex g_constMutex;    // c_scene is one per device: serialise renders that share a device

// Wavefront workspace (path-slot state + queues): scratch memory that holds nothing between renders, so it is kept per
// device and reused by every scene instead of being allocated and freed with each one (12 GB for 8 M slots: the
// cudaMalloc/cudaFree pair cost more than 100 ms of every end-to-end render).  Guarded by g_constMutex, which already
// serialises the renders of a device.  gdb200_release_workspace() frees it.
struct Workspace {
    double *sd = nullptr; int *si = nullptr;
    int *liveList = nullptr, *liveCount = nullptr, *genList = nullptr, *genCount = nullptr;
    int slotCapacity = 0;
};
static Workspace g_workspace[64];

static void freeWorkspace(Workspace &w)
{
    cudaFree(w.sd); cudaFree(w.si); cudaFree(w.liveList); cudaFree(w.liveCount); cudaFree(w.genList); cudaFree(w.genCount);
    w = Workspace();
}

namespace {

void freeSceneBuffers(gdb200_scene *s)
{
    cudaFree(s->film); cudaFree(s->dev64); cudaFree(s->dev32);
    cudaFree(s->counters); cudaFree(s->dScene); s->dScene = nullptr; cudaFree(s->dTables); s->dTables = nullptr;
    s->film = s->dev64 = nullptr; s->dev32 = nullptr; s->counters = nullptr;
}

// Variable-size tables go to one device allocation (once per scene); their device addresses are patched into the
// DScene copies that the kernels read.
int uploadTables(gdb200_scene *s)
{
    if (s->dTables) return GDB200_OK;
    struct Part { const void *src; size_t bytes; size_t offset; };
    std::vector<Part> parts = {
        {s->envTexels.data(), s->envTexels.size() * sizeof(Float), 0}, {s->envRowWeights.data(), s->envRowWeights.size() * sizeof(Float), 0},
        {s->emTriCdf.data(), s->emTriCdf.size() * sizeof(Float), 0}, {s->envCdfRows.data(), s->envCdfRows.size() * sizeof(float), 0},
        {s->envCdfCols.data(), s->envCdfCols.size() * sizeof(float), 0}, {s->emTris.data(), s->emTris.size() * sizeof(DEmTri), 0},
        {s->bvh.data(), s->bvh.size() * sizeof(BvhNode), 0}, {s->bvhTris.data(), s->bvhTris.size() * sizeof(DTri), 0},
        {s->triNormals.data(), s->triNormals.size() * sizeof(V3), 0}};
    size_t total = 0;
    for (Part &p : parts) { p.offset = total; total += (p.bytes + 255) & ~(size_t)255; }
    if (total == 0) return GDB200_OK;
    GDB_CUDA(cudaMalloc(&s->dTables, total));
    char *base = (char *)s->dTables;
    for (const Part &p : parts) if (p.bytes) GDB_CUDA(cudaMemcpy(base + p.offset, p.src, p.bytes, cudaMemcpyHostToDevice));
    DScene &h = s->host;
    h.env.texels = (const Float *)(base + parts[0].offset); h.env.rowWeights = (const Float *)(base + parts[1].offset);
    h.emTriCdf = (const Float *)(base + parts[2].offset); h.env.cdfRows = (const float *)(base + parts[3].offset);
    h.env.cdfCols = (const float *)(base + parts[4].offset); h.emTris = (const DEmTri *)(base + parts[5].offset);
    h.bvh = (const BvhNode *)(base + parts[6].offset); h.bvhTris = (const DTri *)(base + parts[7].offset);
    h.triNormals = (const V3 *)(base + parts[8].offset);
    return GDB200_OK;
}

int uploadScene(gdb200_scene *s)
{
    if (int rc = uploadTables(s)) return rc;
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

void gdb200_release_workspace(void)
{
    std::lock_guard<std::mutex> lock(g_constMutex);
    int current = 0, n = 0;
    if (cudaGetDevice(&current) != cudaSuccess || cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return; }
    for (int d = 0; d < n && d < 64; d++)
        if (g_workspace[d].slotCapacity) { cudaSetDevice(d); freeWorkspace(g_workspace[d]); }
    cudaSetDevice(current);
}

int gdb200_gpt_render(gdb200_scene *s, const gdb200_gpt_params *p, gdb200_buffers *out, gdb200_stats *stats)
{
    if (!s || !p) return set_error(GDB200_ERR_ARGUMENT, "scene/params is NULL");
    GptArgs a;
    const char *capEnv = getenv("GDB200_MAX_SLOTS");        // resident path slots (tuning knob; streams beyond it are dealt out as slots drain)
    if (int rc = setupArgs(*s, p, a, capEnv ? atoi(capEnv) : (1 << 23))) return rc;
    GDB_CUDA(cudaSetDevice(s->device));
    const int nSlots = a.nSlots;
    if (s->device < 0 || s->device >= 64) return set_error(GDB200_ERR_ARGUMENT, "device index %d out of range", s->device);
    classifyMaterials(s, p->shift_threshold);

    std::lock_guard<std::mutex> lock(g_constMutex);
    Workspace &ws = g_workspace[s->device];
    if (nSlots > ws.slotCapacity) {
        freeWorkspace(ws);
        static_assert(IF_COUNT <= 16, "int fields must fit the 16-int slot line");
        cudaError_t e = cudaMalloc(&ws.sd, sizeof(double) * 4 * kRecords * (size_t)nSlots);
        if (e == cudaSuccess) e = cudaMalloc(&ws.si, sizeof(int) * 16 * (size_t)nSlots);
        if (e == cudaSuccess) e = cudaMalloc(&ws.liveList, sizeof(int) * 2 * (size_t)kBuckets * nSlots);
        if (e == cudaSuccess) e = cudaMalloc(&ws.liveCount, sizeof(int) * 2 * kBuckets);
        if (e == cudaSuccess) e = cudaMalloc(&ws.genList, sizeof(int) * 2 * (size_t)nSlots);
        if (e == cudaSuccess) e = cudaMalloc(&ws.genCount, sizeof(int) * 2);
        if (e != cudaSuccess) { freeWorkspace(ws); cudaGetLastError(); return set_error(GDB200_ERR_CUDA, "wavefront workspace for %d path slots: %s", nSlots, cudaGetErrorString(e)); }
        ws.slotCapacity = nSlots;
    }
    if (int rc = uploadScene(s)) return rc;
    GDB_CUDA(cudaMemset(s->film, 0, sizeof(double) * 5 * (size_t)s->width * s->height * 4));
    GDB_CUDA(cudaMemset(s->counters, 0, sizeof(unsigned long long) * 8));

    a.sd = ws.sd; a.si = ws.si;
    a.film = s->film; a.liveList = ws.liveList; a.liveCount = ws.liveCount; a.genList = ws.genList; a.genCount = ws.genCount;
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
    if (getenv("GDB200_NO_TAIL")) tailThreshold = 0;        // test knob: drain the wavefront through its queues to the last path
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
