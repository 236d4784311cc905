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
#include "gpt_stages.cuh"
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

constexpr int kMaxDevices = 64;
// c_scene / c_sceneG / c_bounds and the wavefront workspace exist once per device: renders that share a device are
// serialised, renders on different devices of one process (the Mitsuba plugin driving several GPUs) run concurrently.
static std::mutex g_deviceMutex[kMaxDevices];

// Wavefront workspace (path-slot state + queues): scratch memory that holds nothing between renders, so it is kept per
// device and reused by every scene instead of being allocated and freed with each one (~2.8 KB per resident path: the
// cudaMalloc/cudaFree pair cost more than 100 ms of every end-to-end render).  Guarded by g_deviceMutex[device].
// gdb200_release_workspace() frees it.
struct Workspace {
    double *sd = nullptr; int *si = nullptr;
    int *liveList = nullptr, *liveCount = nullptr, *genList = nullptr, *genCount = nullptr;     // fused-bounce wavefront (A/B)
    double *rays[2] = {nullptr, nullptr};                                                        // staged wavefront
    int *rayOwner[2] = {nullptr, nullptr}, *rayCount = nullptr, *qKey = nullptr, *qList = nullptr, *qCount = nullptr;
    int slotCapacity = 0; bool fused = false;
    std::vector<cudaEvent_t> events;       // timing marks of renders that ask for stats, reused across renders
    unsigned long long *stamps = nullptr;  // phase stamps of the staged wavefront (stampPhase), kStampCapacity entries
    std::vector<unsigned long long> stampsHost;
    // Film buffers of destroyed scenes, kept for the next scene of the same size (a renderer that keeps its film between
    // frames): cudaFree of the 168 MB film is a device-wide synchronisation that took 2 ... 600 ms on the GPU boxes.
    struct FilmSet { double *film = nullptr, *dev64 = nullptr; float *dev32 = nullptr; unsigned long long *counters = nullptr; size_t pixels = 0; };
    std::vector<FilmSet> spareFilms;
};
static Workspace g_workspace[kMaxDevices];
constexpr size_t kStampCapacity = (size_t)1 << 20;     // 8 stamps per tick: 131 072 ticks, then the phase times stop growing

static void freeWorkspace(Workspace &w)
{
    cudaFree(w.sd); cudaFree(w.si); cudaFree(w.liveList); cudaFree(w.liveCount); cudaFree(w.genList); cudaFree(w.genCount);
    cudaFree(w.rays[0]); cudaFree(w.rays[1]); cudaFree(w.rayOwner[0]); cudaFree(w.rayOwner[1]);
    cudaFree(w.rayCount); cudaFree(w.qKey); cudaFree(w.qList); cudaFree(w.qCount);
    for (cudaEvent_t e : w.events) cudaEventDestroy(e);
    cudaFree(w.stamps);
    for (const Workspace::FilmSet &f : w.spareFilms) { cudaFree(f.film); cudaFree(f.dev64); cudaFree(f.dev32); cudaFree(f.counters); }
    w = Workspace();
}

// Restores the caller's current device when an entry point that binds to the scene's device returns.
struct DeviceGuard {
    int previous = -1;
    ~DeviceGuard() { if (previous >= 0) cudaSetDevice(previous); }
    int bind(int device)
    {
        if (device < 0 || device >= kMaxDevices) return set_error(GDB200_ERR_ARGUMENT, "device index %d out of range", device);
        if (cudaGetDevice(&previous) != cudaSuccess) { previous = -1; cudaGetLastError(); }
        GDB_CUDA(cudaSetDevice(device));
        return GDB200_OK;
    }
};

namespace {

constexpr size_t kSpareFilms = 2;

void freeSceneBuffers(gdb200_scene *s)
{
    bool kept = false;
    if (s->film && s->dev64 && s->dev32 && s->counters && s->device >= 0 && s->device < kMaxDevices) {
        std::lock_guard<std::mutex> lock(g_deviceMutex[s->device]);
        Workspace &ws = g_workspace[s->device];
        const size_t pixels = (size_t)s->width * s->height;
        if (ws.spareFilms.size() < kSpareFilms && pixels * 340 <= ((size_t)4 << 30)) {      // 340 B per pixel; larger films are given back
            Workspace::FilmSet f;
            f.film = s->film; f.dev64 = s->dev64; f.dev32 = s->dev32; f.counters = s->counters; f.pixels = (size_t)s->width * s->height;
            ws.spareFilms.push_back(f);
            kept = true;
        }
    }
    if (!kept) { cudaFree(s->film); cudaFree(s->dev64); cudaFree(s->dev32); cudaFree(s->counters); }
    cudaFree(s->dScene); s->dScene = nullptr; cudaFree(s->dTables); s->dTables = nullptr;
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

namespace {

// Timing marks of a render that asked for stats: events come from the workspace's pool (created once, reused).
struct Marks {
    Workspace &ws; bool on; size_t used = 0;
    Marks(Workspace &w, bool enabled) : ws(w), on(enabled) {}
    void mark()
    {
        if (!on) return;
        if (used == ws.events.size()) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) { on = false; return; } ws.events.push_back(e); }
        cudaEventRecord(ws.events[used++]);
    }
    float between(size_t i) const { float ms = 0; cudaEventElapsedTime(&ms, ws.events[i], ws.events[i + 1]); return ms; }
};

int ensureWorkspace(Workspace &ws, int nSlots, bool fused)
{
    if (nSlots <= ws.slotCapacity && fused == ws.fused) return GDB200_OK;
    std::vector<cudaEvent_t> keep; keep.swap(ws.events);
    freeWorkspace(ws);
    ws.events.swap(keep);
    static_assert(IF_COUNT <= 16, "int fields must fit the 16-int slot line");
    const size_t n = (size_t)nSlots;
    cudaError_t e = cudaMalloc(&ws.sd, sizeof(double) * 4 * kRecPitch * n);
    if (e == cudaSuccess) e = cudaMalloc(&ws.si, sizeof(int) * 16 * n);
    if (fused) {
        if (e == cudaSuccess) e = cudaMalloc(&ws.liveList, sizeof(int) * 2 * (size_t)kBuckets * n);
        if (e == cudaSuccess) e = cudaMalloc(&ws.liveCount, sizeof(int) * 2 * kBuckets);
        if (e == cudaSuccess) e = cudaMalloc(&ws.genList, sizeof(int) * 2 * n);
        if (e == cudaSuccess) e = cudaMalloc(&ws.genCount, sizeof(int) * 2);
    } else {
        for (int q = 0; q < 2; q++) {
            if (e == cudaSuccess) e = cudaMalloc(&ws.rays[q], sizeof(double) * 8 * 5 * n);
            if (e == cudaSuccess) e = cudaMalloc(&ws.rayOwner[q], sizeof(int) * (5 * n + 4));      // + the 16-byte rounding of a bulk copy
        }
        if (e == cudaSuccess) e = cudaMalloc(&ws.rayCount, sizeof(int) * 2);
        if (e == cudaSuccess) e = cudaMalloc(&ws.qKey, sizeof(int) * n);
        if (e == cudaSuccess) e = cudaMalloc(&ws.qList, sizeof(int) * (size_t)kStageBuckets * n);
        if (e == cudaSuccess) e = cudaMalloc(&ws.qCount, sizeof(int) * kStageBuckets);
    }
    if (e != cudaSuccess) {
        keep.swap(ws.events); freeWorkspace(ws); ws.events.swap(keep); cudaGetLastError();
        return set_error(GDB200_ERR_CUDA, "wavefront workspace for %d path slots: %s", nSlots, cudaGetErrorString(e));
    }
    ws.slotCapacity = nSlots; ws.fused = fused;
    return GDB200_OK;
}

// Resident CTAs of a persistent kernel on this device: SM count x occupancy (one wave, the CTAs loop over their queue).
template <class K> int persistentGrid(K kernel, int threads, int device)
{
    int perSM = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, threads, 0) != cudaSuccess || perSM < 1) { cudaGetLastError(); perSM = 1; }
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms < 1) { cudaGetLastError(); sms = 148; }
    return perSM * sms;
}

struct RenderCounters { unsigned long long v[8] = {0, 0, 0, 0, 0, 0, 0, 0}; };

// Round-1 wavefront: generate -> compact -> ONE bounce kernel per step, tail kernel at the end (kept for A/B measurements,
// GDB200_GPT_FUSED_BOUNCE).
int renderFused(gdb200_scene *s, const GptArgs &a, Marks &marks, gdb200_stats *stats, int &launches, int spp, RenderCounters &hc)
{
    const int nSlots = a.nSlots;
    gpt_init_kernel<<<(nSlots + 255) / 256, 256>>>(a);
    launches = 1;
    const int genBlocks = (nSlots + kGenThreads - 1) / kGenThreads;
    const int bounceBlocks = (nSlots + 32 * kBuckets + kBounceThreads - 1) / kBounceThreads;
    int parity = 0;
    unsigned long long tailThreshold = (unsigned long long)std::max(nSlots / 16, std::min(nSlots, 16384));
    if (getenv("GDB200_NO_TAIL")) tailThreshold = 0;        // test knob: drain the wavefront through its queues to the last path
    const long long maxSteps = (long long)spp * 4096 + 65536;     // safety net: never spin forever
    size_t firstMark = marks.used;
    for (long long step = 0;; step++) {
        if (step > maxSteps) return set_error(GDB200_ERR_CUDA, "wavefront did not drain after %lld steps", step);
        marks.mark();
        gpt_generate_kernel<<<genBlocks, kGenThreads>>>(a, parity);
        marks.mark();
        gpt_compact_kernel<<<(nSlots + 255) / 256, 256>>>(a, parity);
        marks.mark();
        gpt_bounce_kernel<2><<<bounceBlocks, kBounceThreads>>>(a, parity); launches += 3;
        marks.mark();
        parity ^= 1;
        if ((step & 15) == 15) {
            GDB_CUDA(cudaMemcpy(hc.v, s->counters, sizeof(hc.v), cudaMemcpyDeviceToHost));
            if (hc.v[0] >= (unsigned long long)nSlots) break;
            if (s->cancel) return set_error(GDB200_ERR_CANCELLED, "render cancelled");
            // tail: few pixel streams left => finish them in one launch instead of 3 launches per bounce
            if ((unsigned long long)nSlots - hc.v[0] <= tailThreshold) {
                gpt_tail_kernel<<<(nSlots + kBounceThreads - 1) / kBounceThreads, kBounceThreads>>>(a);
                launches++;
                break;
            }
        }
    }
    GDB_CUDA(cudaDeviceSynchronize());
    if (stats && marks.on)
        for (size_t i = firstMark; i + 3 < marks.used; i += 4) {
            stats->generate_ms += marks.between(i); stats->compact_ms += marks.between(i + 1); stats->bounce_ms += marks.between(i + 2);
            stats->bounce_launches++;
        }
    return GDB200_OK;
}

// Staged wavefront (gpt_stages.cuh): one tick = compact(A) -> primary, shade, resolve -> compact(B) -> prepare, generate -> casts.
int renderStaged(gdb200_scene *s, const GptArgs &a, Marks &marks, gdb200_stats *stats, int &launches, int spp, RenderCounters &hc)
{
    const int nSlots = a.nSlots;
    struct Grids { int primary, shade[3], resolve, prepare, generate, castNearest, castAny; };
    static Grids grids[kMaxDevices];                 // per device, filled on first use (guarded by the device mutex)
    Grids &g = grids[s->device];
    if (!g.primary) {
        g.primary = persistentGrid(gpt_stage_kernel<SK_PRIMARY>, kStageThreads, s->device);
        g.shade[0] = persistentGrid(gpt_stage_kernel<SK_SHADE0>, kStageThreads, s->device);
        g.shade[1] = persistentGrid(gpt_stage_kernel<SK_SHADE1>, kStageThreads, s->device);
        g.shade[2] = persistentGrid(gpt_stage_kernel<SK_SHADE2>, kStageThreads, s->device);
        g.resolve = persistentGrid(gpt_stage_kernel<SK_RESOLVE>, kStageThreads, s->device);
        g.prepare = persistentGrid(gpt_stage_kernel<SK_PREPARE>, kStageThreads, s->device);
        g.generate = persistentGrid(gpt_stage_kernel<SK_GENERATE>, kStageThreads, s->device);
        g.castNearest = persistentGrid(gpt_cast_kernel<false>, kCastThreads, s->device);
        g.castAny = persistentGrid(gpt_cast_kernel<true>, kCastThreads, s->device);
    }
    const int compactBlocks = (nSlots + 255) / 256;
    gpt_stage_init_kernel<<<(std::max(nSlots, kStageBuckets) + 255) / 256, 256>>>(a);
    launches = 1;
    // a sample takes 2 ticks + (1 or 2) per bounce; streams are consumed sequentially by their slot
    const long long maxTicks = ((long long)spp * 8192 + 131072) * std::max(1, a.nStreams / std::max(1, nSlots) + 1);
    // phase times: the kernel that opens a phase stamps the GPU timer (stampPhase), 8 stamps per tick + one at the end
    Workspace &ws = g_workspace[s->device];
    const bool timing = stats && marks.on;
    if (timing && !ws.stamps && cudaMalloc(&ws.stamps, sizeof(unsigned long long) * kStampCapacity) != cudaSuccess) { cudaGetLastError(); ws.stamps = nullptr; }
    size_t nStamps = 0;
    bool stampTick = false, openPhase = false;       // openPhase: the last stamped tick's cast phase still waits for its closing stamp
    GptArgs b = a;                                   // b.mark set: this launch opens a phase
    auto opens = [&]() -> const GptArgs & { b.mark = stampTick ? ws.stamps + nStamps++ : nullptr; return b; };
    const char *printTick = getenv("GDB200_PRINT_QUEUES");      // profiling aid: queue lengths of one tick on stderr (pairs an ncu capture of that tick with its work)
    for (long long tick = 0;; tick++) {
        if (tick > maxTicks) return set_error(GDB200_ERR_CUDA, "wavefront did not drain after %lld ticks", tick);
        stampTick = timing && ws.stamps && nStamps + 9 < kStampCapacity;
        if (!stampTick && openPhase) { b.mark = ws.stamps + nStamps++; openPhase = false; gpt_stage_compact_kernel<0><<<compactBlocks, 256>>>(b); }
        else gpt_stage_compact_kernel<0><<<compactBlocks, 256>>>(opens());
        if (stampTick) openPhase = true;
        if (printTick && tick == atoll(printTick)) {
            int q[kStageBuckets], r[2];
            cudaMemcpy(q, a.qCount, sizeof(q), cudaMemcpyDeviceToHost);
            int shade[3] = {0, 0, 0};
            for (int st = 0; st < 3; st++) for (int t = 0; t < kBsdfTypes; t++) shade[st] += q[QA_SHADE0 + st * kBsdfTypes + t];
            fprintf(stderr, "gdb200 tick %lld: primary %d shade0 %d shade1 %d shade2 %d resolve %d (slots %d)\n", tick, q[QA_PRIMARY], shade[0], shade[1], shade[2], q[QA_RESOLVE], nSlots);
            (void)r;
        }
        gpt_stage_kernel<SK_PRIMARY><<<g.primary, kStageThreads>>>(opens());
        gpt_stage_kernel<SK_SHADE0><<<g.shade[0], kStageThreads>>>(opens());
        gpt_stage_kernel<SK_SHADE1><<<g.shade[1], kStageThreads>>>(a);
        gpt_stage_kernel<SK_SHADE2><<<g.shade[2], kStageThreads>>>(a);
        gpt_stage_kernel<SK_RESOLVE><<<g.resolve, kStageThreads>>>(opens());
        gpt_stage_compact_kernel<1><<<compactBlocks, 256>>>(opens());
        gpt_stage_kernel<SK_PREPARE><<<g.prepare, kStageThreads>>>(opens());
        gpt_stage_kernel<SK_GENERATE><<<g.generate, kStageThreads>>>(opens());
        gpt_cast_kernel<false><<<g.castNearest, kCastThreads>>>(opens());
        gpt_cast_kernel<true><<<g.castAny, kCastThreads>>>(a);
        launches += 11;
        if ((tick & 15) == 15) {
            GDB_CUDA(cudaMemcpy(hc.v, s->counters, sizeof(hc.v), cudaMemcpyDeviceToHost));
            if (hc.v[7]) return set_error(GDB200_ERR_CUDA, "ray queue overflow (%llu rays dropped)", hc.v[7]);
            if (hc.v[0] >= (unsigned long long)nSlots) break;
            if (s->cancel) return set_error(GDB200_ERR_CANCELLED, "render cancelled");
        }
    }
    if (openPhase) { b.mark = ws.stamps + nStamps++; gpt_stamp_kernel<<<1, 1>>>(b); launches++; }      // closes the last tick's cast phase
    GDB_CUDA(cudaDeviceSynchronize());
    if (timing && nStamps % 8 == 1) {
        ws.stampsHost.resize(nStamps);
        GDB_CUDA(cudaMemcpy(ws.stampsHost.data(), ws.stamps, sizeof(unsigned long long) * nStamps, cudaMemcpyDeviceToHost));
        const unsigned long long *t = ws.stampsHost.data();
        double ns[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // compact A, primary, shade, resolve, compact B, prepare, generate, casts
        for (size_t i = 0; i + 8 < nStamps; i += 8) {
            for (int k = 0; k < 8; k++) ns[k] += (double)(t[i + k + 1] - t[i + k]);
            stats->bounce_launches++;
        }
        stats->compact_ms = (ns[0] + ns[4]) * 1e-6; stats->primary_ms = ns[1] * 1e-6; stats->bounce_ms = ns[2] * 1e-6;
        stats->resolve_ms = ns[3] * 1e-6; stats->prepare_ms = ns[5] * 1e-6; stats->generate_ms = ns[6] * 1e-6;
        stats->cast_ms = ns[7] * 1e-6;
    }
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
    if (s->device >= 0 && s->device < kMaxDevices) {
        std::lock_guard<std::mutex> lock(g_deviceMutex[s->device]);
        std::vector<Workspace::FilmSet> &spare = g_workspace[s->device].spareFilms;
        for (size_t i = 0; i < spare.size(); i++)
            if (spare[i].pixels == n) {
                s->film = spare[i].film; s->dev64 = spare[i].dev64; s->dev32 = spare[i].dev32; s->counters = spare[i].counters;
                spare.erase(spare.begin() + i);
                break;
            }
    }
    cudaError_t e = cudaSuccess;
    if (!s->film) {
        e = cudaMalloc(&s->film, sizeof(double) * 5 * n * 4);
        if (e == cudaSuccess) e = cudaMalloc(&s->dev64, sizeof(double) * 5 * n * 3);
        if (e == cudaSuccess) e = cudaMalloc(&s->dev32, sizeof(float) * 5 * n * 3);
        if (e == cudaSuccess) e = cudaMalloc(&s->counters, sizeof(unsigned long long) * 8);
    }
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
    int current = 0, n = 0;
    if (cudaGetDevice(&current) != cudaSuccess || cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return; }
    for (int d = 0; d < n && d < kMaxDevices; d++) {
        std::lock_guard<std::mutex> lock(g_deviceMutex[d]);
        if (g_workspace[d].slotCapacity || !g_workspace[d].events.empty() || !g_workspace[d].spareFilms.empty()) { cudaSetDevice(d); freeWorkspace(g_workspace[d]); }
    }
    cudaSetDevice(current);
}

int gdb200_gpt_render(gdb200_scene *s, const gdb200_gpt_params *p, gdb200_buffers *out, gdb200_stats *stats)
{
    if (!s || !p) return set_error(GDB200_ERR_ARGUMENT, "scene/params is NULL");
    GptArgs a;
    const char *capEnv = getenv("GDB200_MAX_SLOTS");        // developer knob; gdb200_gpt_params.max_slots is the interface
    if (int rc = setupArgs(*s, p, a, capEnv ? atoi(capEnv) : (1 << 23))) return rc;
    DeviceGuard guard;
    if (int rc = guard.bind(s->device)) return rc;
    const int nSlots = a.nSlots;
    const bool fused = (p->flags & GDB200_GPT_FUSED_BOUNCE) != 0;

    std::lock_guard<std::mutex> lock(g_deviceMutex[s->device]);
    classifyMaterials(s, p->shift_threshold);
    Workspace &ws = g_workspace[s->device];
    if (int rc = ensureWorkspace(ws, nSlots, fused)) return rc;
    if (int rc = uploadScene(s)) return rc;
    GDB_CUDA(cudaMemset(s->film, 0, sizeof(double) * 5 * (size_t)s->width * s->height * 4));
    GDB_CUDA(cudaMemset(s->counters, 0, sizeof(unsigned long long) * 8));

    a.sd = ws.sd; a.si = ws.si;
    a.film = s->film; a.liveList = ws.liveList; a.liveCount = ws.liveCount; a.genList = ws.genList; a.genCount = ws.genCount;
    a.counters = s->counters;
    a.rays[0] = ws.rays[0]; a.rays[1] = ws.rays[1]; a.rayOwner[0] = ws.rayOwner[0]; a.rayOwner[1] = ws.rayOwner[1];
    a.rayCount = ws.rayCount; a.rayCapacity = 5 * nSlots; a.qKey = ws.qKey; a.qList = ws.qList; a.qCount = ws.qCount;

    if (stats) memset(stats, 0, sizeof(*stats));
    Marks marks(ws, stats != nullptr);
    marks.mark();                                    // [0]: start of the device work
    s->cancel = 0;
    int launches = 0;
    RenderCounters hc;
    if (int rc = fused ? renderFused(s, a, marks, stats, launches, p->spp, hc) : renderStaged(s, a, marks, stats, launches, p->spp, hc)) {
        cudaDeviceSynchronize(); cudaGetLastError();
        return rc;
    }
    GDB_CUDA(cudaGetLastError());
    float ms = 0.f;
    if (marks.on && marks.used >= 1) {               // events[0] = start of the device work; the staged path records no others
        marks.mark();
        GDB_CUDA(cudaEventSynchronize(ws.events[marks.used - 1]));
        GDB_CUDA(cudaEventElapsedTime(&ms, ws.events[0], ws.events[marks.used - 1]));
    }
    GDB_CUDA(cudaMemcpy(hc.v, s->counters, sizeof(hc.v), cudaMemcpyDeviceToHost));
    if (int rc = developAndCopy(s, out)) return rc;
    if (stats) {
        stats->device_ms = ms; stats->launches = launches + 1;
        stats->samples = (double)hc.v[3]; stats->rays = (double)hc.v[1]; stats->path_vertices = (double)hc.v[2];
        stats->state_bytes = (double)hc.v[4]; stats->path_bounces = (double)hc.v[5];
    }
    return GDB200_OK;
}

int gdb200_debug_check_culling(gdb200_scene *s, int n_rays, unsigned long long seed, unsigned long long *out_mismatches, unsigned long long *out_hits)
{
    if (!s || !out_mismatches || n_rays <= 0) return set_error(GDB200_ERR_ARGUMENT, "scene/out is NULL");
    DeviceGuard guard;
    if (int rc = guard.bind(s->device)) return rc;
    std::lock_guard<std::mutex> lock(g_deviceMutex[s->device]);
    classifyMaterials(s, 0.001);
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
    DeviceGuard guard;
    if (int rc = guard.bind(s->device)) return rc;
    return developAndCopy(s, out);
}

}  // extern "C"
