// TEST INFRASTRUCTURE -- host emulation of the Poisson solver's device code (gradientdomain-mitsuba_b200/csrc/poisson.cu), so
// that the persistent IRLS/CG kernel, its four variants and the multi-GPU ("sharded") protocol can be run when no GPU is present
// (this container has none).  It is not a CPU backend: nothing in the product loads it, libgdb200.so fails loudly without a device.
//
// The SAME SOURCE is compiled for the host (GDB200_EMU): every CUDA thread is an OS thread, every CTA a process (function-local
// __shared__ arrays are then plain statics), every "GPU" a group of CTA processes with a process-shared grid barrier, and all
// "device memory" one shared mapping, so that the shards reach each other's halo rows and mailboxes through plain pointers --
// exactly what peer memory gives the real kernels.  __shfl_xor_sync is a real exchange between the 32 threads of a warp,
// __syncthreads a barrier of the CTA's 256 threads, grid.sync() that plus the barrier between the CTA processes of a GPU.
//
//   poisson_emu <in.bin> <out.bin>      (tests/test_poisson_emu.py writes and reads the files)
#define GDB200_EMU 1
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <vector>
#include <pthread.h>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include "../../include/gdb200.h"

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct float4 { float x, y, z, w; } __attribute__((aligned(16)));
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct EmuDim3 { unsigned x, y, z; };
static thread_local EmuDim3 threadIdx = {0, 0, 0};
static EmuDim3 blockIdx = {0, 0, 0}, gridDim = {1, 1, 1}, blockDim = {256, 1, 1};      // one CTA per process

constexpr int kEmuThreads = 256, kEmuWarps = kEmuThreads / 32;
struct EmuCta {
    pthread_barrier_t block, warp[kEmuWarps];
    double xchg[kEmuWarps][32];
    float4 *dynamicShared;
};
static EmuCta g_cta;
static pthread_barrier_t *g_gridBarrier = nullptr;      // process-shared: thread 0 of every CTA of this "GPU"

static inline void __syncthreads() { pthread_barrier_wait(&g_cta.block); }
static inline double __shfl_xor_sync(unsigned, double v, int o)
{
    const int w = (int)threadIdx.x >> 5, l = (int)threadIdx.x & 31;
    g_cta.xchg[w][l] = v;
    pthread_barrier_wait(&g_cta.warp[w]);
    const double r = g_cta.xchg[w][l ^ o];
    pthread_barrier_wait(&g_cta.warp[w]);
    return r;
}
static inline double __ldcg(const double *p) { return *(const volatile double *)p; }
static inline float __ldcg(const float *p) { return *(const volatile float *)p; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double *p, double v)
{
    unsigned long long *q = reinterpret_cast<unsigned long long *>(p), o = __atomic_load_n(q, __ATOMIC_RELAXED), n;
    double od;
    do { std::memcpy(&od, &o, 8); const double nd = od + v; std::memcpy(&n, &nd, 8); } while (!__atomic_compare_exchange_n(q, &o, n, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return od;
}
static inline void __nanosleep(unsigned) { sched_yield(); }
static inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
static inline unsigned long long emu_timer_ns()
{
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
static inline float4 *emu_dynamic_shared() { return g_cta.dynamicShared; }

namespace cooperative_groups {
struct grid_group {
    void sync() const
    {
        __syncthreads();
        if (threadIdx.x == 0 && gridDim.x > 1) pthread_barrier_wait(g_gridBarrier);
        __syncthreads();
    }
};
inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups

#include "../../gradientdomain-mitsuba_b200/csrc/poisson.cu"

using namespace gdb200;

// ---------------------------------------------------------------------------------------------------------------- harness
struct Header {
    int w, h, variant, nShards, ctasPerShard, skipRank;
    int bounds[kMaxRanks + 1];
    gdb200_poisson_config cfg;
    float alpha;
    int hasDirect;
};

struct ShardMem {
    float *planes; double *red; int *iters; Mail *mail; unsigned *status; pthread_barrier_t *grid;
    size_t planeElems;
};

static char *g_arena = nullptr;
static size_t g_used = 0, g_size = 0;
template <class T> static T *carve(size_t n)
{
    g_used = (g_used + 255) & ~(size_t)255;
    T *p = reinterpret_cast<T *>(g_arena + g_used);
    g_used += n * sizeof(T);
    if (g_used > g_size) { fprintf(stderr, "poisson_emu: arena too small\n"); exit(2); }
    return p;
}

typedef void (*KernelFn)(const PoissonArgs);
static KernelFn kernelFor(int variant, bool shard)
{
    switch (variant) {
    case 1: return shard ? poisson_irls_cg_kernel<1, true> : poisson_irls_cg_kernel<1, false>;
    case 2: return shard ? poisson_irls_cg_kernel<2, true> : poisson_irls_cg_kernel<2, false>;
    case 3: return shard ? poisson_irls_cg_kernel<3, true> : poisson_irls_cg_kernel<3, false>;
    default: return shard ? poisson_irls_cg_kernel<0, true> : poisson_irls_cg_kernel<0, false>;
    }
}

struct ThreadStart { KernelFn fn; const PoissonArgs *args; unsigned tid; };
static void *threadMain(void *p)
{
    const ThreadStart *s = static_cast<const ThreadStart *>(p);
    threadIdx.x = s->tid;
    s->fn(*s->args);
    return nullptr;
}

// One CTA = this process: 256 OS threads run the kernel.
static void runCta(KernelFn fn, const PoissonArgs &args, unsigned block, unsigned grid, pthread_barrier_t *gridBarrier)
{
    blockIdx.x = block; gridDim.x = grid; g_gridBarrier = gridBarrier;
    pthread_barrier_init(&g_cta.block, nullptr, kEmuThreads);
    for (int w = 0; w < kEmuWarps; w++) pthread_barrier_init(&g_cta.warp[w], nullptr, 32);
    g_cta.dynamicShared = static_cast<float4 *>(aligned_alloc(16, kResBytes));
    pthread_attr_t attr; pthread_attr_init(&attr); pthread_attr_setstacksize(&attr, 512 * 1024);
    std::vector<pthread_t> th(kEmuThreads);
    std::vector<ThreadStart> st(kEmuThreads);
    for (int t = 0; t < kEmuThreads; t++) {
        st[t] = ThreadStart{fn, &args, (unsigned)t};
        if (pthread_create(&th[t], &attr, threadMain, &st[t]) != 0) { fprintf(stderr, "poisson_emu: pthread_create failed\n"); _exit(3); }
    }
    for (int t = 0; t < kEmuThreads; t++) pthread_join(th[t], nullptr);
}

int main(int argc, char **argv)
{
    if (argc != 3) { fprintf(stderr, "usage: poisson_emu in.bin out.bin\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("in"); return 2; }
    Header hd;
    if (fread(&hd, sizeof(hd), 1, f) != 1) return 2;
    const int w = hd.w, h = hd.h, wp = (w + 3) & ~3, n = hd.nShards, G = hd.ctasPerShard;
    const size_t img = (size_t)w * h * 3;

    g_size = (size_t)64 << 20;
    for (int r = 0; r < n; r++) g_size += (size_t)wp * (hd.bounds[r + 1] - hd.bounds[r] + 2) * kPlanes * sizeof(float) + (1 << 16);
    g_size += 5 * img * sizeof(float);
    g_arena = static_cast<char *>(mmap(nullptr, g_size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
    if (g_arena == MAP_FAILED) { perror("mmap"); return 2; }

    float *in[4];
    for (int i = 0; i < 4; i++) {
        in[i] = carve<float>(img);
        if (fread(in[i], sizeof(float), img, f) != img) { fprintf(stderr, "poisson_emu: short input\n"); return 2; }
    }
    fclose(f);
    float *out = carve<float>(img);
    for (size_t i = 0; i < img; i++) out[i] = -777.0f;

    std::vector<ShardMem> mem(n);
    pthread_barrierattr_t battr; pthread_barrierattr_init(&battr); pthread_barrierattr_setpshared(&battr, PTHREAD_PROCESS_SHARED);
    for (int r = 0; r < n; r++) {
        ShardMem &m = mem[r];
        m.planeElems = (size_t)wp * (hd.bounds[r + 1] - hd.bounds[r] + 2);
        m.planes = carve<float>(m.planeElems * kPlanes);
        memset(m.planes, 0, m.planeElems * kPlanes * sizeof(float));
        m.red = carve<double>(2 * 3 * G);
        m.iters = carve<int>(2);
        m.mail = carve<Mail>(4 * kMaxRanks); memset(m.mail, 0, sizeof(Mail) * 4 * kMaxRanks);
        m.status = carve<unsigned>(2); m.status[0] = m.status[1] = 0;
        m.grid = carve<pthread_barrier_t>(1);
        pthread_barrier_init(m.grid, &battr, G);
    }

    // the arguments gdb200_poisson_solve_device builds (csrc/poisson.cu, host part)
    std::vector<PoissonArgs> args(n);
    for (int r = 0; r < n; r++) {
        PoissonArgs &a = args[r];
        memset(&a, 0, sizeof(a));
        a.W = w; a.H = h; a.Wp = wp; a.Gx = wp / 4;
        a.tilesX = (a.Gx + kTileGX - 1) / kTileGX;
        a.y0 = hd.bounds[r]; a.y1 = hd.bounds[r + 1]; a.rank = r; a.nRanks = n;
        a.nTiles = a.tilesX * ((a.y1 - a.y0 + kTileY - 1) / kTileY);
        a.aosVec = (w % 4 == 0);
        a.alpha = hd.alpha > 0.f ? hd.alpha : 0.f;
        a.cfg = hd.cfg;
        for (int i = 0; i < kPlanes; i++) a.plane[i] = mem[r].planes + mem[r].planeElems * i;
        a.in_dx = in[0]; a.in_dy = in[1]; a.in_thr = in[2]; a.in_direct = hd.hasDirect ? in[3] : nullptr; a.out_final = out;
        a.red = mem[r].red; a.iters = mem[r].iters;
        if (n > 1) {
            for (int q = 0; q < n; q++) a.mail[q] = mem[q].mail;
            if (r > 0) { a.peerUp = mem[r - 1].planes; a.peerUpRows = hd.bounds[r] - hd.bounds[r - 1]; a.peerUpElems = mem[r - 1].planeElems; }
            if (r < n - 1) { a.peerDown = mem[r + 1].planes; a.peerDownElems = mem[r + 1].planeElems; }
        }
        a.epochBase = 16;
        a.status = mem[r].status;
        if (G > a.nTiles) { fprintf(stderr, "poisson_emu: %d CTAs for %d tiles\n", G, a.nTiles); return 2; }
        if (res_tiles(hd.variant) && a.nTiles > res_tiles(hd.variant) * G) { fprintf(stderr, "poisson_emu: variant %d does not fit\n", hd.variant); return 2; }
    }

    std::vector<pid_t> kids;
    for (int r = 0; r < n; r++) {
        if (r == hd.skipRank) continue;                       // a peer that never starts: the others must give up, not hang
        for (int b = 0; b < G; b++) {
            const pid_t pid = fork();
            if (pid < 0) { perror("fork"); return 2; }
            if (pid == 0) {
                prctl(PR_SET_PDEATHSIG, SIGKILL);             // a CTA process never outlives the harness (e.g. a test's time-out)
                runCta(kernelFor(hd.variant, n > 1), args[r], (unsigned)b, (unsigned)G, mem[r].grid);
                _exit(0);
            }
            kids.push_back(pid);
        }
    }
    int failed = 0;
    for (pid_t pid : kids) {
        int st = 0;
        waitpid(pid, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) failed++;
    }
    if (failed) { fprintf(stderr, "poisson_emu: %d CTA processes failed\n", failed); return 3; }

    f = fopen(argv[2], "wb");
    if (!f) { perror("out"); return 2; }
    fwrite(out, sizeof(float), img, f);
    for (int r = 0; r < n; r++) { fwrite(mem[r].iters, sizeof(int), 2, f); fwrite(mem[r].status, sizeof(unsigned), 2, f); }
    fclose(f);
    return 0;
}
