// Microbenchmark: how fast can a stage kernel's access pattern move path-slot state on a B200?
// State = [slot][64 records] x 32 B (2 KB per slot), a queue holds every ~3rd slot in ascending order (what the tracer's
// compaction produces).  Per queue entry a "stage" reads RD records and writes WR records of the slot's block.
//   A  one thread per slot, all loads issued back to back (the best case of the shipped per-thread pattern)
//   A2 one thread per slot, the record groups loaded one after the other with a dependent fp64 chain in between (a stage's shape)
//   B  the CTA fetches each slot's record groups with one bulk copy per group (cp.async.bulk, 384-byte bursts) into shared
//      memory, double buffered; threads read their slot from shared memory; stores per thread as in A
//   C  same bytes streamed contiguously (the copy regime)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a slot_access.cu -o slot_access ; ./slot_access [slots_log2=22]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

static constexpr int kRecB = 32, kRecsPerSlot = 64, kSlotB = kRecB * kRecsPerSlot;
static constexpr int kGroups = 5, kGroupRecs = 12;       // base + 4 offsets, 12 records (384 B) each
static constexpr int kRdPerGroup = 8, kWrPerGroup = 6;   // 40 records read (1280 B), 30 written (960 B) per slot

struct __align__(32) Rec { double v[4]; };

__device__ __forceinline__ Rec ldrec(const Rec *p)
{
    Rec r;
    asm volatile("ld.global.v2.f64 {%0,%1}, [%4]; ld.global.v2.f64 {%2,%3}, [%4+16];"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
    return r;
}
__device__ __forceinline__ void strec(Rec *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v2.f64 [%0], {%1,%2}; st.global.v2.f64 [%0+16], {%3,%4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

template <int CHAIN> __device__ __forceinline__ double chain(double x)
{
#pragma unroll 8
    for (int i = 0; i < CHAIN; i++) x = fma(x, 0.999999, 1e-9);
    return x;
}

// ---- A / A2
template <int CHAIN, bool WRITE> __global__ void __launch_bounds__(128, 2) perThread(Rec *state, const int *queue, int n, double *sink)
{
    double acc = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        Rec *blk = state + (size_t)queue[q] * kRecsPerSlot;
        double g[kGroups];
#pragma unroll
        for (int k = 0; k < kGroups; k++) {
            double s = CHAIN ? acc * 1e-30 : 0;       // CHAIN: the next group's loads cannot issue before this chain ends
            Rec *grp = blk + k * kGroupRecs + (CHAIN ? (int)(s != s) : 0);
            Rec r[kRdPerGroup];
#pragma unroll
            for (int i = 0; i < kRdPerGroup; i++) r[i] = ldrec(grp + i);
#pragma unroll
            for (int i = 0; i < kRdPerGroup; i++) s += r[i].v[0] + r[i].v[1] + r[i].v[2] + r[i].v[3];
            g[k] = chain<CHAIN>(s);
            acc += g[k];
        }
        if (WRITE) {
#pragma unroll
            for (int k = 0; k < kGroups; k++)
#pragma unroll
                for (int i = 0; i < kWrPerGroup; i++) strec(blk + k * kGroupRecs + i, g[k], acc, i, k);
        }
    }
    if (acc == 1.2345) *sink = acc;
}

// ---- B
__device__ __forceinline__ uint32_t smemAddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t *b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(b)), "r"(count)); }
__device__ __forceinline__ void mbarExpect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemAddr(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbarWait(uint64_t *b, uint32_t phase)
{
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" :: "r"(smemAddr(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulkLoad(void *dst, const void *src, uint32_t bytes, uint64_t *b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(b)) : "memory");
}

// One CTA iteration = T slots; each thread issues the kGroups bulk copies of ITS slot (kRdPerGroup records each) into its row
// of the stage buffer; two buffers.
template <int T, int CHAIN, bool WRITE> __global__ void __launch_bounds__(T) bulkStaged(Rec *state, const int *queue, int n, double *sink)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kRowRecs = kGroups * kRdPerGroup;
    Rec *buf[2] = { reinterpret_cast<Rec *>(smem), reinterpret_cast<Rec *>(smem) + T * kRowRecs };
    __shared__ uint64_t bar[2];
    if (threadIdx.x == 0) { mbarInit(&bar[0], T); mbarInit(&bar[1], T); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    double acc = 0;
    const int stride = gridDim.x * T;
    auto issue = [&](int base, int b) {
        const int q = base + threadIdx.x;
        if (q < n) {
            const Rec *blk = state + (size_t)queue[q] * kRecsPerSlot;
            mbarExpect(&bar[b], kRowRecs * kRecB);
#pragma unroll
            for (int k = 0; k < kGroups; k++)
                bulkLoad(buf[b] + threadIdx.x * kRowRecs + k * kRdPerGroup, blk + k * kGroupRecs, kRdPerGroup * kRecB, &bar[b]);
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smemAddr(&bar[b])) : "memory");
        }
    };
    int base = blockIdx.x * T, it = 0;
    if (base < n) issue(base, 0);
    for (; base < n; base += stride, it++) {
        const int b = it & 1;
        if (base + stride < n) issue(base + stride, b ^ 1);
        mbarWait(&bar[b], (it >> 1) & 1);
        const int q = base + threadIdx.x;
        if (q < n) {
            Rec *blk = state + (size_t)queue[q] * kRecsPerSlot;
            const Rec *row = buf[b] + threadIdx.x * kRowRecs;
            double g[kGroups];
#pragma unroll
            for (int k = 0; k < kGroups; k++) {
                double s = 0;
#pragma unroll
                for (int i = 0; i < kRdPerGroup; i++) { const Rec r = row[k * kRdPerGroup + i]; s += r.v[0] + r.v[1] + r.v[2] + r.v[3]; }
                g[k] = chain<CHAIN>(s);
                acc += g[k];
            }
            if (WRITE) {
#pragma unroll
                for (int k = 0; k < kGroups; k++)
#pragma unroll
                    for (int i = 0; i < kWrPerGroup; i++) strec(blk + k * kGroupRecs + i, g[k], acc, i, k);
            }
        }
        __syncthreads();   // everybody is done with buf[b] before it is refilled two iterations on
    }
    if (acc == 1.2345) *sink = acc;
}

// ---- C
template <bool WRITE> __global__ void __launch_bounds__(256) stream(Rec *state, size_t nRd, size_t nWr, double *sink)
{
    double acc = 0;
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    for (size_t i = t; i < nRd; i += step) { const Rec r = ldrec(state + i); acc += r.v[0] + r.v[3]; }
    if (WRITE) for (size_t i = t; i < nWr; i += step) strec(state + i, acc, 1, 2, 3);
    if (acc == 1.2345) *sink = acc;
}

template <class F> static float timeIt(F f, int reps = 3)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    return best;
}

int main(int argc, char **argv)
{
    const int lg = argc > 1 ? atoi(argv[1]) : 22;
    const size_t nSlots = (size_t)1 << lg;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    Rec *state; cudaMalloc(&state, nSlots * kSlotB); cudaMemset(state, 0, nSlots * kSlotB);
    double *sink; cudaMalloc(&sink, 8);
    for (int every = 1; every <= 9; every += (every == 1 ? 2 : 3)) {
        std::vector<int> q; q.reserve(nSlots / every + 16);
        uint64_t s = 88172645463325252ull;
        for (size_t i = 0; i < nSlots; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; if (s % every == 0) q.push_back((int)i); }
        const int n = (int)q.size();
        int *dq; cudaMalloc(&dq, sizeof(int) * n); cudaMemcpy(dq, q.data(), sizeof(int) * n, cudaMemcpyHostToDevice);
        const double rdB = (double)n * kGroups * kRdPerGroup * kRecB, wrB = (double)n * kGroups * kWrPerGroup * kRecB;
        printf("\n%s: %zu slots (%.1f GB), queue holds every ~%d-th: %d entries, %.0f B read + %.0f B written per entry\n", p.name, nSlots,
               nSlots * (double)kSlotB / 1e9, every, n, rdB / n, wrB / n);
        auto report = [&](const char *name, float ms, double bytes) { printf("  %-58s %8.3f ms  %7.0f GB/s\n", name, ms, bytes / ms * 1e-6); };
        report("A  per-thread loads, independent, read only", timeIt([&] { perThread<0, false><<<sms * 8, 128>>>(state, dq, n, sink); }), rdB);
        report("A  per-thread loads, independent, read+write", timeIt([&] { perThread<0, true><<<sms * 8, 128>>>(state, dq, n, sink); }), rdB + wrB);
        report("A2 per-thread loads, group by group (256 DFMA chain), r", timeIt([&] { perThread<256, false><<<sms * 8, 128>>>(state, dq, n, sink); }), rdB);
        report("A2 per-thread loads, group by group (256 DFMA chain), r+w", timeIt([&] { perThread<256, true><<<sms * 8, 128>>>(state, dq, n, sink); }), rdB + wrB);
        {
            constexpr int T = 64; const int shm = 2 * T * kGroups * kRdPerGroup * kRecB;
            cudaFuncSetAttribute(bulkStaged<T, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm);
            cudaFuncSetAttribute(bulkStaged<T, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm);
            cudaFuncSetAttribute(bulkStaged<T, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm);
            int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bulkStaged<T, 0, true>, T, shm);
            printf("  (B: %d threads, %d B shared per CTA, %d CTAs/SM)\n", T, shm, occ);
            report("B  bulk copies per slot group into shared, read only", timeIt([&] { bulkStaged<T, 0, false><<<sms * occ, T, shm>>>(state, dq, n, sink); }), rdB);
            report("B  bulk copies per slot group into shared, read+write", timeIt([&] { bulkStaged<T, 0, true><<<sms * occ, T, shm>>>(state, dq, n, sink); }), rdB + wrB);
            report("B2 bulk copies + 256 DFMA chain per group, read+write", timeIt([&] { bulkStaged<T, 256, true><<<sms * occ, T, shm>>>(state, dq, n, sink); }), rdB + wrB);
        }
        report("C  same bytes streamed, read only", timeIt([&] { stream<false><<<sms * 8, 256>>>(state, (size_t)(rdB / kRecB), 0, sink); }), rdB);
        report("C  same bytes streamed, read+write", timeIt([&] { stream<true><<<sms * 8, 256>>>(state, (size_t)(rdB / kRecB), (size_t)(wrB / kRecB), sink); }), rdB + wrB);
        cudaFree(dq);
    }
    return 0;
}
