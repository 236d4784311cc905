// TEST INFRASTRUCTURE — a minimal host shim for the CUDA constructs used by the tracer's per-slot
// device routines (gradientdomain-mitsuba_b200/csrc/gpt_device.cuh, gpt_kernels.cuh), so that
// tests/emu/gpt_emu.cpp can run the *same source* serially on the CPU, one "thread" at a time, and
// compare it with the oracle when no GPU is present (this container has none).  It is not a CPU
// backend: nothing in the product loads it, libgdb200.so fails loudly without a device.
#pragma once
#define GDB200_EMU 1
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __constant__ static
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__
#define CUDART_INF (std::numeric_limits<double>::infinity())

struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
#include <pthread.h>
struct EmuDim3 { unsigned x, y, z; };
static thread_local EmuDim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
static EmuDim3 blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

// Two modes.  Serial (emu_block == nullptr): one "thread" at a time, the collectives are identities.  Block mode (the
// wavefront test, gpt_emu.cpp): the threads of one CTA run as OS threads; __syncthreads is a barrier and __ballot_sync a
// block-wide exchange (valid where every thread of the CTA reaches it, which is how the kernels use it).  Warp-AGGREGATED
// helpers that key on __activemask() (appendGen, countWarp) see a one-lane mask, i.e. every thread is its own leader.
struct EmuBlock { pthread_barrier_t bar; int nThreads; unsigned char pred[1024]; };
static EmuBlock *emu_block = nullptr;

using std::min;
using std::max;
using std::isfinite;
inline double atomicAdd(double *p, double v)
{
    unsigned long long *q = reinterpret_cast<unsigned long long *>(p), o = __atomic_load_n(q, __ATOMIC_RELAXED), n;
    double od;
    do { std::memcpy(&od, &o, 8); const double nd = od + v; std::memcpy(&n, &nd, 8); } while (!__atomic_compare_exchange_n(q, &o, n, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return od;
}
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned __activemask() { return 1u << (threadIdx.x & 31); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <class T> inline T __shfl_sync(unsigned, T v, int) { return v; }
inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
inline void __syncthreads() { if (emu_block) pthread_barrier_wait(&emu_block->bar); }
inline unsigned __ballot_sync(unsigned, bool p)
{
    if (!emu_block) return p ? 1u : 0u;
    const int t = (int)threadIdx.x, w = t >> 5;
    emu_block->pred[t] = p;
    pthread_barrier_wait(&emu_block->bar);
    unsigned m = 0;
    for (int l = 0; l < 32 && w * 32 + l < emu_block->nThreads; l++) if (emu_block->pred[w * 32 + l]) m |= 1u << l;
    pthread_barrier_wait(&emu_block->bar);
    return m;
}
inline bool __any_sync(unsigned, bool p) { return p; }
inline void __syncwarp(unsigned = 0xffffffffu) {}
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __frcp_rn(float x) { return 1.0f / x; }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
inline double __ldg(const double *p) { return *p; }
inline float __ldg(const float *p) { return *p; }
inline int __ldg(const int *p) { return *p; }
