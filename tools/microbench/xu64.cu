// Microbenchmark: throughput of the 64-bit "XU" operations the fp64 tracer leans on (division / sqrt seeds and
// fp64<->fp32/int conversions) against DFMA and the 32-bit MUFU, in warp-instructions per cycle per SM.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a xu64.cu -o xu64 ; ./xu64
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096
#define CHAINS 8

template <int OP> __device__ __forceinline__ double op(double x, double k)
{
    if (OP == 0) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r + k; }          // MUFU.RCP64H
    if (OP == 1) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r + k; }        // MUFU.RSQ64H
    if (OP == 2) { float f; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(x)); return (double)__float_as_int(f) * 0 + x + k + f; }   // F2F.F32.F64 (+F2F back)
    if (OP == 3) { return k / x; }                                                                                      // full IEEE division
    if (OP == 4) { return sqrt(x) + k; }                                                                               // full IEEE sqrt
    if (OP == 5) { return fma(x, k, 0.5); }                                                                            // DFMA
    if (OP == 6) { float f = __int_as_float(__double2hiint(x)); float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(f)); return __hiloint2double(__float_as_int(r), __double2loint(x)) + k; }   // MUFU.RCP (32-bit)
    if (OP == 7) { unsigned long long u = (unsigned long long)__double_as_longlong(x) >> 11; return (double)u * 1.1102230246251565e-16 + k; }   // I2F.F64.U64
    if (OP == 8) { int i; asm volatile("cvt.rmi.s32.f64 %0, %1;" : "=r"(i) : "d"(x)); return x + (double)(i & 1) * 0 + k + __int_as_float(i) * 0.0; } // F2I.F64.FLOOR (+I2F)
    return x;
}

template <int OP> __global__ void bench(double *out, double seed)
{
    double v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) v[c] = seed + threadIdx.x * 1e-3 + c;
    for (int i = 0; i < ITER; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) v[c] = op<OP>(v[c], 1.25);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char *name, double *d, int sms, double ghz)
{
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    bench<OP><<<blocks, threads>>>(d, 1.5); cudaDeviceSynchronize();
    cudaEventRecord(a);
    bench<OP><<<blocks, threads>>>(d, 1.5);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warpOps = (double)blocks * threads / 32 * ITER * CHAINS;
    printf("%-28s %8.3f ms  %8.4f op-warp-inst/cycle/SM (at %.3f GHz)\n", name, ms, warpOps / (ms * 1e-3 * ghz * 1e9) / sms, ghz);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    double *d; cudaMalloc(&d, sizeof(double) * p.multiProcessorCount * 8 * 256);
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    run<5>("DFMA", d, p.multiProcessorCount, ghz);
    run<0>("MUFU.RCP64H (+DADD)", d, p.multiProcessorCount, ghz);
    run<1>("MUFU.RSQ64H (+DADD)", d, p.multiProcessorCount, ghz);
    run<6>("MUFU.RCP fp32 (+DADD)", d, p.multiProcessorCount, ghz);
    run<2>("F2F.F32.F64 (+back, DADD)", d, p.multiProcessorCount, ghz);
    run<7>("I2F.F64.U64 (+DFMA)", d, p.multiProcessorCount, ghz);
    run<8>("F2I.F64.FLOOR (+DADD)", d, p.multiProcessorCount, ghz);
    run<3>("div.rn.f64", d, p.multiProcessorCount, ghz);
    run<4>("sqrt.rn.f64 (+DADD)", d, p.multiProcessorCount, ghz);
    return 0;
}
