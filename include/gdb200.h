/* gdb200 — C ABI of the B200-native G-PT + screened-Poisson hot path.
 *
 * This is the drop-in boundary: plain C, caller-owned buffers,
 * status codes, no exceptions, no torch/C++ types.  A Mitsuba `gpt` integrator
 * plugin shim calls these instead of
 *   - GradientPathIntegrator::render's block scheduler + renderBlock
 *     (reference src/integrators/gpt/gpt.cpp:1358-1413, 1220-1355), and
 *   - poisson::Solver::{importImagesMTS,setupBackend,solveIndirect,
 *     exportImagesMTS} (reference src/integrators/poisson_solver/Solver.hpp:113-117,
 *     call site gpt.cpp:1445-1462).
 *
 * Image buffers: row-major, top-left origin, interleaved RGB (= Vec3f /
 * Spectrum AoS, Solver.cpp:224-227).  Every function returns 0 on success;
 * on failure the message is available from gdb200_last_error() (thread-local).
 * All functions fail (never fall back to a CPU path) when no CUDA device or
 * kernel image is available.
 */
#ifndef GDB200_H
#define GDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDB200_OK            0
#define GDB200_ERR_ARGUMENT  1
#define GDB200_ERR_CUDA      2
#define GDB200_ERR_NO_DEVICE 3
#define GDB200_ERR_CANCELLED 4

/* ------------------------------------------------------------------ misc */

int         gdb200_version(void);                 /* 100*major + minor */
const char *gdb200_last_error(void);              /* thread-local, never NULL */
int         gdb200_device_count(int *out_count);
int         gdb200_set_device(int device);
/* Pinned host allocation helpers so callers can stage film buffers for full
 * PCIe/NVLink-C2C copy bandwidth. */
int         gdb200_host_alloc(void **out_ptr, size_t bytes);
int         gdb200_host_free(void *ptr);

typedef struct gdb200_stats {
    double device_ms;       /* CUDA-event time of the device work of this call   */
    double h2d_ms, d2h_ms;  /* copies performed by host-pointer entry points     */
    int    launches;        /* kernels of this library launched by this call     */
    int    irls_iters;      /* IRLS iterations executed (solver)                 */
    int    cg_iters;        /* total CG iterations executed (solver)             */
    int    reserved0;
    double samples;         /* evaluatePoint() samples traced (tracer)           */
    double rays;            /* rays cast (tracer)                                */
    double path_vertices;   /* sum of base-path depths (tracer; "Average path length", gpt.cpp:1178-1179) */
    double state_bytes;     /* wavefront state bytes moved through HBM (tracer)  */
} gdb200_stats;

/* ------------------------------------------------ screened Poisson solver */

/* Mirrors poisson::Solver::Params' solver configuration (Solver.hpp:82-90). */
typedef struct gdb200_poisson_config {
    int   irlsIterMax;
    float irlsRegInit;
    float irlsRegIter;
    int   cgIterMax;
    int   cgIterCheck;
    float cgTolerance;
} gdb200_poisson_config;

/* Solver::Params::setConfigPreset (Solver.cpp:90-164): "L1D","L1Q","L1L","L2D","L2Q". */
int gdb200_poisson_preset(const char *preset, gdb200_poisson_config *out_cfg);

typedef struct gdb200_poisson_plan gdb200_poisson_plan;

/* Allocates the device workspace for a w x h solve on the current device
 * (replaces Solver::setupBackend's allocVector calls, Solver.cpp:296-313). */
int  gdb200_poisson_plan_create(int w, int h, gdb200_poisson_plan **out_plan);
void gdb200_poisson_plan_destroy(gdb200_poisson_plan *plan);

/* Device-pointer entry: inputs/outputs are resident in HBM (w*h*3 floats each).
 * d_throughput may be NULL (alpha := 0, x0 := 0; Solver.cpp:319,334-337) and
 * d_direct may be NULL (final := x; Solver.cpp:561-562).  `stream` is a
 * cudaStream_t (NULL = default stream).  Asynchronous unless stats != NULL. */
int gdb200_poisson_solve_device(gdb200_poisson_plan *plan,
                                const float *d_dx, const float *d_dy,
                                const float *d_throughput, const float *d_direct,
                                float alpha, const gdb200_poisson_config *cfg,
                                float *d_out_final, void *stream, gdb200_stats *stats);

/* Host-pointer entry = importImagesMTS + setupBackend + solveIndirect +
 * exportImagesMTS (gpt.cpp:1456-1462) in one call.  Copies in, solves, copies out. */
int gdb200_poisson_solve(const float *dx, const float *dy, const float *throughput,
                         const float *direct, int w, int h, float alpha,
                         const char *preset, float *out_final, gdb200_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* GDB200_H */
