/* TEST INFRASTRUCTURE — CPU oracle for the screened-Poisson reconstruction.
 *
 * A plain-C, single-threaded fp32 restatement of the reference algorithm
 *   src/integrators/poisson_solver/Solver.cpp:90-164   (presets)
 *   src/integrators/poisson_solver/Solver.cpp:296-337  (b, x initialisation)
 *   src/integrators/poisson_solver/Solver.cpp:374-509  (IRLS over CG)
 *   src/integrators/poisson_solver/Solver.cpp:561-580  (final = direct + x)
 *   src/integrators/poisson_solver/Backend.cpp:154-376 (the vector ops)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * it; the product (gdb200 CUDA library) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_poisson_oracle.py checks this file
 * bit-for-bit against the reference's own sources compiled unmodified
 * (oracle/_ref/libref_poisson.so, both built with -ffp-contract=off) and against
 * checksums committed under tests/golden/.
 *
 * Layout: interleaved RGB, row-major, top-left origin (= Vec3f AoS,
 * Solver.cpp:224-227).  All sums are sequential fp32, as in Backend.cpp.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int   irlsIterMax;
    float irlsRegInit, irlsRegIter;
    int   cgIterMax, cgIterCheck;
    float cgTolerance;
} oracle_preset;

/* Solver.cpp:90-164.  cgPrecond is false in every preset and is not restated. */
static int oracle_preset_lookup(const char *name, oracle_preset *p)
{
    p->irlsIterMax = 1; p->irlsRegInit = 0.f; p->irlsRegIter = 0.f;
    p->cgIterMax = 1; p->cgIterCheck = 100; p->cgTolerance = 0.f;
    if (!strcmp(name, "L1D")) { p->irlsIterMax = 20; p->irlsRegInit = 0.05f;  p->irlsRegIter = 0.5f; p->cgIterMax = 50;    return 1; }
    if (!strcmp(name, "L1Q")) { p->irlsIterMax = 64; p->irlsRegInit = 1.0f;   p->irlsRegIter = 0.7f; p->cgIterMax = 1000;  return 1; }
    if (!strcmp(name, "L1L")) { p->irlsIterMax = 7;  p->irlsRegInit = 1.0e-4f; p->irlsRegIter = 1.0e-1f; p->cgIterMax = 20000; p->cgTolerance = 1.0e-20f; return 1; }
    if (!strcmp(name, "L2D")) { p->cgIterMax = 50;  return 1; }
    if (!strcmp(name, "L2Q")) { p->cgIterMax = 500; return 1; }
    return 0;
}

static float maxf(float a, float b) { return a > b ? a : b; }

/* Reduction accumulators.  Faithful mode (default) rounds to fp32 after every add =
 * the reference's sequential `float` sums.  acc64 mode keeps fp64 sums and exists only so
 * tests can measure how far reduction order alone moves the result (the parity noise floor). */
static int g_acc64 = 0;
static inline void acc_add(double *acc, float v)
{
    if (g_acc64) *acc += (double)v;
    else *acc = (double)(float)((float)*acc + v);
}

/* e = b - P*x   (Backend.cpp:165-186 then :256-272 with a = -1) */
static void residual(float *e, const float *b, const float *x, int w, int h, float alpha)
{
    size_t n = (size_t)w * h;
    for (int yy = 0; yy < h; yy++)
        for (int xx = 0; xx < w; xx++) {
            size_t i = (size_t)yy * w + xx;
            for (int c = 0; c < 3; c++) {
                float xi = x[3 * i + c];
                float p0 = xi * alpha;
                float p1 = (xx != w - 1) ? x[3 * (i + 1) + c] - xi : 0.0f;
                float p2 = (yy != h - 1) ? x[3 * (i + w) + c] - xi : 0.0f;
                e[3 * (0 * n + i) + c] = -1.0f * p0 + b[3 * (0 * n + i) + c];
                e[3 * (1 * n + i) + c] = -1.0f * p1 + b[3 * (1 * n + i) + c];
                e[3 * (2 * n + i) + c] = -1.0f * p2 + b[3 * (2 * n + i) + c];
            }
        }
}

/* Backend.cpp:351-376 */
static void calc_w2(float *w2, const float *e, size_t n3, float reg)
{
    double sum = 0.0;
    for (size_t i = 0; i < n3; i++) {
        const float *v = e + 3 * i;
        float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        float wi = 1.0f / (len + reg);
        w2[i] = wi;
        acc_add(&sum, wi);
    }
    float coef = (float)(int)n3 / (float)sum;
    for (size_t i = 0; i < n3; i++) w2[i] *= coef;
}

/* r = P' diag(w2) e   (Backend.cpp:190-217) */
static void calc_PTW2x(float *r, const float *w2, const float *e, int w, int h, float alpha)
{
    size_t n = (size_t)w * h;
    for (int yy = 0; yy < h; yy++)
        for (int xx = 0; xx < w; xx++) {
            size_t i = (size_t)yy * w + xx;
            for (int c = 0; c < 3; c++) {
                float v = w2[0 * n + i] * e[3 * (0 * n + i) + c] * alpha;
                if (xx != 0)     v += w2[1 * n + i - 1] * e[3 * (1 * n + i - 1) + c];
                if (xx != w - 1) v -= w2[1 * n + i]     * e[3 * (1 * n + i) + c];
                if (yy != 0)     v += w2[2 * n + i - w] * e[3 * (2 * n + i - w) + c];
                if (yy != h - 1) v -= w2[2 * n + i]     * e[3 * (2 * n + i) + c];
                r[3 * i + c] = v;
            }
        }
}

/* Ap = P' W2 P p ; pAp = sum p*Ap   (Backend.cpp:221-252) */
static void calc_Ax_xAx(float *Ap, float pAp_out[3], const float *w2, const float *p, int w, int h, float alpha)
{
    size_t n = (size_t)w * h;
    float alphaSqr = alpha * alpha;
    double pAp[3] = {0.0, 0.0, 0.0};
    for (int yy = 0; yy < h; yy++)
        for (int xx = 0; xx < w; xx++) {
            size_t i = (size_t)yy * w + xx;
            for (int c = 0; c < 3; c++) {
                float xi = p[3 * i + c];
                float v = w2[0 * n + i] * xi * alphaSqr;
                if (xx != 0)     v += w2[1 * n + i - 1] * (xi - p[3 * (i - 1) + c]);
                if (xx != w - 1) v += w2[1 * n + i]     * (xi - p[3 * (i + 1) + c]);
                if (yy != 0)     v += w2[2 * n + i - w] * (xi - p[3 * (i - w) + c]);
                if (yy != h - 1) v += w2[2 * n + i]     * (xi - p[3 * (i + w) + c]);
                Ap[3 * i + c] = v;
                acc_add(&pAp[c], xi * v);
            }
        }
    for (int c = 0; c < 3; c++) pAp_out[c] = (float)pAp[c];
}

/* out_err / out_errL: optional outputs of Solver::evaluateMetricsMTS (Solver.cpp:511-541) on the solved x */
static int oracle_solve(const float *dx, const float *dy, const float *throughput,
                        const float *direct, int w, int h, float alpha,
                        const char *preset, float *out_final, float *out_err, float *out_errL)
{
    oracle_preset ps;
    if (!oracle_preset_lookup(preset, &ps) || w <= 0 || h <= 0 || !dx || !dy) return 1;
    /* Params::sanitize, Solver.cpp:168-178 */
    alpha = maxf(alpha, 0.0f);
    if (!throughput) alpha = 0.0f;                    /* Solver.cpp:319 */

    size_t n = (size_t)w * h, n3 = 3 * n;
    float *b  = (float *)malloc(sizeof(float) * 3 * n3);
    float *e  = (float *)malloc(sizeof(float) * 3 * n3);
    float *w2 = (float *)malloc(sizeof(float) * n3);
    float *x  = (float *)malloc(sizeof(float) * n3);
    float *r  = (float *)malloc(sizeof(float) * n3);
    float *p  = (float *)malloc(sizeof(float) * n3);
    float *Ap = (float *)malloc(sizeof(float) * n3);
    if (!b || !e || !w2 || !x || !r || !p || !Ap) return 2;

    /* Solver.cpp:321-337 */
    for (size_t i = 0; i < n3; i++) {
        b[0 * n3 + i] = throughput ? throughput[i] * alpha : 0.0f;
        b[1 * n3 + i] = dx[i];
        b[2 * n3 + i] = dy[i];
        x[i] = throughput ? throughput[i] : 0.0f;
    }

    for (int irls = 0; irls < ps.irlsIterMax; irls++) {
        residual(e, b, x, w, h, alpha);                                   /* :386-387 */
        if (irls == 0)
            for (size_t i = 0; i < n3; i++) w2[i] = 1.0f;                 /* :392 */
        else
            calc_w2(w2, e, n3, ps.irlsRegInit * powf(ps.irlsRegIter, (float)(irls - 1))); /* :395-396 */

        float rzA[3], rzB[3], pAp[3];
        float *rz = rzA, *rz2 = rzB;
        calc_PTW2x(r, w2, e, w, h, alpha);                                /* :403 */
        double acc[3] = {0.0, 0.0, 0.0};                                  /* :404 */
        for (size_t i = 0; i < n; i++)
            for (int c = 0; c < 3; c++) acc_add(&acc[c], r[3 * i + c] * r[3 * i + c]);
        for (int c = 0; c < 3; c++) rz[c] = (float)acc[c];
        memcpy(p, r, sizeof(float) * n3);                                 /* :405 */

        for (int cg = 0;; cg++) {
            if (cg % ps.cgIterCheck == 0 || cg == ps.cgIterMax) {         /* :411-445 */
                float errL2W = rz[0] + rz[1] + rz[2];
                if (cg == ps.cgIterMax || errL2W <= ps.cgTolerance) break;
            }
            { float *t = rz; rz = rz2; rz2 = t; }                         /* :466 */
            calc_Ax_xAx(Ap, pAp, w2, p, w, h, alpha);                     /* :467 */
            float a[3], bb[3];
            for (int c = 0; c < 3; c++) a[c] = rz2[c] / maxf(pAp[c], FLT_MIN);
            acc[0] = acc[1] = acc[2] = 0.0;                               /* :468, Backend.cpp:296-321 */
            for (size_t i = 0; i < n; i++)
                for (int c = 0; c < 3; c++) {
                    float ri = r[3 * i + c] - Ap[3 * i + c] * a[c];
                    r[3 * i + c] = ri;
                    acc_add(&acc[c], ri * ri);
                }
            for (int c = 0; c < 3; c++) rz[c] = (float)acc[c];
            for (int c = 0; c < 3; c++) bb[c] = rz[c] / maxf(rz2[c], FLT_MIN);
            for (size_t i = 0; i < n; i++)                                /* :469, Backend.cpp:325-347 */
                for (int c = 0; c < 3; c++) {
                    float pi = p[3 * i + c];
                    x[3 * i + c] += pi * a[c];
                    p[3 * i + c] = r[3 * i + c] + pi * bb[c];
                }
        }
    }

    /* Solver.cpp:561-567: final = 1*direct + x, or x when there is no direct image */
    if (out_final)
        for (size_t i = 0; i < n3; i++)
            out_final[i] = direct ? 1.0f * direct[i] + x[i] : x[i];

    if (out_errL) {      /* Solver::evaluateMetricsMTS, Solver.cpp:511-541: e = b - P*x, mean |e| and mean |e|^2 over its 3n RGB elements */
        residual(e, b, x, w, h, alpha);
        float errL1 = 0.0f, errL2 = 0.0f;
        for (size_t i = 0; i < n3; i++) {
            const float *v = e + 3 * i;
            errL1 += sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);     /* length(), Defs.hpp */
            errL2 += v[0] * v[0] + v[1] * v[1] + v[2] * v[2];            /* lenSqr() */
        }
        out_errL[0] = errL1 / (float)(int)n3;
        out_errL[1] = errL2 / (float)(int)n3;
        if (out_err) memcpy(out_err, e, sizeof(float) * n3);             /* the primal block of e */
    }

    free(b); free(e); free(w2); free(x); free(r); free(p); free(Ap);
    return 0;
}

int gdb200_oracle_poisson_solve(const float *dx, const float *dy, const float *throughput,
                                const float *direct, int w, int h, float alpha,
                                const char *preset, float *out_final)
{
    return oracle_solve(dx, dy, throughput, direct, w, h, alpha, preset, out_final, NULL, NULL);
}

/* Solve, then Solver::evaluateMetricsMTS: out_err = w*h*3 (the primal block of b - P*x), out_errL = {errL1, errL2} */
int gdb200_oracle_poisson_metrics(const float *dx, const float *dy, const float *throughput,
                                  const float *direct, int w, int h, float alpha,
                                  const char *preset, float *out_final, float *out_err, float *out_errL)
{
    return oracle_solve(dx, dy, throughput, direct, w, h, alpha, preset, out_final, out_err, out_errL);
}

/* Same algorithm with exact (fp64) reduction sums: NOT the reference's arithmetic, only a
 * yardstick for reduction-order sensitivity. */
int gdb200_oracle_poisson_solve_acc64(const float *dx, const float *dy, const float *throughput,
                                      const float *direct, int w, int h, float alpha,
                                      const char *preset, float *out_final)
{
    g_acc64 = 1;
    int rc = gdb200_oracle_poisson_solve(dx, dy, throughput, direct, w, h, alpha, preset, out_final);
    g_acc64 = 0;
    return rc;
}

#ifdef __cplusplus
}
#endif
