// Test infrastructure: thin extern "C" entry over the UNMODIFIED reference
// solver (src/integrators/poisson_solver/Solver.{hpp,cpp}), built by
// oracle/Makefile into oracle/_ref/.  Mirrors the call sequence of
// gpt.cpp:1445-1462.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the resulting library.
#include "Solver.hpp"
#include <string>

// backend: "Auto" (what gpt.cpp leaves), "OpenMP", "Naive" or -- in the build with the reference's CUDA backend compiled in
// (oracle/_ref/libref_poisson_cuda.so) -- "CUDA" (Solver.cpp:262-294).
extern "C" int ref_poisson_solve_backend(const float *dx, const float *dy, const float *throughput,
                                         const float *direct, int w, int h, float alpha,
                                         const char *preset, const char *backend, float *out_final, float *out_seconds)
{
    poisson::Solver::Params params;
    if (!params.setConfigPreset(preset)) return 1;
    params.alpha = alpha;
    params.backend = backend;
    static float s_seconds; s_seconds = -1.f;
    params.setLogFunction(poisson::Solver::Params::LogFunction([](const std::string &m) {
        float v; if (sscanf(m.c_str(), "Execution time = %f s", &v) == 1) s_seconds = v; }));
    poisson::Solver solver(params);
    solver.importImagesMTS(const_cast<float*>(dx), const_cast<float*>(dy),
                           const_cast<float*>(throughput), const_cast<float*>(direct), w, h);
    solver.setupBackend();
    solver.solveIndirect();
    solver.exportImagesMTS(out_final);
    if (out_seconds) *out_seconds = s_seconds;
    return 0;
}

extern "C" int ref_poisson_solve(const float *dx, const float *dy, const float *throughput,
                                 const float *direct, int w, int h, float alpha,
                                 const char *preset, float *out_final, float *out_seconds)
{
#ifdef REF_POISSON_HAS_CUDA
    return ref_poisson_solve_backend(dx, dy, throughput, direct, w, h, alpha, preset, "OpenMP", out_final, out_seconds);
#else
    return ref_poisson_solve_backend(dx, dy, throughput, direct, w, h, alpha, preset, "Auto", out_final, out_seconds);
#endif
}

// Solve, then the reference's own Solver::evaluateMetricsMTS (Solver.cpp:511-541): err = w*h*3, errL = {errL1, errL2}.
extern "C" int ref_poisson_metrics(const float *dx, const float *dy, const float *throughput,
                                   const float *direct, int w, int h, float alpha,
                                   const char *preset, float *out_final, float *out_err, float *out_errL)
{
    poisson::Solver::Params params;
    if (!params.setConfigPreset(preset)) return 1;
    params.alpha = alpha;
#ifdef REF_POISSON_HAS_CUDA
    params.backend = "OpenMP";
#endif
    params.setLogFunction(poisson::Solver::Params::LogFunction([](const std::string &) {}));
    poisson::Solver solver(params);
    solver.importImagesMTS(const_cast<float*>(dx), const_cast<float*>(dy),
                           const_cast<float*>(throughput), const_cast<float*>(direct), w, h);
    solver.setupBackend();
    solver.solveIndirect();
    solver.evaluateMetricsMTS(out_err, out_errL[0], out_errL[1]);
    solver.exportImagesMTS(out_final);
    return 0;
}
