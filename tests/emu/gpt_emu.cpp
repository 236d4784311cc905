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
    sd.assign((size_t)4 * kRecords * a.nSlots, 0.0); si.assign((size_t)16 * a.nSlots, 0); film.assign(5 * n * 4, 0.0);
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
