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

// ---- single-plugin entry points of the DEVICE routines for tests/test_chisquare.py (the reference's own consistency
// test of BSDF::sample/eval/pdf and Emitter::sampleDirect/pdfDirect, src/tests/test_chisquare.cpp)
static void emuMaterial(const gdb200_material *m, HostScene &hs)
{
    memset(&hs.host, 0, sizeof(hs.host));
    hs.mats.assign(1, *m);
    classifyMaterials(&hs, 0.001);
}
extern "C" int gdb200_emu_bsdf_sample_batch(const gdb200_material *m, const double *wi, int n, const double *samples,
                                            double *wo, double *weight, double *pdf, int *sampledType)
{
    static HostScene hs; emuMaterial(m, hs);
    for (int i = 0; i < n; i++) {
        BSDFSample bs;
        bsdfSample(hs.host.materials[0], mk(wi[0], wi[1], wi[2]), samples[2 * i], samples[2 * i + 1], bs);
        wo[3 * i] = bs.wo.x; wo[3 * i + 1] = bs.wo.y; wo[3 * i + 2] = bs.wo.z;
        weight[3 * i] = bs.weight.x; weight[3 * i + 1] = bs.weight.y; weight[3 * i + 2] = bs.weight.z;
        pdf[i] = bs.pdf; sampledType[i] = (int)bs.sampledType;
    }
    return 0;
}
extern "C" int gdb200_emu_bsdf_eval_batch(const gdb200_material *m, const double *wi, int n, const double *wo, int measure,
                                          double *value, double *pdf)
{
    static HostScene hs; emuMaterial(m, hs);
    for (int i = 0; i < n; i++) {
        Spec f; Float p;
        bsdfEvalPdf(hs.host.materials[0], mk(wi[0], wi[1], wi[2]), mk(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), measure ? EDiscrete : ESolidAngle, f, p);
        value[3 * i] = f.x; value[3 * i + 1] = f.y; value[3 * i + 2] = f.z; pdf[i] = p;
    }
    return 0;
}
static int emuEnvScene(const gdb200_scene_desc *desc, HostScene &hs)
{
    if (int rc = flattenScene(desc, &hs)) return rc;
    if (!hs.host.env.present) return set_error(GDB200_ERR_ARGUMENT, "scene has no environment emitter");
    classifyMaterials(&hs, 0.001);
    // isolate the environment emitter: one-entry selection CDF
    hs.host.emitters[0] = hs.host.emitters[hs.host.env.emitter]; hs.host.emitters[0].pdfDiscrete = 1.0;
    hs.host.nEmitters = 1; hs.host.emCdf[0] = 0; hs.host.emCdf[1] = 1; hs.host.env.emitter = 0;
    hs.host.env.texels = hs.envTexels.data(); hs.host.env.rowWeights = hs.envRowWeights.data();
    hs.host.env.cdfRows = hs.envCdfRows.data(); hs.host.env.cdfCols = hs.envCdfCols.data();
    hs.host.emTris = hs.emTris.data(); hs.host.emTriCdf = hs.emTriCdf.data(); hs.host.triNormals = hs.triNormals.data();
    hs.host.bvh = hs.bvh.data(); hs.host.bvhTris = hs.bvhTris.data();
    c_scene = hs.host; static DScene sceneCopy; sceneCopy = hs.host; c_sceneG = &sceneCopy;
    memcpy(c_bounds, hs.bounds, sizeof(hs.bounds));
    return 0;
}
extern "C" int gdb200_emu_envmap_sample_batch(const gdb200_scene_desc *desc, const double *ref, int n, const double *samples, double *d, double *pdf)
{
    static HostScene hs;
    if (int rc = emuEnvScene(desc, hs)) return rc;
    for (int i = 0; i < n; i++) {
        DRec r; r.ref = mk(ref[0], ref[1], ref[2]); r.refN = mk(0, 0, 0);
        bool vis;
        sampleEmitterDirectVisible(r, samples[2 * i], samples[2 * i + 1], vis);
        d[3 * i] = r.d.x; d[3 * i + 1] = r.d.y; d[3 * i + 2] = r.d.z; pdf[i] = r.pdf;
    }
    return 0;
}
extern "C" int gdb200_emu_envmap_pdf_batch(const gdb200_scene_desc *desc, int n, const double *d, double *pdf)
{
    static HostScene hs;
    if (int rc = emuEnvScene(desc, hs)) return rc;
    for (int i = 0; i < n; i++) pdf[i] = envPdfDirection(xfVector(c_scene.env.toObject, mk(d[3 * i], d[3 * i + 1], d[3 * i + 2])));
    return 0;
}
