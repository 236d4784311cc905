// TEST INFRASTRUCTURE -- the REFERENCE's own G-BDPT integrator (src/integrators/gbdpt/{gbdpt,gbdpt_proc,gbdpt_wr}.cpp over
// src/libbidir, compiled unmodified into oracle/_ref/libref_gbdpt.so by oracle/Makefile) driven from a gdb200_scene_desc.
// The Scene is the one ref_gpt_shim.cpp builds (the reference's own Scene / Shape / BSDF / Emitter / Sensor / MultiFilm
// classes, gdb200_counter sampler); the integrator is swapped for `gbdpt` and rendered the way Mitsuba renders anything:
// a RenderJob on the Scheduler with local workers (renderjob.cpp:88-140 -> gbdpt.cpp:140-262 -> GBDPTProcess), then
// MultiFilm::develop, which writes <dest>-L1 / -L2 / -gradientNegY / -gradientNegX / -gradientPosX / -gradientPosY /
// -primal .pfm (gbdpt.cpp:164, multifilm.cpp:423-516).  This is the pin for SURVEY.md §8f-1 (BASELINE config 4): golden
// outputs of the reference for the G-BDPT path, before any restatement or kernel exists.  Nothing in the product loads it.
#include <mitsuba/render/scene.h>
#include <mitsuba/render/renderjob.h>
#include <mitsuba/render/renderqueue.h>
#include <mitsuba/core/sched.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/logger.h>
#include <cstdlib>
#include <mutex>
#include <string>
#include "../include/gdb200.h"

using namespace mitsuba;

extern "C" void *CreateInstance_gbdpt(const Properties &props);
extern "C" void *gdbref_build_scene(const gdb200_scene_desc *, const gdb200_gpt_params *, double, const char *);
extern "C" void gdbref_release_scene(void *);
extern "C" const char *gdbref_gpt_last_error();

namespace {
std::string g_error;
}

extern "C" const char *gdbref_gbdpt_last_error() { return g_error.c_str(); }

// prm: maxDepth (-1 is clamped to 12 by the reference, gbdpt_proc.cpp:104-107), rrDepth, shiftThreshold, spp, seed.
// light_image: the `lightImage` parameter (paths that hit the sensor by chance are splatted into separate light images).
extern "C" int gdbref_gbdpt_render(const gdb200_scene_desc *desc, const gdb200_gpt_params *prm, double fov_x_deg, const char *rfilter,
                                   int light_image, double reconstruct_alpha, int threads, const char *dest)
{
    void *handle = gdbref_build_scene(desc, prm, fov_x_deg, rfilter);
    if (!handle) { g_error = gdbref_gpt_last_error(); return 1; }
    Scene *scene = static_cast<Scene *>(handle);
    int rc = 0;
    try {
        if (getenv("GDBREF_LOG")) Thread::getThread()->getLogger()->setLogLevel(EInfo);   // the library keeps Mitsuba's log at EError otherwise
        Scheduler *sched = Scheduler::getInstance();
        std::vector<ref<Worker> > workers;                                          // started and stopped per call: no thread outlives it
        for (int i = 0; i < std::max(1, threads); i++) { workers.push_back(new LocalWorker(i, formatString("wrk%i", i))); sched->registerWorker(workers.back()); }
        sched->start();
        struct StopScheduler { Scheduler *s; std::vector<ref<Worker> > &w; ~StopScheduler() { s->stop(); for (size_t i = 0; i < w.size(); i++) s->unregisterWorker(w[i]); } } stop = {sched, workers};
        Properties ip("gbdpt");
        ip.setInteger("maxDepth", prm->max_depth); ip.setInteger("rrDepth", prm->rr_depth);
        ip.setFloat("shiftThreshold", prm->shift_threshold); ip.setBoolean("lightImage", light_image != 0);
        ip.setBoolean("reconstructL1", true); ip.setBoolean("reconstructL2", false); ip.setFloat("reconstructAlpha", reconstruct_alpha);
        ref<Integrator> integ = static_cast<Integrator *>(static_cast<ConfigurableObject *>(CreateInstance_gbdpt(ip)));
        integ->configure();
        scene->setIntegrator(integ.get());
        scene->setDestinationFile(fs::path(dest));
        ref<RenderQueue> queue = new RenderQueue();
        ref<RenderJob> job = new RenderJob("gbdpt", scene, queue.get(), -1, -1, -1, true, false);
        queue->addJob(job);                                                          // mitsuba.cpp: addJob, start, waitLeft, join
        job->start();
        queue->waitLeft(0);
        queue->join();
    } catch (const std::exception &e) { g_error = e.what(); rc = 1; }
    gdbref_release_scene(handle);
    return rc;
}
