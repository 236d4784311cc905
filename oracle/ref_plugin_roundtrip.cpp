// Test infrastructure: the integrator plugin's scene flattening (gradientdomain-mitsuba_b200/plugin/gpt_plugin.cpp, FlatScene::build --
// Mitsuba Scene -> gdb200_scene_desc) run on a REAL Mitsuba Scene.  The scene is built by oracle/_ref/libref_mitsuba.so from a
// gdb200_scene_desc, so the test is a round trip: desc -> Mitsuba objects (the reference's own classes) -> plugin -> desc'.
// Linked against libgdb200.so because the plugin source also holds the render() that calls it; nothing here renders.
#include "../gradientdomain-mitsuba_b200/plugin/gpt_plugin.cpp"

extern "C" void *gdbref_build_scene(const gdb200_scene_desc *, const gdb200_gpt_params *, double, const char *);
extern "C" void gdbref_release_scene(void *);
extern "C" const char *gdbref_gpt_last_error();

namespace { std::string g_rt_error; }

extern "C" const char *gdbref_roundtrip_last_error() { return g_rt_error.c_str(); }

extern "C" const gdb200_scene_desc *gdbref_plugin_flatten(const gdb200_scene_desc *desc, const gdb200_gpt_params *prm, double fov_x_deg, const char *rfilter)
{
    static mitsuba::FlatScene *flat = NULL;
    void *handle = gdbref_build_scene(desc, prm, fov_x_deg, rfilter);
    if (!handle) { g_rt_error = gdbref_gpt_last_error(); return NULL; }
    try {
        delete flat;
        flat = new mitsuba::FlatScene();
        mitsuba::Scene *scene = static_cast<mitsuba::Scene *>(handle);
        flat->build(scene, scene->getSensor());
        gdbref_release_scene(handle);
        return &flat->desc;
    } catch (const std::exception &e) { g_rt_error = e.what(); gdbref_release_scene(handle); return NULL; }
}

// The plugin's render() driven the way Mitsuba's RenderJob drives an integrator (renderjob.cpp:36-70, 88-140): a real Scene
// (sensor, MultiFilm, gdb200_counter sampler: the reference's own classes, built by libref_mitsuba.so), the scene / sensor /
// sampler registered as Scheduler resources, a RenderQueue, then Integrator::render and MultiFilm::develop, which writes
// <dest>-final.pfm, -throughput, -dx, -dy, -direct (multifilm.cpp:423-516).  Needs a GPU: render() calls libgdb200.
extern "C" int gdbref_plugin_render(const gdb200_scene_desc *desc, const gdb200_gpt_params *prm, double fov_x_deg, const char *rfilter,
                                    int reconstructL1, int reconstructL2, double reconstructAlpha, const char *dest)
{
    using namespace mitsuba;
    void *handle = gdbref_build_scene(desc, prm, fov_x_deg, rfilter);
    if (!handle) { g_rt_error = gdbref_gpt_last_error(); return 1; }
    Scene *scene = static_cast<Scene *>(handle);
    int rc = 0, sceneRes = -1, sensorRes = -1, samplerRes = -1;
    Scheduler *sched = Scheduler::getInstance();
    ref<Worker> worker;                                  // a host has at least one local worker (mitsuba.cpp:259-262); the plugin schedules nothing on it
    try {
        if (sched->getCoreCount() == 0) { worker = new LocalWorker(0, "wrk0"); sched->registerWorker(worker.get()); }
        if (!sched->isRunning()) sched->start();
        Properties ip("gpt");                                                      // the XML parameters of <integrator type="gpt">
        ip.setInteger("maxDepth", prm->max_depth); ip.setInteger("rrDepth", prm->rr_depth);
        ip.setBoolean("strictNormals", prm->strict_normals != 0); ip.setFloat("shiftThreshold", prm->shift_threshold);
        ip.setBoolean("reconstructL1", reconstructL1 != 0); ip.setBoolean("reconstructL2", reconstructL2 != 0);
        ip.setFloat("reconstructAlpha", reconstructAlpha);
        ip.setSize("streamsPerPixel", (size_t) std::max(1, prm->streams_per_pixel));
        if (prm->flags & GDB200_GPT_REF_UNINIT_MEASURE) ip.setBoolean("refUninitMeasure", true);
        ref<GDB200GradientPathIntegrator> integ = new GDB200GradientPathIntegrator(ip);
        integ->configure();
        ref<Sensor> sensor = scene->getSensor();
        ref<Sampler> sampler = scene->getSampler();
        sceneRes = sched->registerResource(scene);
        sensorRes = sched->registerResource(sensor);
        std::vector<SerializableObject *> samplers(sched->getCoreCount());             // one sampler per core, renderjob.cpp:60-69
        for (size_t i = 0; i < samplers.size(); ++i) { ref<Sampler> c = sampler->clone(); c->incRef(); samplers[i] = c.get(); }
        samplerRes = sched->registerMultiResource(samplers);
        for (size_t i = 0; i < samplers.size(); ++i) samplers[i]->decRef();
        ref<RenderQueue> queue = new RenderQueue();
        ref<Film> film = sensor->getFilm();
        film->setDestinationFile(fs::path(dest), 0);
        if (!integ->render(scene, queue.get(), NULL, sceneRes, sensorRes, samplerRes)) { g_rt_error = "render() returned false"; rc = 2; }
        else film->develop(scene, 0);
    } catch (const std::exception &e) { g_rt_error = e.what(); rc = 1; }
    if (samplerRes >= 0) sched->unregisterResource(samplerRes);
    if (sensorRes >= 0) sched->unregisterResource(sensorRes);
    if (sceneRes >= 0) sched->unregisterResource(sceneRes);
    if (sched->isRunning()) sched->stop();
    if (worker) sched->unregisterWorker(worker.get());
    gdbref_release_scene(handle);
    return rc;
}
