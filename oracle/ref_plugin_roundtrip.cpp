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
