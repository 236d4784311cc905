// Test infrastructure: the REFERENCE's G-PT integrator -- src/integrators/gpt/gpt.cpp with its evaluatePoint / shift
// mappings / MIS and renderBlock's 15 film splats, plus the scene (kd-tree), shapes, emitters, BSDFs, perspective / thinlens
// sensor, reconstruction filters and ImageBlock::put it runs on -- compiled from /root/reference by oracle/Makefile into
// oracle/_ref/libref_mitsuba.so and driven from a gdb200_scene_desc, the same bytes the CUDA tracer and the CPU restatement
// take.  Random numbers come from the repo's own sampler plugin (plugin/samplers/gdb200_counter.cpp, compiled against the
// real headers here), so the three implementations consume identical per-pixel streams.
//
// What is NOT the reference here: the loop over image blocks (the reference hands blocks to its scheduler's workers,
// gpt_proc.cpp:86-117; this file hands them to std::threads) and the construction of the scene objects from the
// C structs instead of from XML.  Each block is rendered by GradientPathIntegrator::renderBlock (gpt.cpp:1219-1352) and
// merged with ImageBlock::put (imageblock.h:109-113), exactly what GPTRenderProcess::processResult does
// (gpt_proc.cpp:119-128 -> MultiFilm::put).
#include <mitsuba/render/scene.h>
#include <mitsuba/render/trimesh.h>
#include <mitsuba/render/imageblock.h>
#include <mitsuba/render/renderproc.h>
#include <mitsuba/core/statistics.h>
#include <mitsuba/core/sched.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/bitmap.h>
#include <sstream>
#include <thread>
#include <chrono>
#include <mutex>
#include <atomic>
#define private public                   /* GradientPathIntegrator::m_config is filled by render() (gpt.cpp:1365-1370), which is bypassed */
#include "/root/reference/src/integrators/gpt/gpt.h"
#undef private
#include "/root/reference/src/integrators/gpt/gpt_wr.h"
#include "../include/gdb200.h"

using namespace mitsuba;

extern "C" {
void *CreateInstance_diffuse(const Properties &); void *CreateInstance_roughconductor(const Properties &);
void *CreateInstance_conductor(const Properties &); void *CreateInstance_dielectric(const Properties &);
void *CreateInstance_plastic(const Properties &); void *CreateInstance_roughdielectric(const Properties &);
void *CreateInstance_twosided(const Properties &);
void *CreateInstance_rectangle(const Properties &); void *CreateInstance_sphere(const Properties &);
void *CreateInstance_envmap(const Properties &); void *CreateInstance_area(const Properties &); void *CreateInstance_point(const Properties &); void *CreateInstance_spot(const Properties &);
void *CreateInstance_perspective(const Properties &); void *CreateInstance_thinlens(const Properties &);
void *CreateInstance_multifilm(const Properties &);
void *CreateInstance_box(const Properties &); void *CreateInstance_gaussian(const Properties &); void *CreateInstance_tent(const Properties &);
void *CreateInstance_gdb200_counter(const Properties &);
void *CreateInstance_gpt(const Properties &);
}

namespace {
std::string g_error;
std::once_flag g_init;

void staticInit()
{                                                                                // the order of src/mitsuba/mitsuba.cpp:362-373
    Class::staticInitialization();
    Object::staticInitialization();
    Statistics::staticInitialization();
    Thread::staticInitialization();
    Logger::staticInitialization();
    Spectrum::staticInitialization();
    Bitmap::staticInitialization();
    Scheduler::staticInitialization();
    Thread::getThread()->getLogger()->setLogLevel(EError);                      // the default appender writes to stdout, which callers (bench.py) keep for their own output
}

template <typename T> T *make(void *(*factory)(const Properties &), const Properties &props)
{
    T *obj = static_cast<T *>(static_cast<ConfigurableObject *>(factory(props)));
    obj->incRef();
    return obj;
}
void attach(ConfigurableObject *parent, ConfigurableObject *child)               // scenehandler.cpp: addChild, then setParent
{
    parent->addChild(child);
    child->setParent(parent);
}
Spectrum rgb(const double *v) { Spectrum s; s.fromLinearRGB(v[0], v[1], v[2]); return s; }
Transform transformOf(const double *m)
{
    Matrix4x4 mat;
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) mat.m[r][c] = m[4 * r + c];
    return Transform(mat);
}

BSDF *makeBSDF(const gdb200_material &m)
{
    BSDF *bsdf = NULL;
    const char *distr = m.distribution == GDB200_MICROFACET_BECKMANN ? "beckmann" : "ggx";
    switch (m.type) {
        case GDB200_BSDF_DIFFUSE: { Properties p("diffuse"); p.setSpectrum("reflectance", rgb(m.reflectance)); bsdf = make<BSDF>(CreateInstance_diffuse, p); break; }
        case GDB200_BSDF_ROUGHCONDUCTOR: case GDB200_BSDF_CONDUCTOR: {
            Properties p(m.type == GDB200_BSDF_CONDUCTOR ? "conductor" : "roughconductor");
            p.setString("material", "none"); p.setSpectrum("eta", rgb(m.eta)); p.setSpectrum("k", rgb(m.k)); p.setFloat("extEta", 1.0);
            p.setSpectrum("specularReflectance", rgb(m.specular_reflectance));
            if (m.type == GDB200_BSDF_ROUGHCONDUCTOR) { p.setFloat("alpha", m.alpha); p.setString("distribution", distr); }
            bsdf = make<BSDF>(m.type == GDB200_BSDF_CONDUCTOR ? CreateInstance_conductor : CreateInstance_roughconductor, p);
            break;
        }
        case GDB200_BSDF_DIELECTRIC: case GDB200_BSDF_ROUGHDIELECTRIC: {
            Properties p(m.type == GDB200_BSDF_DIELECTRIC ? "dielectric" : "roughdielectric");
            p.setFloat("intIOR", m.ior_ratio); p.setFloat("extIOR", 1.0);
            p.setSpectrum("specularReflectance", rgb(m.specular_reflectance)); p.setSpectrum("specularTransmittance", rgb(m.specular_transmittance));
            if (m.type == GDB200_BSDF_ROUGHDIELECTRIC) { p.setFloat("alpha", m.alpha); p.setString("distribution", distr); }
            bsdf = make<BSDF>(m.type == GDB200_BSDF_DIELECTRIC ? CreateInstance_dielectric : CreateInstance_roughdielectric, p);
            break;
        }
        case GDB200_BSDF_PLASTIC: {
            Properties p("plastic");
            p.setFloat("intIOR", m.ior_ratio); p.setFloat("extIOR", 1.0); p.setBoolean("nonlinear", m.nonlinear != 0);
            p.setSpectrum("diffuseReflectance", rgb(m.reflectance)); p.setSpectrum("specularReflectance", rgb(m.specular_reflectance));
            bsdf = make<BSDF>(CreateInstance_plastic, p);
            break;
        }
        default: throw std::runtime_error("unknown material type");
    }
    bsdf->configure();
    if (m.twosided) {
        BSDF *outer = make<BSDF>(CreateInstance_twosided, Properties("twosided"));
        attach(outer, bsdf);
        outer->configure();
        bsdf = outer;
    }
    return bsdf;
}

struct Built {
    ref<Scene> scene;
    ref<GradientPathIntegrator> gpt;
    ref<Sampler> sampler;
    std::vector<ref<Sampler> > chunkSamplers;   // streams_per_pixel = C > 1: one sampler per chunk (sampleCount and `chunk` set), see renderBlocks
};

Built buildScene(const gdb200_scene_desc *d, const gdb200_gpt_params *prm, double fovX, const char *rfilterName, bool gptIntegrator)
{
    Built out;
    ref<Scene> scene = new Scene(Properties("scene"));

    // ---- sensor <- film <- rfilter, sampler
    const gdb200_camera &cam = d->camera;
    Properties sp(cam.aperture_radius > 0 ? "thinlens" : "perspective");
    sp.setTransform("toWorld", transformOf(cam.camera_to_world));
    sp.setFloat("fov", fovX); sp.setString("fovAxis", "x");
    sp.setFloat("nearClip", cam.near_clip); sp.setFloat("farClip", cam.far_clip);
    if (cam.aperture_radius > 0) { sp.setFloat("apertureRadius", cam.aperture_radius); sp.setFloat("focusDistance", cam.focus_distance); }
    Sensor *sensor = make<Sensor>(cam.aperture_radius > 0 ? CreateInstance_thinlens : CreateInstance_perspective, sp);
    Properties fp("multifilm");
    fp.setInteger("width", cam.width); fp.setInteger("height", cam.height); fp.setBoolean("banner", false); fp.setString("fileFormat", "pfm"); fp.setString("componentFormat", "float32");
    Film *film = make<Film>(CreateInstance_multifilm, fp);
    const std::string rf(rfilterName);
    ReconstructionFilter *filter = make<ReconstructionFilter>(rf == "gaussian" ? CreateInstance_gaussian : rf == "tent" ? CreateInstance_tent : CreateInstance_box,
                                                              Properties(rf));
    filter->configure();
    attach(film, filter);
    film->configure();
    Properties smp("gdb200_counter");
    smp.setSize("sampleCount", (size_t) prm->spp); smp.setSize("seed", (size_t) prm->seed);
    Sampler *sampler = make<Sampler>(CreateInstance_gdb200_counter, smp);
    sampler->configure();
    attach(sensor, film);
    attach(sensor, sampler);
    sensor->configure();
    scene->addChild(sensor);
    out.sampler = sampler;
    // gdb200's chunked sample streams: chunk c of a pixel holds spp/C samples (+1 for c < spp%C) and draws from the stream
    // that the sampler plugin's `chunk` property selects.  The reference renders them as C passes into one film.
    const int C = std::max(1, prm->streams_per_pixel);
    for (int c = 0; c < C && C > 1; c++) {
        const int count = prm->spp / C + (c < prm->spp % C ? 1 : 0);
        if (count == 0) continue;
        Properties cp("gdb200_counter");
        cp.setSize("sampleCount", (size_t) count); cp.setSize("seed", (size_t) prm->seed); cp.setSize("chunk", (size_t) c);
        Sampler *cs = make<Sampler>(CreateInstance_gdb200_counter, cp);
        cs->configure();
        out.chunkSamplers.push_back(cs);
    }

    // ---- integrator
    Properties ip(gptIntegrator ? "gpt" : "path");
    ip.setInteger("maxDepth", prm->max_depth); ip.setInteger("rrDepth", prm->rr_depth); ip.setBoolean("strictNormals", prm->strict_normals != 0);
    if (gptIntegrator) {
        ip.setFloat("shiftThreshold", prm->shift_threshold); ip.setBoolean("reconstructL1", false); ip.setBoolean("reconstructL2", false);
        GradientPathIntegrator *gpt = make<GradientPathIntegrator>(CreateInstance_gpt, ip);
        gpt->configure();
        gpt->m_config.m_maxDepth = prm->max_depth;                               // gpt.cpp:1365-1370
        gpt->m_config.m_minDepth = 1;
        gpt->m_config.m_rrDepth = prm->rr_depth;
        gpt->m_config.m_strictNormals = prm->strict_normals != 0;
        scene->addChild(gpt);
        out.gpt = gpt;
    } else throw std::runtime_error("only the gpt integrator is wired");

    // ---- emitters that are not attached to a shape keep their slot: Scene::m_emitters follows addChild order
    std::vector<BSDF *> bsdfs(d->n_materials);
    for (int i = 0; i < d->n_materials; i++) bsdfs[i] = makeBSDF(d->materials[i]);
    std::vector<ConfigurableObject *> ordered(d->n_emitters, (ConfigurableObject *) NULL);
    for (int i = 0; i < d->n_emitters; i++) {
        const gdb200_emitter &e = d->emitters[i];
        if (e.type == GDB200_EMITTER_AREA) continue;
        if (e.type == GDB200_EMITTER_ENVMAP) {                                     // envmap.cpp:108-181: the map handed over in memory ("bitmap" data property)
            const gdb200_envmap *env = d->envmap;
            if (!env) throw std::runtime_error("envmap emitter without gdb200_scene_desc.envmap");
            ref<Bitmap> bmp = new Bitmap(Bitmap::ERGB, Bitmap::EFloat32, Vector2i(env->width, env->height));
            memcpy(bmp->getFloat32Data(), env->rgb, sizeof(float) * 3 * (size_t) env->width * env->height);
            bmp->incRef();                                                           // the emitter keeps no reference to its input
            Properties p("envmap");
            Properties::Data data; data.ptr = (uint8_t *) bmp.get(); data.size = sizeof(Bitmap);
            p.setData("bitmap", data);
            p.setFloat("scale", env->scale); p.setTransform("toWorld", transformOf(env->to_world)); p.setFloat("samplingWeight", e.sampling_weight);
            p.setBoolean("cache", false);
            Emitter *em = make<Emitter>(CreateInstance_envmap, p);
            em->configure();
            ordered[i] = em;
            continue;
        }
        Properties p(e.type == GDB200_EMITTER_POINT ? "point" : "spot");
        p.setSpectrum("intensity", rgb(e.radiance)); p.setFloat("samplingWeight", e.sampling_weight);
        if (e.type == GDB200_EMITTER_POINT) p.setPoint("position", Point(e.position[0], e.position[1], e.position[2]));
        else {
            Matrix4x4 inv;                                                           // to_local = rows of the inverse rotation; the world matrix is its inverse
            inv.setIdentity();
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) inv.m[r][c] = e.to_local[3 * r + c];
            Matrix4x4 rot;
            if (!inv.invert(rot)) throw std::runtime_error("spot: singular to_local");
            for (int r = 0; r < 3; r++) rot.m[r][3] = e.position[r];
            p.setTransform("toWorld", Transform(rot));
            p.setFloat("cutoffAngle", radToDeg(e.cutoff_angle)); p.setFloat("beamWidth", radToDeg(e.beam_width));
        }
        Emitter *em = make<Emitter>(e.type == GDB200_EMITTER_POINT ? CreateInstance_point : CreateInstance_spot, p);
        em->configure();
        ordered[i] = em;
    }

    // ---- shapes (with their BSDF and area emitter)
    std::vector<Shape *> shapes(d->n_shapes);
    for (int i = 0; i < d->n_shapes; i++) {
        const gdb200_shape &s = d->shapes[i];
        Shape *shape = NULL;
        if (s.type == GDB200_SHAPE_RECTANGLE) {
            Properties p("rectangle"); p.setTransform("toWorld", transformOf(s.to_world));
            shape = make<Shape>(CreateInstance_rectangle, p);
        } else if (s.type == GDB200_SHAPE_SPHERE) {
            Properties p("sphere"); p.setPoint("center", Point(s.center[0], s.center[1], s.center[2])); p.setFloat("radius", s.radius);
            p.setBoolean("flipNormals", s.flip_normals != 0);
            shape = make<Shape>(CreateInstance_sphere, p);
        } else {
            std::map<int, uint32_t> remap;                                         // the mesh's own vertex table, in first-use order
            std::vector<int> verts;
            for (int t = 0; t < s.tri_count; t++) for (int k = 0; k < 3; k++) {
                const int v = d->triangles[3 * (s.first_tri + t) + k];
                if (!remap.count(v)) { remap[v] = (uint32_t) verts.size(); verts.push_back(v); }
            }
            const bool smooth = s.has_vertex_normals != 0;
            TriMesh *mesh = new TriMesh("mesh", (size_t) s.tri_count, verts.size(), smooth, false, false, false, !smooth);
            mesh->incRef();
            for (size_t v = 0; v < verts.size(); v++) {
                mesh->getVertexPositions()[v] = Point(d->vertices[3 * verts[v]], d->vertices[3 * verts[v] + 1], d->vertices[3 * verts[v] + 2]);
                if (smooth) mesh->getVertexNormals()[v] = Normal(d->normals[3 * verts[v]], d->normals[3 * verts[v] + 1], d->normals[3 * verts[v] + 2]);
            }
            for (int t = 0; t < s.tri_count; t++) for (int k = 0; k < 3; k++)
                mesh->getTriangles()[t].idx[k] = remap[d->triangles[3 * (s.first_tri + t) + k]];
            shape = mesh;
        }
        attach(shape, bsdfs[s.material]);
        if (s.emitter >= 0) {
            const gdb200_emitter &e = d->emitters[s.emitter];
            Properties p("area"); p.setSpectrum("radiance", rgb(e.radiance)); p.setFloat("samplingWeight", e.sampling_weight);
            Emitter *em = make<Emitter>(CreateInstance_area, p);
            em->configure();
            attach(shape, em);
        }
        shape->configure();
        shapes[i] = shape;
    }
    // Scene::m_emitters (the emitter CDF): scene-level emitters in addChild order (scene.cpp:496-516), then the shapes' emitters
    // in shape order when Scene::initialize() runs addShape (scene.cpp:570-571).  The desc must list them the same way.
    for (int i = 0; i < d->n_emitters; i++) if (ordered[i]) scene->addChild(ordered[i]);
    for (int i = 0; i < d->n_shapes; i++) scene->addChild(shapes[i]);

    scene->configure();
    scene->initialize();
    const ref_vector<Emitter> &ems = scene->getEmitters();
    if ((int) ems.size() != d->n_emitters) throw std::runtime_error("emitter count differs from the scene description");
    for (int i = 0; i < d->n_emitters; i++) {
        const Emitter *expect = ordered[i] ? static_cast<const Emitter *>(ordered[i]) : shapes[d->emitters[i].shape]->getEmitter();
        if (ems[i].get() != expect) throw std::runtime_error("the scene description lists its emitters in an order Mitsuba cannot produce (scene-level emitters first, then the shapes' emitters in shape order)");
    }
    out.scene = scene;
    return out;
}

// Block loop: 32x32 blocks (scene.cpp: blockSize default), any order (every pixel re-keys the sampler), `threads` workers.
void renderBlocks(Built &b, int threads, double *out5, double *blockSeconds = nullptr)
{
    Scene *scene = b.scene.get();
    Sensor *sensor = scene->getSensor();
    Film *film = sensor->getFilm();
    const Vector2i size = film->getCropSize();
    const ReconstructionFilter *rfilter = film->getReconstructionFilter();
    const int bs = 32;
    const int nbx = (size.x + bs - 1) / bs, nby = (size.y + bs - 1) / bs;
    ref<GPTWorkResult> total = new GPTWorkResult(rfilter, size, 1);
    total->clear();
    std::mutex merge;
    std::atomic<int> next(0);
    std::string failure;
    auto worker = [&]() {
        try {
            std::vector<ref<Sampler> > samplers;               // one pass per sample stream of a pixel (a single one for the reference's own mode)
            if (b.chunkSamplers.empty()) samplers.push_back(b.sampler->clone());
            else for (size_t c = 0; c < b.chunkSamplers.size(); c++) samplers.push_back(b.chunkSamplers[c]->clone());
            ref<GPTWorkResult> block = new GPTWorkResult(rfilter, Vector2i(bs, bs), 1);
            const bool stop = false;
            for (;;) {
                const int id = next++;
                if (id >= nbx * nby) break;
                const Point2i off((id % nbx) * bs, (id / nbx) * bs);
                const Vector2i sz(std::min(bs, size.x - off.x), std::min(bs, size.y - off.y));
                block->setOffset(off); block->setSize(sz);
                std::vector<TPoint2<uint8_t> > points;
                for (int y = 0; y < sz.y; y++) for (int x = 0; x < sz.x; x++) points.push_back(TPoint2<uint8_t>((uint8_t) x, (uint8_t) y));
                for (size_t c = 0; c < samplers.size(); c++) {
                    b.gpt->renderBlock(scene, sensor, samplers[c].get(), block.get(), stop, points);   // clears the block, then accumulates
                    std::lock_guard<std::mutex> guard(merge);
                    total->put(block.get());
                }
            }
        } catch (const std::exception &e) { std::lock_guard<std::mutex> guard(merge); failure = e.what(); }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (threads <= 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.push_back(std::thread([&]() {
            ref<Thread> self = Thread::registerUnmanagedThread("gdbref");            // Mitsuba services (logger, TLS) for a foreign thread
            worker();
        }));
        for (auto &t : pool) t.join();
    }
    if (blockSeconds) *blockSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!failure.empty()) throw std::runtime_error(failure);
    // develop: value / weight per pixel (Bitmap::convert of ESpectrumAlphaWeight, what MultiFilm::develop writes)
    for (int buf = 0; buf < 5; buf++) {
        const ImageBlock *ib = total->getImageBlock(buf);
        const Bitmap *bmp = ib->getBitmap();
        const int border = ib->getBorderSize(), stride = bmp->getWidth();
        const Float *data = bmp->getFloatData();
        for (int y = 0; y < size.y; y++) for (int x = 0; x < size.x; x++) {
            const Float *px = data + ((size_t) (y + border) * stride + (x + border)) * (SPECTRUM_SAMPLES + 2);
            const Float w = px[SPECTRUM_SAMPLES + 1], inv = w != 0 ? 1 / w : 0;
            double *o = out5 + (((size_t) buf * size.y + y) * size.x + x) * 3;
            for (int c = 0; c < 3; c++) o[c] = px[c] * inv;
        }
    }
}
}

extern "C" {

const char *gdbref_gpt_last_error() { return g_error.c_str(); }

// Mitsuba's start-up sequence, once per process (every entry point of the library calls it first).
void gdbref_static_init() { std::call_once(g_init, staticInit); }

// The Mitsuba Scene built from `desc` (reference counted; gdbref_release_scene drops it) -- for
// oracle/ref_plugin_roundtrip.cpp, which runs the integrator plugin's scene flattening on it.
void *gdbref_build_scene(const gdb200_scene_desc *desc, const gdb200_gpt_params *prm, double fov_x_deg, const char *rfilter)
{
    try {
        std::call_once(g_init, staticInit);
        Built b = buildScene(desc, prm, fov_x_deg, rfilter, true);
        b.scene->incRef();
        return b.scene.get();
    } catch (const std::exception &e) { g_error = e.what(); return NULL; }
}
void gdbref_release_scene(void *scene) { if (scene) static_cast<Scene *>(scene)->decRef(); }

// GradientPathIntegrator::Li (gpt.cpp:1489-1662, the plain MIS path tracer G-PT carries for sub-surface preprocessing) averaged
// over the pixel's samples: out = [h][w][3].  The camera sample loop around it mirrors SamplingIntegrator::renderBlock; the
// sampler is keyed with seed ^ 0x5bd1e995 like the restatement's gdb200_oracle_path_render.
int gdbref_gpt_li_render(const gdb200_scene_desc *desc, const gdb200_gpt_params *prm, double fov_x_deg, const char *rfilter, double *out)
{
    try {
        std::call_once(g_init, staticInit);
        gdb200_gpt_params q = *prm;
        q.seed = prm->seed ^ 0x5bd1e995u;
        Built b = buildScene(desc, &q, fov_x_deg, rfilter, true);
        Scene *scene = b.scene.get();
        Sensor *sensor = scene->getSensor();
        Sampler *sampler = b.sampler.get();
        const Vector2i size = sensor->getFilm()->getCropSize();
        const bool needsAperture = sensor->needsApertureSample();
        RadianceQueryRecord rRec(scene, sampler);
        for (int y = 0; y < size.y; y++) for (int x = 0; x < size.x; x++) {
            sampler->generate(Point2i(x, y));
            Spectrum sum(0.0f);
            for (int j = 0; j < prm->spp; j++) {
                rRec.newQuery(RadianceQueryRecord::ERadiance, sensor->getMedium());
                const Point2 samplePos(Point2(Point2i(x, y)) + Vector2(rRec.nextSample2D()));
                Point2 apertureSample(0.5f);
                if (needsAperture) apertureSample = rRec.nextSample2D();
                RayDifferential ray;
                Spectrum spec = sensor->sampleRayDifferential(ray, samplePos, apertureSample, 0.5f);
                spec *= b.gpt->Li(ray, rRec);
                sum += spec;
            }
            const Spectrum mean = sum / (Float) prm->spp;
            Float r, g, bl; mean.toLinearRGB(r, g, bl);
            double *o = out + ((size_t) y * size.x + x) * 3;
            o[0] = r; o[1] = g; o[2] = bl;
        }
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return 1; }
}

// PerspectiveCamera's horizontal field of view for a sensor described the way a scene file does (sensor.cpp:244-307):
// `fov` with `fovAxis` (x, y, diagonal, smaller, larger) when fov >= 0, else `focalLength` (e.g. "50mm"); film width x height.
double gdbref_sensor_xfov(double fov, const char *fovAxis, const char *focalLength, int width, int height)
{
    try {
        std::call_once(g_init, staticInit);
        Properties sp("perspective");
        if (fov >= 0) { sp.setFloat("fov", fov); sp.setString("fovAxis", fovAxis); }
        else if (focalLength && focalLength[0]) sp.setString("focalLength", focalLength);
        Sensor *sensor = make<Sensor>(CreateInstance_perspective, sp);
        Properties fp("multifilm");
        fp.setInteger("width", width); fp.setInteger("height", height); fp.setBoolean("banner", false); fp.setString("fileFormat", "pfm"); fp.setString("componentFormat", "float32");
        Film *film = make<Film>(CreateInstance_multifilm, fp);
        ReconstructionFilter *filter = make<ReconstructionFilter>(CreateInstance_box, Properties("box"));
        filter->configure();
        attach(film, filter);
        film->configure();
        Sampler *sampler = make<Sampler>(CreateInstance_gdb200_counter, Properties("gdb200_counter"));
        sampler->configure();
        attach(sensor, film);
        attach(sensor, sampler);
        sensor->configure();
        return static_cast<PerspectiveCamera *>(sensor)->getXFov();
    } catch (const std::exception &e) { g_error = e.what(); return -1; }
}

// Scene::rayIntersect (the kd-tree) for n rays with mint = Epsilon, maxt = inf: hit distance (inf = miss), index of the hit
// shape in the description, geometric normal, shading frame (s, t, n).
int gdbref_intersect_batch(void *handle, int n, const double *o, const double *d, double *t, int *shape, double *geoN, double *shFrame)
{
    try {
        Scene *scene = static_cast<Scene *>(handle);
        const ref_vector<Shape> &shapes = scene->getShapes();
        for (int i = 0; i < n; i++) {
            Ray ray(Point(o[3 * i], o[3 * i + 1], o[3 * i + 2]), Vector(d[3 * i], d[3 * i + 1], d[3 * i + 2]), Epsilon, std::numeric_limits<Float>::infinity(), 0.0f);
            Intersection its;
            const bool hit = scene->rayIntersect(ray, its);
            t[i] = hit ? its.t : std::numeric_limits<Float>::infinity();
            shape[i] = -1;
            if (hit) for (size_t k = 0; k < shapes.size(); k++) if (shapes[k].get() == its.shape) shape[i] = (int) k;
            const Vector f[4] = {hit ? Vector(its.geoFrame.n) : Vector(0.0f), hit ? its.shFrame.s : Vector(0.0f), hit ? its.shFrame.t : Vector(0.0f), hit ? Vector(its.shFrame.n) : Vector(0.0f)};
            for (int c = 0; c < 3; c++) { geoN[3 * i + c] = f[0][c]; for (int k = 0; k < 3; k++) shFrame[9 * i + 3 * k + c] = f[1 + k][c]; }
        }
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return 1; }
}

// Scene::getAABB() as the environment emitter's createShape saw it (scene.cpp:386-396): out = min[3], max[3] of the kd-tree's
// box, then min[3], max[3] of the sensor's box.
int gdbref_scene_bounds(void *handle, double *out)
{
    try {
        Scene *scene = static_cast<Scene *>(handle);
        const AABB a = scene->getKDTree()->getAABB(), b = scene->getSensor()->getAABB();
        for (int k = 0; k < 3; k++) { out[k] = a.min[k]; out[3 + k] = a.max[k]; out[6 + k] = b.min[k]; out[9 + k] = b.max[k]; }
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return 1; }
}

// EnvironmentMap::sampleDirect / pdfDirect / evalEnvironment of the scene's environment emitter (envmap.cpp:516-556,376-409)
// from the reference point `ref`: per sample the world direction, the solid-angle density sampleDirect reports, the density
// pdfDirect reports for that direction, value / pdf as returned (RGB) and evalEnvironment along the direction (RGB).
int gdbref_envmap_sample(void *handle, const double *ref, int n, const double *samples, double *dir, double *pdfSample, double *pdfEval,
                         double *weight, double *radiance)
{
    try {
        Scene *scene = static_cast<Scene *>(handle);
        const Emitter *env = scene->getEnvironmentEmitter();
        if (!env) throw std::runtime_error("the scene has no environment emitter");
        for (int i = 0; i < n; i++) {
            DirectSamplingRecord dRec(Point(ref[0], ref[1], ref[2]), 0.0f);
            const Spectrum w = env->sampleDirect(dRec, Point2(samples[2 * i], samples[2 * i + 1]));
            dir[3 * i] = dRec.d.x; dir[3 * i + 1] = dRec.d.y; dir[3 * i + 2] = dRec.d.z;
            pdfSample[i] = dRec.pdf;
            pdfEval[i] = env->pdfDirect(dRec);
            Float r, g, b; w.toLinearRGB(r, g, b);
            weight[3 * i] = r; weight[3 * i + 1] = g; weight[3 * i + 2] = b;
            const Spectrum L = env->evalEnvironment(RayDifferential(Ray(dRec.ref, dRec.d, 0.0f)));
            L.toLinearRGB(r, g, b);
            radiance[3 * i] = r; radiance[3 * i + 1] = g; radiance[3 * i + 2] = b;
        }
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return 1; }
}

// The reference G-PT tracer on `desc`: out5 = [5][h][w][3] developed buffers in the order -final (preview), -throughput,
// -dx, -dy, -direct.  fov_x_deg and rfilter are what the desc's matrices / filter table were made from.
int gdbref_gpt_render(const gdb200_scene_desc *desc, const gdb200_gpt_params *prm, double fov_x_deg, const char *rfilter, int threads, double *out5)
{
    try {
        std::call_once(g_init, staticInit);
        Built b = buildScene(desc, prm, fov_x_deg, rfilter, true);
        renderBlocks(b, threads, out5);
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return 1; }
}

// The same, reporting the wall time of the block loop alone -- worker start to last join: what Mitsuba's "Render time" covers
// (renderjob.cpp:108), without building the Scene / kd-tree before it or developing the film after it.  bench.py's CPU baseline.
int gdbref_gpt_render_timed(const gdb200_scene_desc *desc, const gdb200_gpt_params *prm, double fov_x_deg, const char *rfilter, int threads,
                            double *out5, double *render_seconds)
{
    try {
        std::call_once(g_init, staticInit);
        Built b = buildScene(desc, prm, fov_x_deg, rfilter, true);
        renderBlocks(b, threads, out5, render_seconds);
        return 0;
    } catch (const std::exception &e) { g_error = e.what(); return 1; }
}

}
