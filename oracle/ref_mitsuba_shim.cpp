// Test infrastructure: C entry points over the REFERENCE's own BSDF plugins, microfacet distribution, Fresnel and warp
// routines.  `make -C oracle ref_mitsuba` compiles those sources unmodified where they lie under /root/reference
// (src/bsdfs/{diffuse,roughconductor,conductor,dielectric,plastic,roughdielectric,twosided}.cpp, src/bsdfs/microfacet.h,
// src/libcore/{util,warp,math,quad,spectrum,object,class,serialization}.cpp, src/librender/{bsdf,texture,shader,sampler}.cpp,
// src/libhw/basicshader.cpp) against oracle/refstubs (a few boost stand-ins + inert runtime services) into
// oracle/_ref/libref_mitsuba.so.  tests/test_ref_mitsuba.py pins the restatement in oracle/gpt_oracle.cpp -- and the
// device code through tests/emu -- against these entry points.  Never linked into the product.
#include <mitsuba/render/bsdf.h>
#include <mitsuba/render/sampler.h>
#include <mitsuba/render/scene.h>
#include <mitsuba/core/properties.h>
#include <mitsuba/core/warp.h>
#include <mitsuba/core/frame.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/fstream.h>
#include <mitsuba/render/trimesh.h>
#include "/root/reference/src/bsdfs/microfacet.h"
#include <stdexcept>
#include <string>
#include <cstring>

using namespace mitsuba;

extern "C" {
void *CreateInstance_diffuse(const Properties &);
void *CreateInstance_roughconductor(const Properties &);
void *CreateInstance_conductor(const Properties &);
void *CreateInstance_dielectric(const Properties &);
void *CreateInstance_plastic(const Properties &);
void *CreateInstance_roughdielectric(const Properties &);
void *CreateInstance_twosided(const Properties &);
void *CreateInstance_obj(const Properties &);
}

namespace {
// The FakeSampler of src/tests/test_chisquare.cpp: hands out a fixed value (the in-BSDF draw of roughdielectric).
class FixedSampler : public Sampler {
public:
    FixedSampler() : Sampler(Properties()), value(0) {}
    Float next1D() { return value; }
    Point2 next2D() { return Point2(value, value); }
    Float value;
};
std::string g_error;
template <typename F> int guarded(F f)
{
    try { f(); return 0; } catch (const std::exception &e) { g_error = e.what(); return 1; }
}
}

extern "C" {

void gdbref_static_init();                                                     // ref_gpt_shim.cpp
const char *gdbref_last_error() { return g_error.c_str(); }

// kinds: 0 float, 1 spectrum (RGB), 2 string, 3 boolean.  `nested` (a BSDF from this function) is attached as a child
// before configure() -- twosided.
void *gdbref_bsdf_create(const char *plugin, int n, const char **keys, const int *kinds, const double *values, const char **strings, void *nested)
{
    void *result = NULL;
    guarded([&] {
        gdbref_static_init();                                                  // incl. Class::staticInitialization: the super-class links derivesFrom() walks
        Properties props(plugin);
        for (int i = 0; i < n; i++) {
            if (kinds[i] == 0) props.setFloat(keys[i], values[3 * i]);
            else if (kinds[i] == 1) { Spectrum s; s.fromLinearRGB(values[3 * i], values[3 * i + 1], values[3 * i + 2]); props.setSpectrum(keys[i], s); }
            else if (kinds[i] == 2) props.setString(keys[i], strings[i]);
            else props.setBoolean(keys[i], values[3 * i] != 0);
        }
        const std::string p(plugin);
        void *obj = p == "diffuse" ? CreateInstance_diffuse(props) : p == "roughconductor" ? CreateInstance_roughconductor(props)
                  : p == "conductor" ? CreateInstance_conductor(props) : p == "dielectric" ? CreateInstance_dielectric(props)
                  : p == "plastic" ? CreateInstance_plastic(props) : p == "roughdielectric" ? CreateInstance_roughdielectric(props)
                  : p == "twosided" ? CreateInstance_twosided(props) : NULL;
        if (!obj) throw std::runtime_error("unknown BSDF plugin " + p);
        BSDF *bsdf = static_cast<BSDF *>(static_cast<ConfigurableObject *>(obj));
        bsdf->incRef();
        if (nested) bsdf->addChild(static_cast<BSDF *>(nested));
        bsdf->configure();
        result = bsdf;
    });
    return result;
}

int gdbref_bsdf_info(void *handle, unsigned *type, double *eta)
{
    return guarded([&] { BSDF *b = static_cast<BSDF *>(handle); *type = b->getType(); *eta = b->getEta(); });
}

// BSDF::eval / BSDF::pdf for one wi and n outgoing directions (local frame); measure 0 = solid angle, 1 = discrete.
int gdbref_bsdf_eval(void *handle, const double *wi, int n, const double *wo, int measure, double *value, double *pdf)
{
    return guarded([&] {
        BSDF *bsdf = static_cast<BSDF *>(handle);
        Intersection its;
        for (int i = 0; i < n; i++) {
            BSDFSamplingRecord bRec(its, Vector(wi[0], wi[1], wi[2]), Vector(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), ERadiance);
            const EMeasure m = measure ? EDiscrete : ESolidAngle;
            const Spectrum f = bsdf->eval(bRec, m);
            Float r, g, b; f.toLinearRGB(r, g, b);
            value[3 * i] = r; value[3 * i + 1] = g; value[3 * i + 2] = b;
            pdf[i] = bsdf->pdf(bRec, m);
        }
    });
}

// BSDF::sample(bRec, pdf, sample) as gpt.cpp:456-457 calls it; samples = (sx, sy, in-BSDF sampler draw) per entry.
int gdbref_bsdf_sample(void *handle, const double *wi, int n, const double *samples, double *wo, double *weight, double *pdf,
                       double *eta, int *sampledType)
{
    return guarded([&] {
        BSDF *bsdf = static_cast<BSDF *>(handle);
        Intersection its;
        ref<FixedSampler> sampler = new FixedSampler();
        for (int i = 0; i < n; i++) {
            sampler->value = samples[3 * i + 2];
            BSDFSamplingRecord bRec(its, sampler.get(), ERadiance);
            bRec.wi = Vector(wi[0], wi[1], wi[2]);
            Float p = 0;
            const Spectrum w = bsdf->sample(bRec, p, Point2(samples[3 * i], samples[3 * i + 1]));
            Float r, g, b; w.toLinearRGB(r, g, b);
            weight[3 * i] = r; weight[3 * i + 1] = g; weight[3 * i + 2] = b;
            wo[3 * i] = bRec.wo.x; wo[3 * i + 1] = bRec.wo.y; wo[3 * i + 2] = bRec.wo.z;
            pdf[i] = p; eta[i] = bRec.eta; sampledType[i] = (int) bRec.sampledType;
        }
    });
}

// MicrofacetDistribution (src/bsdfs/microfacet.h): op 0 eval(m), 1 pdf(wi, m), 2 smithG1(wi, m), 3 G(wi, wo, m),
// 4 sample(wi, sample) -> out = (m.x, m.y, m.z, pdf).  in: 9 doubles per entry (wi, wo-or-sample, m).
int gdbref_microfacet(int type, double alpha, int sampleVisible, int op, int n, const double *in, double *out)
{
    return guarded([&] {
        MicrofacetDistribution d((MicrofacetDistribution::EType) type, alpha, sampleVisible != 0);
        for (int i = 0; i < n; i++) {
            const double *q = in + 9 * i;
            const Vector wi(q[0], q[1], q[2]), wo(q[3], q[4], q[5]); const Normal m(q[6], q[7], q[8]);
            if (op == 0) out[i] = d.eval(m);
            else if (op == 1) out[i] = d.pdf(wi, m);
            else if (op == 2) out[i] = d.smithG1(wi, m);
            else if (op == 3) out[i] = d.G(wi, wo, m);
            else { Float pdf; const Normal s = d.sample(wi, Point2(q[3], q[4]), pdf); out[4 * i] = s.x; out[4 * i + 1] = s.y; out[4 * i + 2] = s.z; out[4 * i + 3] = pdf; }
        }
    });
}

// src/libcore/util.cpp: op 0 fresnelDielectricExt(cosThetaI, eta) -> (F, cosThetaT); 1 fresnelConductorExact(cosThetaI, eta, k)
// -> F (one channel); 2 fresnelDiffuseReflectance(eta, fast = false) -> F; 3 the same with fast = true.  in: 3 doubles per entry.
int gdbref_fresnel(int op, int n, const double *in, double *out)
{
    return guarded([&] {
        for (int i = 0; i < n; i++) {
            const double *q = in + 3 * i;
            if (op == 0) { Float cosThetaT; out[2 * i] = fresnelDielectricExt(q[0], cosThetaT, q[1]); out[2 * i + 1] = cosThetaT; }
            else if (op == 1) out[i] = fresnelConductorExact(q[0], q[1], q[2]);
            else out[i] = fresnelDiffuseReflectance(q[0], op == 3);
        }
    });
}

// src/libcore/warp.cpp: kind 0 squareToCosineHemisphere, 1 squareToUniformDiskConcentric, 2 squareToUniformTriangle,
// 3 squareToUniformCone(cosCutoff = param), 4 squareToUniformSphere.  in: 2 doubles, out: 3 doubles per entry.
int gdbref_warp(int kind, double param, int n, const double *in, double *out)
{
    return guarded([&] {
        for (int i = 0; i < n; i++) {
            const Point2 s(in[2 * i], in[2 * i + 1]);
            Vector v(0.0f);
            if (kind == 0) v = warp::squareToCosineHemisphere(s);
            else if (kind == 1) { const Point2 p = warp::squareToUniformDiskConcentric(s); v = Vector(p.x, p.y, 0); }
            else if (kind == 2) { const Point2 p = warp::squareToUniformTriangle(s); v = Vector(p.x, p.y, 0); }
            else if (kind == 3) v = warp::squareToUniformCone(param, s);
            else v = warp::squareToUniformSphere(s);
            out[3 * i] = v.x; out[3 * i + 1] = v.y; out[3 * i + 2] = v.z;
        }
    });
}


// src/libcore/transform.cpp: kind 0 lookAt(origin, target, up), 1 rotate(axis, angle in degrees = a[3]), 2 scale(v), 3 translate(v),
// 4 perspective(fov = a[0], near = a[1], far = a[2]); out = the 4x4 matrix followed by the 4x4 inverse the Transform carries.
int gdbref_transform(int kind, const double *a, double *out)
{
    return guarded([&] {
        Transform t;
        if (kind == 0) t = Transform::lookAt(Point(a[0], a[1], a[2]), Point(a[3], a[4], a[5]), Vector(a[6], a[7], a[8]));
        else if (kind == 1) t = Transform::rotate(Vector(a[0], a[1], a[2]), a[3]);
        else if (kind == 2) t = Transform::scale(Vector(a[0], a[1], a[2]));
        else if (kind == 3) t = Transform::translate(Vector(a[0], a[1], a[2]));
        else t = Transform::perspective(a[0], a[1], a[2]);
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) { out[4 * r + c] = t.getMatrix()(r, c); out[16 + 4 * r + c] = t.getInverseMatrix()(r, c); }
    });
}


// Bitmap::write(EPFM) / the EPFM reader (src/libcore/bitmap.cpp:3745-3850): writes rgb [h][w][3] float32 to `path` with the
// reference's writer, or reads `path` with the reference's reader into rgb (which must hold w*h*3 floats; w, h are returned).
int gdbref_pfm(int write, const char *path, int *w, int *h, float *rgb)
{
    return guarded([&] {
        gdbref_static_init();
        if (write) {
            ref<Bitmap> bmp = new Bitmap(Bitmap::ERGB, Bitmap::EFloat32, Vector2i(*w, *h));
            memcpy(bmp->getFloat32Data(), rgb, sizeof(float) * 3 * (size_t) *w * *h);
            ref<FileStream> fs = new FileStream(path, FileStream::ETruncWrite);
            bmp->write(Bitmap::EPFM, fs);
        } else {
            ref<FileStream> fs = new FileStream(path, FileStream::EReadOnly);
            ref<Bitmap> bmp = new Bitmap(Bitmap::EPFM, fs);
            if (bmp->getChannelCount() != 3 || bmp->getComponentFormat() != Bitmap::EFloat32) throw std::runtime_error("unexpected PFM layout");
            if (rgb && *w == bmp->getWidth() && *h == bmp->getHeight()) memcpy(rgb, bmp->getFloat32Data(), sizeof(float) * 3 * (size_t) *w * *h);
            *w = bmp->getWidth(); *h = bmp->getHeight();
        }
    });
}


// Triangle meshes through the reference's loaders.  kind 0: the `serialized` container (TriMesh(Stream *, index),
// trimesh.cpp:80-86,175-252), 2: the `obj` plugin (first mesh of the file); then TriMesh::configure()
// (computeNormals, trimesh.cpp:608-681).  Returns the counts; gdbref_mesh_copy hands out the arrays of the last mesh.
static ref<TriMesh> g_mesh;
int gdbref_mesh_load(int kind, const char *path, int index, int faceNormals, int flipNormals, int *counts /* vertices, triangles, has normals */)
{
    return guarded([&] {
        gdbref_static_init();
        if (kind == 0) {
            ref<FileStream> fs = new FileStream(path, FileStream::EReadOnly);
            fs->setByteOrder(Stream::ELittleEndian);
            g_mesh = new TriMesh(fs, index);
        } else {
            if (kind == 1) throw std::runtime_error("the ply plugin needs boost::mpl (ply_parser.hpp) and is not in this build");
            Properties p("obj");
            p.setString("filename", path); p.setBoolean("faceNormals", faceNormals != 0); p.setBoolean("flipNormals", flipNormals != 0);
            ConfigurableObject *obj = static_cast<ConfigurableObject *>(CreateInstance_obj(p));
            Shape *shape = static_cast<Shape *>(obj);
            if (kind == 2) {                                              // the obj plugin is a compound shape: its meshes are elements
                ref<Shape> keep = shape;
                shape->configure();
                g_mesh = static_cast<TriMesh *>(shape->getElement(0));
                if (g_mesh == NULL) throw std::runtime_error("obj: no mesh");
            } else g_mesh = static_cast<TriMesh *>(shape);
        }
        g_mesh->configure();
        counts[0] = (int) g_mesh->getVertexCount(); counts[1] = (int) g_mesh->getTriangleCount(); counts[2] = g_mesh->hasVertexNormals() ? 1 : 0;
    });
}
int gdbref_mesh_copy(double *vertices, double *normals, int *triangles)
{
    return guarded([&] {
        if (g_mesh == NULL) throw std::runtime_error("no mesh loaded");
        for (size_t i = 0; i < g_mesh->getVertexCount(); i++) for (int k = 0; k < 3; k++) {
            vertices[3 * i + k] = g_mesh->getVertexPositions()[i][k];
            if (g_mesh->hasVertexNormals()) normals[3 * i + k] = g_mesh->getVertexNormals()[i][k];
        }
        for (size_t t = 0; t < g_mesh->getTriangleCount(); t++) for (int k = 0; k < 3; k++) triangles[3 * t + k] = (int) g_mesh->getTriangles()[t].idx[k];
    });
}

}
