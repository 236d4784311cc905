// TEST INFRASTRUCTURE — CPU oracle for the G-PT tracer (fp64, scalar, OpenMP over row bands).
//
// A restatement of the reference's algorithm for the hot path, each function citing the
// reference file:line it follows:
//   src/integrators/gpt/gpt.cpp:84-114,176-231,242-369,397-436,439-463,468-1180,1220-1355
// plus the slice of Mitsuba the BASELINE scenes exercise: perspective sensor, rectangle /
// sphere / flat triangle shapes, nearest-hit and shadow-ray epsilons of the kd-tree front end,
// area emitters on rectangles, diffuse / roughconductor / conductor / dielectric BSDFs,
// box reconstruction filter, ImageBlock::put and MultiFilm::developMulti semantics.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this.  The product (libgdb200.so) never links or calls it.
//
// PARITY PINNED against the reference itself: oracle/_ref/libref_mitsuba.so is the reference's gpt.cpp with the scene,
// kd-tree, shape, emitter, BSDF, sensor, film and filter sources it runs on, compiled from /root/reference as they are
// (recipe and the stand-ins for the missing third-party headers: oracle/Makefile, oracle/refstubs).  tests/test_ref_gpt.py
// renders twenty-one scene / parameter cases with both on the same scene bytes and sample streams: every buffer agrees to
// 1e-11 with no differing pixel; tests/test_ref_mitsuba.py does the same per BSDF plugin (sample / eval / pdf, 1e-12).
// The reference ships no test, scene or golden image for gpt; the invariants of tests/test_gpt_oracle.py (primal equals a
// plain MIS path tracer, gradients are differences of the primal, weight bookkeeping) stay as independent checks.
// One deliberate deviation: gpt.cpp:957 leaves shiftedDRec.measure uninitialised (undefined behaviour); the intended
// ESolidAngle is used here and in the product.  gdb200_gpt_params.flags & GDB200_GPT_REF_UNINIT_MEASURE reproduces what the g++ build of the
// reference does instead (see the note at the use).
//
// Random numbers: the `gdb200_counter` sampler.  Sampler::generate(pixel) (gpt.cpp:1250-1251)
// re-keys a splitmix64 stream from (seed, pixel.x, pixel.y); next1D/next2D then consume that
// stream sequentially over all spp samples of the pixel (gpt never calls advance()).
#include "../include/gdb200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <vector>

namespace {

typedef double Float;
const Float Epsilon = 1e-7, ShadowEpsilon = 1e-5;          // constants.h:25-26 (DOUBLE_PRECISION)
const Float DeltaEpsilon = (Float)1e-3f;                   // constants.h:31 (a float literal)
const Float D_EPSILON = 1e-14;                             // gpt.cpp:63
const Float PI = 3.14159265358979323846, INV_PI = 0.31830988618379067154;
const Float INF = std::numeric_limits<Float>::infinity();

// ---------------------------------------------------------------- vectors / spectra
struct V3 { Float x, y, z; };
inline V3 v3(Float x, Float y, Float z) { V3 r = {x, y, z}; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, Float f) { return v3(a.x * f, a.y * f, a.z * f); }
inline V3 operator*(Float f, V3 a) { return v3(a.x * f, a.y * f, a.z * f); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }   // spectra
inline V3 operator/(V3 a, Float f) { Float r = (Float)1 / f; return v3(a.x * r, a.y * r, a.z * r); }  // vector.h:535-542
inline Float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline Float lengthSquared(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline Float length(V3 a) { return std::sqrt(lengthSquared(a)); }
inline V3 normalize(V3 a) { return a / length(a); }                               // vector.h:625-627
inline bool isZero(V3 a) { return a.x == 0 && a.y == 0 && a.z == 0; }
inline Float maxComp(V3 a) { return std::max(a.x, std::max(a.y, a.z)); }
typedef V3 Spec;
inline Spec spec(Float v) { return v3(v, v, v); }
inline Spec specOf(const double *p) { return v3(p[0], p[1], p[2]); }
inline Spec safeSqrt(Spec s) { return v3(std::sqrt(std::max(0.0, s.x)), std::sqrt(std::max(0.0, s.y)), std::sqrt(std::max(0.0, s.z))); }
inline Spec operator/(Spec a, Spec b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }

struct Frame { V3 s, t, n; };
inline V3 toLocal(const Frame &f, V3 v) { return v3(dot(v, f.s), dot(v, f.t), dot(v, f.n)); }     // frame.h:74-80
inline V3 toWorld(const Frame &f, V3 v) { return f.s * v.x + f.t * v.y + f.n * v.z; }             // frame.h:83-85

// transform.h:108-125 (points, projective), :128-137 (affine), :175-183 (vectors), :203-211 (normals)
inline V3 xfPoint(const double *m, V3 p)
{
    Float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    Float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    Float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    Float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w == 1.0) return v3(x, y, z);
    return v3(x, y, z) / w;
}
inline V3 xfAffine(const double *m, V3 p)
{
    return v3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
              m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
inline V3 xfVector(const double *m, V3 v)
{
    return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
              m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
inline V3 xfNormal(const double *inv, V3 v)
{
    return v3(inv[0] * v.x + inv[4] * v.y + inv[8] * v.z, inv[1] * v.x + inv[5] * v.y + inv[9] * v.z,
              inv[2] * v.x + inv[6] * v.y + inv[10] * v.z);
}

// util.cpp:592-601
inline void coordinateSystem(V3 a, V3 &b, V3 &c)
{
    if (std::abs(a.x) > std::abs(a.y)) {
        Float invLen = 1.0 / std::sqrt(a.x * a.x + a.z * a.z);
        c = v3(a.z * invLen, 0.0, -a.x * invLen);
    } else {
        Float invLen = 1.0 / std::sqrt(a.y * a.y + a.z * a.z);
        c = v3(0.0, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}
// util.cpp:603-608
inline void computeShadingFrame(V3 n, V3 dpdu, Frame &f)
{
    f.n = n;
    f.s = normalize(dpdu - f.n * dot(f.n, dpdu));
    f.t = cross(f.n, f.s);
}

// ---------------------------------------------------------------- sampler (gdb200_counter)
struct Sampler {
    uint64_t key, n;
    bool useForced = false; Float forced = 0;      // chi-square harness only: next1D() returns a supplied value
    static uint64_t mix(uint64_t z)
    {
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    void generate(uint64_t seed, int px, int py)   // Sampler::generate(const Point2i&)
    {
        key = mix(mix(seed + 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)(uint32_t)px | ((uint64_t)(uint32_t)py << 32)));
        n = 0;
    }
    Float next1D()
    {
        n++;
        if (useForced) return forced;
        return (Float)(mix(key + n * 0x9E3779B97F4A7C15ULL) >> 11) * (1.0 / 9007199254740992.0);
    }
    void next2D(Float &x, Float &y) { x = next1D(); y = next1D(); }
};

// ---------------------------------------------------------------- scene
enum { EDiffuseReflection = 0x1, EGlossyReflection = 0x4, EGlossyTransmission = 0x8, EDeltaReflection = 0x10, EDeltaTransmission = 0x20,
       ESmooth = 0xF, EDelta = 0x30, ETransmissionBits = 0x2 | 0x8 | 0x20, EBackSide = 0x20000, EFrontSide = 0x10000 };
enum Measure { ESolidAngle, EDiscrete };

struct Tri { V3 p0, p1, p2; int k; Float n_u, n_v, n_d, a_u, a_v, b_nu, b_nv, c_nu, c_nv; int shape; bool hasNormals; V3 n0, n1, n2; };

struct Shape {
    gdb200_shape d;
    V3 dpdu, dpdv; Frame frame; Float invArea;     // rectangle.cpp:100-110
};

// EnvironmentMap (envmap.cpp): top MIP level + the sampling tables of configure(), envmap.cpp:263-320
struct EnvMap {
    bool present; int w, h, emitter;
    std::vector<Float> texels;                 // Spectrum per texel (RGB -> Float)
    std::vector<float> cdfRows, cdfCols; std::vector<Float> rowWeights;
    Float scale, normalization, pixelSizeX, pixelSizeY;
    double toWorld[16], toObject[16];
    V3 center; Float radius;                   // m_sceneBSphere
    EnvMap() : present(false) {}
};
// TriMesh area sampling table (trimesh.cpp:388-403): per emitting mesh shape
struct MeshSampling { std::vector<Float> cdf; std::vector<V3> verts, normals; Float invSurfaceArea; };

struct Scene {
    gdb200_camera cam;
    Float invResX, invResY, filterRadius;
    std::vector<Shape> shapes;
    std::vector<gdb200_material> mats;
    std::vector<gdb200_emitter> ems;
    std::vector<Tri> tris;
    std::vector<Float> emCdf;      // DiscreteDistribution over samplingWeight (scene.cpp:357-380)
    Float emNormalization;
    EnvMap env;
    std::vector<MeshSampling> meshSampling;    // indexed by shape (empty for non-emitting shapes)
    std::vector<Float> fdrInt, fdrExt;         // per material: plastic.cpp:188-190
};

struct Its {
    Float t; V3 p; V3 geoN; Frame sh; V3 wi; int shape;
    bool valid() const { return t != INF; }
};

struct Ray { V3 o, d; Float mint, maxt; };
inline V3 at(const Ray &r, Float t) { return r.o + r.d * t; }

// triaccel.h:61-95
bool triLoad(Tri &T, V3 A, V3 B, V3 C)
{
    static const int waldModulo[4] = {1, 2, 0, 1};
    V3 b = C - A, c = B - A, N = cross(c, b);
    const Float Nv[3] = {N.x, N.y, N.z}, bv[3] = {b.x, b.y, b.z}, cv[3] = {c.x, c.y, c.z}, Av[3] = {A.x, A.y, A.z};
    int k = 0;
    for (int j = 0; j < 3; j++) if (std::abs(Nv[j]) > std::abs(Nv[k])) k = j;
    int u = waldModulo[k], v = waldModulo[k + 1];
    const Float n_k = Nv[k], denom = bv[u] * cv[v] - bv[v] * cv[u];
    T.p0 = A; T.p1 = B; T.p2 = C;
    if (denom == 0) { T.k = 3; return false; }
    T.k = k;
    T.n_u = Nv[u] / n_k; T.n_v = Nv[v] / n_k; T.n_d = dot(A, N) / n_k;
    T.b_nu = bv[u] / denom; T.b_nv = -bv[v] / denom; T.a_u = Av[u]; T.a_v = Av[v];
    T.c_nu = cv[v] / denom; T.c_nv = -cv[u] / denom;
    return true;
}
// triaccel.h:97-158
inline bool triIntersect(const Tri &T, const Ray &ray, Float mint, Float maxt, Float &u, Float &v, Float &t)
{
    Float o_u, o_v, o_k, d_u, d_v, d_k;
    switch (T.k) {
        case 0: o_u = ray.o.y; o_v = ray.o.z; o_k = ray.o.x; d_u = ray.d.y; d_v = ray.d.z; d_k = ray.d.x; break;
        case 1: o_u = ray.o.z; o_v = ray.o.x; o_k = ray.o.y; d_u = ray.d.z; d_v = ray.d.x; d_k = ray.d.y; break;
        case 2: o_u = ray.o.x; o_v = ray.o.y; o_k = ray.o.z; d_u = ray.d.x; d_v = ray.d.y; d_k = ray.d.z; break;
        default: return false;
    }
    t = (T.n_d - o_u * T.n_u - o_v * T.n_v - o_k) / (d_u * T.n_u + d_v * T.n_v + d_k);
    if (t < mint || t > maxt) return false;
    const Float hu = o_u + t * d_u - T.a_u, hv = o_v + t * d_v - T.a_v;
    u = hv * T.b_nu + hu * T.b_nv;
    v = hu * T.c_nu + hv * T.c_nv;
    return u >= 0 && v >= 0 && u + v <= 1.0;
}

// rectangle.cpp:125-151
inline bool rectIntersect(const Shape &s, const Ray &r, Float mint, Float maxt, Float &t)
{
    V3 o = xfAffine(s.d.to_object, r.o), d = xfVector(s.d.to_object, r.d);
    Float hit = -o.z / d.z;
    if (!(hit >= mint && hit <= maxt)) return false;
    V3 local = o + d * hit;
    if (std::abs(local.x) <= 1 && std::abs(local.y) <= 1) { t = hit; return true; }
    return false;
}

// util.cpp:487-525
inline bool solveQuadratic(double a, double b, double c, double &x0, double &x1)
{
    if (a == 0) { if (b != 0) { x0 = x1 = -c / b; return true; } return false; }
    double discrim = b * b - 4.0 * a * c;
    if (discrim < 0) return false;
    double temp, sqrtDiscrim = std::sqrt(discrim);
    if (b < 0) temp = -0.5 * (b - sqrtDiscrim); else temp = -0.5 * (b + sqrtDiscrim);
    x0 = temp / a; x1 = c / temp;
    if (x0 > x1) std::swap(x0, x1);
    return true;
}
// sphere.cpp:163-187
inline bool sphereIntersect(const Shape &s, const Ray &r, Float mint, Float maxt, Float &t)
{
    V3 o = r.o - specOf(s.d.center), d = r.d;
    double A = lengthSquared(d), B = 2 * dot(o, d), C = lengthSquared(o) - s.d.radius * s.d.radius;
    double nearT, farT;
    if (!solveQuadratic(A, B, C, nearT, farT)) return false;
    if (!(nearT <= maxt && farT >= mint)) return false;
    if (nearT < mint) { if (farT > maxt) return false; t = farT; } else t = nearT;
    return true;
}

// Nearest hit in [mint, maxt] over all primitives = what the kd-tree traversal returns
// (sahkdtree3.h:179-308: every candidate is tested against the shrinking [mint,maxt]).
// Returns primitive kind/index through (shape, tri, u, v).
bool nearestHit(const Scene &sc, const Ray &ray, Float mint, Float maxt, bool anyHit, Float &tOut, int &shapeOut,
                int &triOut, Float &uOut, Float &vOut)
{
    bool found = false;
    for (size_t i = 0; i < sc.shapes.size(); i++) {
        const Shape &s = sc.shapes[i];
        Float t;
        if (s.d.type == GDB200_SHAPE_RECTANGLE) {
            if (rectIntersect(s, ray, mint, maxt, t)) { if (anyHit) return true; maxt = t; found = true; shapeOut = (int)i; triOut = -1; }
        } else if (s.d.type == GDB200_SHAPE_SPHERE) {
            if (sphereIntersect(s, ray, mint, maxt, t)) { if (anyHit) return true; maxt = t; found = true; shapeOut = (int)i; triOut = -1; }
        }
    }
    for (size_t i = 0; i < sc.tris.size(); i++) {
        Float t, u, v;
        if (triIntersect(sc.tris[i], ray, mint, maxt, u, v, t)) {
            if (anyHit) return true;
            maxt = t; found = true; shapeOut = sc.tris[i].shape; triOut = (int)i; uOut = u; vOut = v;
        }
    }
    tOut = maxt;
    return found;
}

inline Float maxAbs(V3 o) { return std::max(std::max(std::abs(o.x), std::abs(o.y)), std::abs(o.z)); }

// ShapeKDTree::rayIntersect(ray, its): skdtree.cpp:112-147 + skdtree.h:343-428
bool rayIntersect(const Scene &sc, const Ray &ray, Its &its)
{
    its.t = INF;
    Float rayMinT = ray.mint;
    if (rayMinT == Epsilon) rayMinT *= std::max(maxAbs(ray.o), Epsilon);
    if (!(ray.maxt > rayMinT)) return false;
    Float t, u = 0, v = 0; int shape = -1, tri = -1;
    if (!nearestHit(sc, ray, rayMinT, ray.maxt, false, t, shape, tri, u, v)) return false;
    const Shape &s = sc.shapes[shape];
    its.t = t; its.shape = shape;
    V3 dpdu;
    if (tri >= 0) {                                            // skdtree.h:348-419, BarycentricPos = true
        const Tri &T = sc.tris[tri];
        const V3 b = v3(1 - u - v, u, v);
        its.p = T.p0 * b.x + T.p1 * b.y + T.p2 * b.z;
        V3 side1 = T.p1 - T.p0, side2 = T.p2 - T.p0, faceNormal = cross(side1, side2);
        Float len = length(faceNormal);
        if (!isZero(faceNormal)) faceNormal = faceNormal / len;
        dpdu = side1;
        if (T.hasNormals) {                                        // skdtree.h:383-394
            its.sh.n = normalize(T.n0 * b.x + T.n1 * b.y + T.n2 * b.z);
            if (dot(faceNormal, its.sh.n) < 0) faceNormal = -faceNormal;
        } else its.sh.n = faceNormal;
        its.geoN = faceNormal;
    } else if (s.d.type == GDB200_SHAPE_RECTANGLE) {           // rectangle.cpp:158-171
        its.geoN = s.frame.n;
        its.sh.n = s.frame.n;
        dpdu = s.dpdu;
        its.p = at(ray, its.t);
    } else {                                                   // sphere.cpp:197-240 (identity rotation)
        its.p = at(ray, its.t);
        V3 c = specOf(s.d.center), local = its.p - c;
        dpdu = v3(-local.y, local.x, 0) * (2 * PI);
        its.geoN = normalize(its.p - c);
        if (s.d.flip_normals) its.geoN = its.geoN * -1.0;
        its.sh.n = its.geoN;
    }
    computeShadingFrame(its.sh.n, dpdu, its.sh);               // skdtree.h:425
    its.wi = toLocal(its.sh, -ray.d);                          // skdtree.h:426
    return true;
}

// ShapeKDTree::rayIntersect(ray) — shadow rays: skdtree.cpp:206-226 (no Epsilon floor on the scale)
bool rayOccluded(const Scene &sc, const Ray &ray)
{
    Float rayMinT = ray.mint;
    if (rayMinT == Epsilon) rayMinT *= maxAbs(ray.o);
    if (!(ray.maxt > rayMinT)) return false;
    Float t, u, v; int a, b;
    return nearestHit(sc, ray, rayMinT, ray.maxt, true, t, a, b, u, v);
}

// ---------------------------------------------------------------- BSDFs
inline unsigned nestedType(const gdb200_material &m)
{
    switch (m.type) {
        case GDB200_BSDF_DIFFUSE:   // diffuse.cpp:97-101: no component at all when the reflectance is black
            return (std::max(m.reflectance[0], std::max(m.reflectance[1], m.reflectance[2])) > 0) ? (EDiffuseReflection | EFrontSide) : 0;
        case GDB200_BSDF_ROUGHCONDUCTOR: return EGlossyReflection | EFrontSide;
        case GDB200_BSDF_CONDUCTOR: return EDeltaReflection | EFrontSide;
        case GDB200_BSDF_PLASTIC: return EDeltaReflection | EDiffuseReflection | EFrontSide;     // plastic.cpp:210-214
        case GDB200_BSDF_ROUGHDIELECTRIC: return EGlossyReflection | EGlossyTransmission | EFrontSide | EBackSide;   // roughdielectric.cpp:246-252
        default: return EDeltaReflection | EDeltaTransmission | EFrontSide | EBackSide;
    }
}
inline int nestedComponentCount(const gdb200_material &m)
{
    if (m.type == GDB200_BSDF_DIELECTRIC || m.type == GDB200_BSDF_PLASTIC || m.type == GDB200_BSDF_ROUGHDIELECTRIC) return 2;
    if (m.type == GDB200_BSDF_DIFFUSE) return nestedType(m) ? 1 : 0;
    return 1;
}
// The plugins keep their roughness in a ConstantFloatTexture built from the (clamped) MicrofacetDistribution of the
// constructor and read it back through Spectrum::average() on every query (roughconductor.cpp:196,273-274;
// roughdielectric.cpp:206): three additions and a multiplication by the SINGLE-precision quotient 1.0f / 3
// (spectrum.h:481-486), so the roughness the distribution sees is alpha * (1 + 3e-8).
inline Float textureAverage(Float v) { Float result = 0.0f; for (int i = 0; i < 3; i++) result += v; return result * (1.0f / 3); }
inline Float effectiveAlpha(const gdb200_material &m) { return textureAverage(std::max((Float)m.alpha, (Float)1e-4f)); }

inline Float nestedRoughness(const gdb200_material &m, int component)
{
    switch (m.type) {
        case GDB200_BSDF_DIFFUSE: return INF;                               // diffuse.cpp:167-169
        case GDB200_BSDF_ROUGHCONDUCTOR: return 0.5f * (effectiveAlpha(m) + effectiveAlpha(m));  // roughconductor.cpp:437-440
        case GDB200_BSDF_ROUGHDIELECTRIC: return 0.5f * (effectiveAlpha(m) + effectiveAlpha(m)); // roughdielectric.cpp:642-645
        case GDB200_BSDF_PLASTIC: return component == 0 ? 0.0 : INF;        // plastic.cpp:442-449
        default: return 0.0;                                                // conductor.cpp:287, dielectric.cpp:393
    }
}
// twosided.cpp:85-101,205-210: the nested BRDF's components once as front-side and once as back-side components
inline unsigned bsdfType(const gdb200_material &m)
{
    unsigned t = nestedType(m);
    if (m.twosided && t) t |= EBackSide;
    return t;
}
inline int bsdfComponentCount(const gdb200_material &m) { return nestedComponentCount(m) * (m.twosided ? 2 : 1); }
inline Float bsdfRoughness(const gdb200_material &m, int component)
{
    const int n = nestedComponentCount(m);
    return nestedRoughness(m, component < n ? component : component - n);
}
inline Float bsdfEta(const gdb200_material &m) { return (m.type == GDB200_BSDF_DIELECTRIC || m.type == GDB200_BSDF_ROUGHDIELECTRIC) ? m.ior_ratio : 1.0; }  // bsdf.cpp:62-64 (plastic, twosided), dielectric.cpp:389

// warp.cpp:81-102
inline void squareToUniformDiskConcentric(Float sx, Float sy, Float &ox, Float &oy)
{
    Float r1 = 2.0 * sx - 1.0, r2 = 2.0 * sy - 1.0, phi, r;
    if (r1 == 0 && r2 == 0) { r = phi = 0; }
    else if (r1 * r1 > r2 * r2) { r = r1; phi = (PI / 4.0) * (r2 / r1); }
    else { r = r2; phi = (PI / 2.0) - (r1 / r2) * (PI / 4.0); }
    ox = r * std::cos(phi); oy = r * std::sin(phi);
}
// warp.cpp:43-52
inline V3 squareToCosineHemisphere(Float sx, Float sy)
{
    Float px, py;
    squareToUniformDiskConcentric(sx, sy, px, py);
    Float z = std::sqrt(std::max(0.0, 1.0 - px * px - py * py));
    if (z == 0) z = (Float)1e-10f;
    return v3(px, py, z);
}

// ---- microfacet.h (isotropic, sampleVisible = true)
struct Microfacet {
    int type; Float alpha;
    Microfacet(int t, Float a) : type(t), alpha(std::max(a, (Float)1e-4f)) {}        // microfacet.h:67-71
    Float eval(V3 m) const                                                            // :191-235
    {
        if (m.z <= 0) return 0.0;
        Float cosTheta2 = m.z * m.z;
        Float beckmannExponent = ((m.x * m.x) / (alpha * alpha) + (m.y * m.y) / (alpha * alpha)) / cosTheta2;
        Float result;
        if (type == GDB200_MICROFACET_BECKMANN)
            result = std::exp(-beckmannExponent) / (PI * alpha * alpha * cosTheta2 * cosTheta2);
        else {
            Float root = ((Float)1 + beckmannExponent) * cosTheta2;
            result = (Float)1 / (PI * alpha * alpha * root * root);
        }
        if (result * m.z < (Float)1e-20f) result = 0;
        return result;
    }
    Float smithG1(V3 v, V3 m) const                                                   // :470-508
    {
        if (dot(v, m) * v.z <= 0) return 0.0;
        Float temp = 1 - v.z * v.z;
        Float tanTheta = temp <= 0.0 ? 0.0 : std::abs(std::sqrt(temp) / v.z);         // frame.h tanTheta
        if (tanTheta == 0.0) return 1.0;
        if (type == GDB200_MICROFACET_BECKMANN) {
            Float a = 1.0 / (alpha * tanTheta);
            if (a >= (Float)1.6f) return 1.0;
            Float aSqr = a * a;
            return ((Float)3.535f * a + (Float)2.181f * aSqr) / (1.0 + (Float)2.276f * a + (Float)2.577f * aSqr);
        }
        Float root = alpha * tanTheta, r;                                             // math.cpp:89-101 hypot2(1, root)
        if (1.0 > std::abs(root)) { r = root / 1.0; r = 1.0 * std::sqrt(1.0 + r * r); }
        else if (root != 0.0) { r = 1.0 / root; r = std::abs(root) * std::sqrt(1.0 + r * r); }
        else r = 0.0;
        return 2.0 / (1.0 + r);
    }
    Float G(V3 wi, V3 wo, V3 m) const { return smithG1(wi, m) * smithG1(wo, m); }     // :514-516
    Float pdfVisible(V3 wi, V3 m) const                                               // :455-459
    {
        if (wi.z == 0) return 0.0;
        return smithG1(wi, m) * std::abs(dot(wi, m)) * eval(m) / std::abs(wi.z);
    }
    void sampleVisible11(Float thetaI, Float sx, Float sy, Float &slopeX, Float &slopeY) const   // :573-696
    {
        const Float SQRT_PI_INV = 1 / std::sqrt(PI);
        if (type == GDB200_MICROFACET_BECKMANN) {
            if (thetaI < (Float)1e-4f) {
                Float r = std::sqrt(-std::log(1.0 - sx));
                Float sinPhi = std::sin(2 * PI * sy), cosPhi = std::cos(2 * PI * sy);
                slopeX = r * cosPhi; slopeY = r * sinPhi; return;
            }
            Float tanThetaI = std::tan(thetaI), cotThetaI = 1 / tanThetaI;
            Float a = -1, c = mtsErf(cotThetaI);
            Float sample_x = std::max(sx, (Float)1e-6f);
            Float fit = 1 + thetaI * ((Float)-0.876f + thetaI * ((Float)0.4265f - (Float)0.0594f * thetaI));
            Float b = c - (1 + c) * std::pow(1 - sample_x, fit);
            Float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * std::exp(-cotThetaI * cotThetaI));
            int it = 0;
            while (++it < 10) {
                if (!(b >= a && b <= c)) b = 0.5 * (a + c);
                Float invErf = erfinv(b);
                Float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * std::exp(-invErf * invErf)) - sample_x;
                Float derivative = normalization * (1 - invErf * tanThetaI);
                if (std::abs(value) < (Float)1e-5f) break;
                if (value > 0) c = b; else a = b;
                b -= value / derivative;
            }
            slopeX = erfinv(b);
            slopeY = erfinv(2.0 * std::max(sy, (Float)1e-6f) - 1.0);
            return;
        }
        if (thetaI < (Float)1e-4f) {
            Float r = std::sqrt(std::max(0.0, sx / (1 - sx)));
            Float sinPhi = std::sin(2 * PI * sy), cosPhi = std::cos(2 * PI * sy);
            slopeX = r * cosPhi; slopeY = r * sinPhi; return;
        }
        Float tanThetaI = std::tan(thetaI);
        Float a = 1 / tanThetaI;
        Float G1 = 2.0 / (1.0 + std::sqrt(std::max(0.0, 1.0 + 1.0 / (a * a))));
        Float A = 2.0 * sx / G1 - 1.0;
        if (std::abs(A) == 1) A -= std::copysign(1.0, A) * Epsilon;
        Float tmp = 1.0 / (A * A - 1.0);
        Float B = tanThetaI;
        Float D = std::sqrt(std::max(0.0, B * B * tmp * tmp - (A * A - B * B) * tmp));
        Float slope_x_1 = B * tmp - D, slope_x_2 = B * tmp + D;
        slopeX = (A < 0.0 || slope_x_2 > 1.0 / tanThetaI) ? slope_x_1 : slope_x_2;
        Float S;
        if (sy > 0.5) { S = 1.0; sy = 2.0 * (sy - 0.5); } else { S = -1.0; sy = 2.0 * (0.5 - sy); }
        Float z = (sy * (sy * (sy * (-(Float)0.365728915865723) + (Float)0.790235037209296) - (Float)0.424965825137544) + (Float)0.000152998850436920) /
                  (sy * (sy * (sy * (sy * (Float)0.169507819808272 - (Float)0.397203533833404) - (Float)0.232500544458471) + (Float)1) - (Float)0.539825872510702);
        slopeY = S * z * std::sqrt(1.0 + slopeX * slopeX);
    }
    // math.cpp:55-72 math::erf (Abramowitz & Stegun 7.1.26, not libm's erf)
    static Float mtsErf(Float x)
    {
        const Float a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027, a5 = 1.061405429, p = 0.3275911;
        Float sign = std::copysign(1.0, x);
        x = std::abs(x);
        Float t = 1.0 / (1.0 + p * x);
        Float y = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * std::exp(-x * x);
        return sign * y;
    }
    // math.cpp:25-53 math::erfinv (Giles' polynomial)
    static Float erfinv(Float x)
    {
        Float w = -std::log(((Float)1 - x) * ((Float)1 + x)), p;
        if (w < (Float)5) {
            w = w - (Float)2.5;
            p = (Float)2.81022636e-08; p = (Float)3.43273939e-07 + p * w; p = (Float)-3.5233877e-06 + p * w;
            p = (Float)-4.39150654e-06 + p * w; p = (Float)0.00021858087 + p * w; p = (Float)-0.00125372503 + p * w;
            p = (Float)-0.00417768164 + p * w; p = (Float)0.246640727 + p * w; p = (Float)1.50140941 + p * w;
        } else {
            w = std::sqrt(w) - (Float)3;
            p = (Float)-0.000200214257; p = (Float)0.000100950558 + p * w; p = (Float)0.00134934322 + p * w;
            p = (Float)-0.00367342844 + p * w; p = (Float)0.00573950773 + p * w; p = (Float)-0.0076224613 + p * w;
            p = (Float)0.00943887047 + p * w; p = (Float)1.00167406 + p * w; p = (Float)2.83297682 + p * w;
        }
        return p * x;
    }
    V3 sampleVisible(V3 _wi, Float sx, Float sy) const                                // :421-452
    {
        V3 wi = normalize(v3(alpha * _wi.x, alpha * _wi.y, _wi.z));
        Float theta = 0, phi = 0;
        if (wi.z < (Float)0.99999) { theta = std::acos(wi.z); phi = std::atan2(wi.y, wi.x); }
        Float sinPhi = std::sin(phi), cosPhi = std::cos(phi);
        Float slx, sly;
        sampleVisible11(theta, sx, sy, slx, sly);
        Float rx = cosPhi * slx - sinPhi * sly, ry = sinPhi * slx + cosPhi * sly;
        rx *= alpha; ry *= alpha;
        Float normalization = (Float)1 / std::sqrt(rx * rx + ry * ry + (Float)1.0);
        return v3(-rx * normalization, -ry * normalization, normalization);
    }
};

// util.cpp:739-761
inline Spec fresnelConductorExact(Float cosThetaI, Spec eta, Spec k)
{
    Float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1 - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    Spec temp1 = eta * eta - k * k - spec(sinThetaI2);
    Spec a2pb2 = safeSqrt(temp1 * temp1 + k * k * eta * eta * 4.0);
    Spec a = safeSqrt((a2pb2 + temp1) * 0.5);
    Spec term1 = a2pb2 + spec(cosThetaI2), term2 = a * (2 * cosThetaI);
    Spec Rs2 = (term1 - term2) / (term1 + term2);
    Spec term3 = a2pb2 * cosThetaI2 + spec(sinThetaI4), term4 = term2 * sinThetaI2;
    Spec Rp2 = Rs2 * (term3 - term4) / (term3 + term4);
    return 0.5 * (Rp2 + Rs2);
}
// util.cpp:651-681
inline Float fresnelDielectricExt(Float cosThetaI_, Float &cosThetaT_, Float eta)
{
    if (eta == 1) { cosThetaT_ = -cosThetaI_; return 0.0; }
    Float scale = (cosThetaI_ > 0) ? 1 / eta : eta, cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0) { cosThetaT_ = 0.0; return 1.0; }
    Float cosThetaI = std::abs(cosThetaI_), cosThetaT = std::sqrt(cosThetaTSqr);
    Float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    Float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5 * (Rs * Rs + Rp * Rp);
}
inline V3 reflectLocal(V3 wi) { return v3(-wi.x, -wi.y, wi.z); }
inline V3 refractLocal(const gdb200_material &m, V3 wi, Float cosThetaT)              // dielectric.cpp:223-226
{
    Float scale = -(cosThetaT < 0 ? 1.0 / m.ior_ratio : m.ior_ratio);
    return v3(scale * wi.x, scale * wi.y, cosThetaT);
}

struct PlasticTerms { Float fdrInt, specularSamplingWeight, invEta2; };
// plastic.cpp:186-206 (fdrInt by the adaptive quadrature of util.cpp:855-859, tabulated per material at scene build)
const std::vector<Float> *g_fdrInt = 0;   // set by the render entry points before tracing (index = material pointer offset)
const gdb200_material *g_matBase = 0;
inline PlasticTerms plasticTerms(const gdb200_material &m)
{
    PlasticTerms t;
    t.fdrInt = (*g_fdrInt)[&m - g_matBase];
    const Float dAvg = m.reflectance[0] * 0.212671f + m.reflectance[1] * 0.715160f + m.reflectance[2] * 0.072169f;
    const Float sAvg = m.specular_reflectance[0] * 0.212671f + m.specular_reflectance[1] * 0.715160f + m.specular_reflectance[2] * 0.072169f;
    t.specularSamplingWeight = sAvg / (dAvg + sAvg);
    t.invEta2 = 1 / (m.ior_ratio * m.ior_ratio);
    return t;
}
inline Spec plasticDiffuse(const gdb200_material &m, const PlasticTerms &t)         // plastic.cpp:266-271
{
    Spec diff = specOf(m.reflectance);
    if (m.nonlinear) diff = diff / (spec(1.0f) - diff * t.fdrInt);
    else diff = diff / (1 - t.fdrInt);
    return diff;
}

Spec nestedEval(const gdb200_material &m, V3 wi, V3 wo, Measure measure)
{
    switch (m.type) {
    case GDB200_BSDF_DIFFUSE:                                                          // diffuse.cpp:110-119
        if (measure != ESolidAngle || wi.z <= 0 || wo.z <= 0) return spec(0);
        return specOf(m.reflectance) * (INV_PI * wo.z);
    case GDB200_BSDF_ROUGHCONDUCTOR: {                                                 // roughconductor.cpp:256-292
        if (measure != ESolidAngle || wi.z <= 0 || wo.z <= 0) return spec(0);
        V3 H = normalize(wo + wi);
        Microfacet distr(m.distribution, effectiveAlpha(m));
        const Float D = distr.eval(H);
        if (D == 0) return spec(0);
        const Spec F = fresnelConductorExact(dot(wi, H), specOf(m.eta), specOf(m.k)) * specOf(m.specular_reflectance);
        const Float G = distr.G(wi, wo, H);
        Float model = D * G / (4.0 * wi.z);
        return F * model;
    }
    case GDB200_BSDF_CONDUCTOR:                                                        // conductor.cpp:221-235
        if (measure != EDiscrete || wi.z <= 0 || wo.z <= 0 || std::abs(dot(reflectLocal(wi), wo) - 1) > DeltaEpsilon) return spec(0);
        return specOf(m.specular_reflectance) * fresnelConductorExact(wi.z, specOf(m.eta), specOf(m.k));
    case GDB200_BSDF_ROUGHDIELECTRIC: {                                                // roughdielectric.cpp:277-351 (mode ERadiance)
        if (measure != ESolidAngle || wi.z == 0) return spec(0);
        const Float m_eta = m.ior_ratio, m_invEta = 1 / m.ior_ratio;
        bool reflect = wi.z * wo.z > 0;
        V3 H;
        if (reflect) H = normalize(wo + wi);
        else { Float eta = wi.z > 0 ? m_eta : m_invEta; H = normalize(wi + wo * eta); }
        H = H * std::copysign(1.0, H.z);
        Microfacet distr(m.distribution, effectiveAlpha(m));
        const Float D = distr.eval(H);
        if (D == 0) return spec(0);
        Float unused; const Float F = fresnelDielectricExt(dot(wi, H), unused, m_eta);
        const Float G = distr.G(wi, wo, H);
        if (reflect) {
            Float value = F * D * G / (4.0f * std::abs(wi.z));
            return specOf(m.specular_reflectance) * value;
        } else {
            Float eta = wi.z > 0.0f ? m_eta : m_invEta;
            Float sqrtDenom = dot(wi, H) + eta * dot(wo, H);
            Float value = ((1 - F) * D * G * eta * eta * dot(wi, H) * dot(wo, H)) / (wi.z * sqrtDenom * sqrtDenom);
            Float factor = wi.z > 0 ? m_invEta : m_eta;
            return specOf(m.specular_transmittance) * std::abs(value * factor * factor);
        }
    }
    case GDB200_BSDF_PLASTIC: {                                                        // plastic.cpp:243-275 (typeMask EAll, component -1)
        const bool hasSpecular = measure == EDiscrete, hasDiffuse = measure == ESolidAngle;
        if (wo.z <= 0 || wi.z <= 0) return spec(0);
        Float unused; const Float Fi = fresnelDielectricExt(wi.z, unused, m.ior_ratio);
        const PlasticTerms t = plasticTerms(m);
        if (hasSpecular) {
            if (std::abs(dot(reflectLocal(wi), wo) - 1) < DeltaEpsilon) return specOf(m.specular_reflectance) * Fi;
        } else if (hasDiffuse) {
            const Float Fo = fresnelDielectricExt(wo.z, unused, m.ior_ratio);
            return plasticDiffuse(m, t) * ((INV_PI * wo.z) * t.invEta2 * (1 - Fi) * (1 - Fo));
        }
        return spec(0);
    }
    default: {                                                                         // dielectric.cpp:228-254
        bool discrete = measure == EDiscrete;
        Float cosThetaT, F = fresnelDielectricExt(wi.z, cosThetaT, m.ior_ratio);
        if (wi.z * wo.z >= 0) {
            if (!discrete || std::abs(dot(reflectLocal(wi), wo) - 1) > DeltaEpsilon) return spec(0);
            return specOf(m.specular_reflectance) * F;
        }
        if (!discrete || std::abs(dot(refractLocal(m, wi, cosThetaT), wo) - 1) > DeltaEpsilon) return spec(0);
        Float factor = cosThetaT < 0 ? 1.0 / m.ior_ratio : m.ior_ratio;
        return specOf(m.specular_transmittance) * factor * factor * (1 - F);
    }
    }
}

Float nestedPdf(const gdb200_material &m, V3 wi, V3 wo, Measure measure)
{
    switch (m.type) {
    case GDB200_BSDF_DIFFUSE:                                                          // diffuse.cpp:121-129
        if (measure != ESolidAngle || wi.z <= 0 || wo.z <= 0) return 0.0;
        return INV_PI * wo.z;
    case GDB200_BSDF_ROUGHCONDUCTOR: {                                                 // roughconductor.cpp:294-320
        if (measure != ESolidAngle || wi.z <= 0 || wo.z <= 0) return 0.0;
        V3 H = normalize(wo + wi);
        Microfacet distr(m.distribution, effectiveAlpha(m));
        return distr.eval(H) * distr.smithG1(wi, H) / (4.0 * wi.z);
    }
    case GDB200_BSDF_CONDUCTOR:                                                        // conductor.cpp:237-250
        if (measure != EDiscrete || wi.z <= 0 || wo.z <= 0 || std::abs(dot(reflectLocal(wi), wo) - 1) > DeltaEpsilon) return 0.0;
        return 1.0;
    case GDB200_BSDF_ROUGHDIELECTRIC: {                                                // roughdielectric.cpp:353-422
        if (measure != ESolidAngle) return 0.0;
        const Float m_eta = m.ior_ratio, m_invEta = 1 / m.ior_ratio;
        bool reflect = wi.z * wo.z > 0;
        V3 H; Float dwh_dwo;
        if (reflect) { H = normalize(wo + wi); dwh_dwo = 1.0f / (4.0f * dot(wo, H)); }
        else {
            Float eta = wi.z > 0 ? m_eta : m_invEta;
            H = normalize(wi + wo * eta);
            Float sqrtDenom = dot(wi, H) + eta * dot(wo, H);
            dwh_dwo = (eta * eta * dot(wo, H)) / (sqrtDenom * sqrtDenom);
        }
        H = H * std::copysign(1.0, H.z);
        Microfacet sampleDistr(m.distribution, effectiveAlpha(m));
        Float prob = sampleDistr.pdfVisible(wi * std::copysign(1.0, wi.z), H);
        Float unused; Float F = fresnelDielectricExt(dot(wi, H), unused, m_eta);
        prob *= reflect ? F : (1 - F);
        return std::abs(prob * dwh_dwo);
    }
    case GDB200_BSDF_PLASTIC: {                                                        // plastic.cpp:277-302
        if (wo.z <= 0 || wi.z <= 0) return 0.0;
        Float unused; const Float Fi = fresnelDielectricExt(wi.z, unused, m.ior_ratio);
        const PlasticTerms t = plasticTerms(m);
        const Float probSpecular = (Fi * t.specularSamplingWeight) / (Fi * t.specularSamplingWeight + (1 - Fi) * (1 - t.specularSamplingWeight));
        if (measure == EDiscrete) {
            if (std::abs(dot(reflectLocal(wi), wo) - 1) < DeltaEpsilon) return probSpecular;
        } else if (measure == ESolidAngle) return (INV_PI * wo.z) * (1 - probSpecular);
        return 0.0;
    }
    default: {                                                                         // dielectric.cpp:256-275
        bool discrete = measure == EDiscrete;
        Float cosThetaT, F = fresnelDielectricExt(wi.z, cosThetaT, m.ior_ratio);
        if (wi.z * wo.z >= 0) {
            if (!discrete || std::abs(dot(reflectLocal(wi), wo) - 1) > DeltaEpsilon) return 0.0;
            return F;
        }
        if (!discrete || std::abs(dot(refractLocal(m, wi, cosThetaT), wo) - 1) > DeltaEpsilon) return 0.0;
        return 1 - F;
    }
    }
}

struct BSDFSample { V3 wi, wo; Float eta; unsigned sampledType; Spec weight; Float pdf; };

// BSDF::sample(bRec, pdf, sample) with pdf pre-set to 0 by the caller (gpt.cpp:450-457)
void nestedSample(const gdb200_material &m, BSDFSample &r, Float sx, Float sy, Sampler &sampler)
{
    r.weight = spec(0); r.pdf = 0; r.eta = 1.0; r.sampledType = 0; r.wo = v3(0, 0, 0);
    switch (m.type) {
    case GDB200_BSDF_DIFFUSE:                                                          // diffuse.cpp:143-153
        if (r.wi.z <= 0) return;
        r.wo = squareToCosineHemisphere(sx, sy);
        r.sampledType = EDiffuseReflection;
        r.pdf = INV_PI * r.wo.z;
        r.weight = specOf(m.reflectance);
        return;
    case GDB200_BSDF_ROUGHCONDUCTOR: {                                                 // roughconductor.cpp:369-419
        if (r.wi.z < 0) return;
        Microfacet distr(m.distribution, effectiveAlpha(m));
        V3 mm = distr.sampleVisible(r.wi, sx, sy);
        Float temporaryPdf = distr.pdfVisible(r.wi, mm);
        if (temporaryPdf == 0) return;
        r.wo = 2 * dot(r.wi, mm) * mm - r.wi;
        r.sampledType = EGlossyReflection;
        if (r.wo.z <= 0) return;
        Spec F = fresnelConductorExact(dot(r.wi, mm), specOf(m.eta), specOf(m.k)) * specOf(m.specular_reflectance);
        Float weight = distr.smithG1(r.wo, mm);
        if (weight > 0) { r.pdf = temporaryPdf / (4.0 * dot(r.wo, mm)); r.weight = F * weight; }
        return;
    }
    case GDB200_BSDF_CONDUCTOR:                                                        // conductor.cpp:268-285
        if (r.wi.z <= 0) return;
        r.sampledType = EDeltaReflection;
        r.wo = reflectLocal(r.wi);
        r.pdf = 1;
        r.weight = specOf(m.specular_reflectance) * fresnelConductorExact(r.wi.z, specOf(m.eta), specOf(m.k));
        return;
    case GDB200_BSDF_ROUGHDIELECTRIC: {                                                // roughdielectric.cpp:505-614 (pdf-returning overload)
        const Float m_eta = m.ior_ratio, m_invEta = 1 / m.ior_ratio;
        Microfacet distr(m.distribution, effectiveAlpha(m));
        const V3 wiS = r.wi * std::copysign(1.0, r.wi.z);
        const V3 mm = distr.sampleVisible(wiS, sx, sy);
        const Float microfacetPDF = distr.pdfVisible(wiS, mm);
        if (microfacetPDF == 0) return;
        float temporaryPdf = microfacetPDF;                                            // :533, a float in the reference
        Float cosThetaT;
        Float F = fresnelDielectricExt(dot(r.wi, mm), cosThetaT, m_eta);
        Spec weight = spec(1.0f);
        bool sampleReflection = true;
        if (sampler.next1D() > F) { sampleReflection = false; temporaryPdf *= 1 - F; } else { temporaryPdf *= F; }   // EUsesSampler
        Float dwh_dwo;
        if (sampleReflection) {
            r.wo = 2 * dot(r.wi, mm) * mm - r.wi;                                      // reflect(wi, m), util.cpp:763-765
            r.eta = 1.0f; r.sampledType = EGlossyReflection;
            if (r.wi.z * r.wo.z <= 0) return;
            weight = weight * specOf(m.specular_reflectance);
            dwh_dwo = 1.0f / (4.0f * dot(r.wo, mm));
        } else {
            if (cosThetaT == 0) return;
            Float e = cosThetaT < 0 ? 1 / m_eta : m_eta;                               // refract(wi, m, eta, cosThetaT), util.cpp:767-772
            r.wo = mm * (dot(r.wi, mm) * e + cosThetaT) - r.wi * e;
            r.eta = cosThetaT < 0 ? m_eta : m_invEta; r.sampledType = EGlossyTransmission;
            if (r.wi.z * r.wo.z >= 0) return;
            Float factor = cosThetaT < 0 ? m_invEta : m_eta;
            weight = weight * (specOf(m.specular_transmittance) * (factor * factor));
            Float sqrtDenom = dot(r.wi, mm) + r.eta * dot(r.wo, mm);
            dwh_dwo = (r.eta * r.eta * dot(r.wo, mm)) / (sqrtDenom * sqrtDenom);
        }
        weight = weight * distr.smithG1(r.wo, mm);
        temporaryPdf *= std::abs(dwh_dwo);
        r.pdf = temporaryPdf; r.weight = weight;
        return;
    }
    case GDB200_BSDF_PLASTIC: {                                                        // plastic.cpp:372-414 (both components requested)
        if (r.wi.z <= 0) return;
        Float unused; const Float Fi = fresnelDielectricExt(r.wi.z, unused, m.ior_ratio);
        const PlasticTerms t = plasticTerms(m);
        const Float probSpecular = (Fi * t.specularSamplingWeight) / (Fi * t.specularSamplingWeight + (1 - Fi) * (1 - t.specularSamplingWeight));
        if (sx < probSpecular) {
            r.sampledType = EDeltaReflection; r.wo = reflectLocal(r.wi);
            r.pdf = probSpecular;
            r.weight = specOf(m.specular_reflectance) * Fi / probSpecular;
        } else {
            r.sampledType = EDiffuseReflection;
            r.wo = squareToCosineHemisphere((sx - probSpecular) / (1 - probSpecular), sy);
            const Float Fo = fresnelDielectricExt(r.wo.z, unused, m.ior_ratio);
            r.pdf = (1 - probSpecular) * (INV_PI * r.wo.z);
            r.weight = plasticDiffuse(m, t) * (t.invEta2 * (1 - Fi) * (1 - Fo) / (1 - probSpecular));
        }
        return;
    }
    default: {                                                                         // dielectric.cpp:277-305
        Float cosThetaT, F = fresnelDielectricExt(r.wi.z, cosThetaT, m.ior_ratio);
        if (sx <= F) {
            r.sampledType = EDeltaReflection; r.wo = reflectLocal(r.wi); r.eta = 1.0; r.pdf = F;
            r.weight = specOf(m.specular_reflectance);
        } else {
            r.sampledType = EDeltaTransmission; r.wo = refractLocal(m, r.wi, cosThetaT);
            r.eta = cosThetaT < 0 ? m.ior_ratio : 1.0 / m.ior_ratio; r.pdf = 1 - F;
            Float factor = cosThetaT < 0 ? 1.0 / m.ior_ratio : m.ior_ratio;
            r.weight = specOf(m.specular_transmittance) * (factor * factor);
        }
        return;
    }
    }
}

// TwoSidedBRDF::eval / pdf / sample (twosided.cpp:109-183) around the nested BRDF (same BRDF on both sides)
Spec bsdfEval(const gdb200_material &m, V3 wi, V3 wo, Measure measure)
{
    if (!m.twosided || wi.z > 0) return nestedEval(m, wi, wo, measure);
    wi.z *= -1; wo.z *= -1;
    return nestedEval(m, wi, wo, measure);
}
Float bsdfPdf(const gdb200_material &m, V3 wi, V3 wo, Measure measure)
{
    if (!m.twosided || wi.z > 0) return nestedPdf(m, wi, wo, measure);
    wi.z *= -1; wo.z *= -1;
    return nestedPdf(m, wi, wo, measure);
}
void bsdfSample(const gdb200_material &m, BSDFSample &r, Float sx, Float sy, Sampler &sampler)
{
    bool flipped = false;
    if (m.twosided && r.wi.z < 0) { r.wi.z *= -1; flipped = true; }
    nestedSample(m, r, sx, sy, sampler);
    if (flipped) {
        r.wi.z *= -1;
        if (!isZero(r.weight) && r.pdf != 0) r.wo.z *= -1;
    }
}

// ---------------------------------------------------------------- emitters
struct DRec { V3 ref, refN, p, n, d; Float dist, pdf; int emitter; bool discrete; };   // DirectSamplingRecord (measure: solid angle unless `discrete`)

inline const gdb200_material &matOf(const Scene &sc, const Its &its) { return sc.mats[sc.shapes[its.shape].d.material]; }
inline bool isEmitter(const Scene &sc, const Its &its) { return sc.shapes[its.shape].d.emitter >= 0; }
// Intersection::Le -> AreaLight::eval, area.cpp:104-109
inline Spec emittedLe(const Scene &sc, const Its &its, V3 d)
{
    if (dot(its.sh.n, d) <= 0) return spec(0);
    return specOf(sc.ems[sc.shapes[its.shape].d.emitter].radiance);
}
// records.inl:160-165
inline void initDRec(const Scene &sc, const Its &ref, DRec &r)
{
    r.ref = ref.p;
    r.refN = v3(0, 0, 0);
    if ((bsdfType(matOf(sc, ref)) & (ETransmissionBits | EBackSide)) == 0) r.refN = ref.sh.n;
}

// DiscreteDistribution::sampleReuse, pmf.h:124-188
inline size_t pmfSampleReuse(const Float *cdf, size_t n, Float &sampleValue, Float &pdf)
{
    size_t entry = std::lower_bound(cdf, cdf + n + 1, sampleValue) - cdf;
    size_t index = std::min(n - 1, (size_t)std::max((ptrdiff_t)0, (ptrdiff_t)entry - 1));
    while (cdf[index + 1] - cdf[index] == 0 && index < n) ++index;
    pdf = cdf[index + 1] - cdf[index];
    sampleValue = (sampleValue - cdf[index]) / (cdf[index + 1] - cdf[index]);
    return index;
}

// ---- EnvironmentMap, envmap.cpp
const Float INV_TWOPI = 0.15915494309189533577;
inline Float luminance(Spec s) { return s.x * 0.212671f + s.y * 0.715160f + s.z * 0.072169f; }   // spectrum.h:725-727
inline Float safe_acos(Float v) { return std::acos(std::min(1.0, std::max(-1.0, v))); }
inline int floorToInt(Float v) { return (int)std::floor(v); }
inline int modulo(int a, int b) { int r = a % b; return (r < 0) ? r + b : r; }
// MIPMap::evalTexel(0, x, y) with ERepeat in u and EClamp in v (mipmap.h:503-566, envmap.cpp:178-179)
inline Spec envTexel(const EnvMap &e, int x, int y)
{
    if (x < 0 || x >= e.w) x = modulo(x, e.w);
    if (y < 0 || y >= e.h) y = std::min(std::max(y, 0), e.h - 1);
    const Float *t = &e.texels[((size_t)y * e.w + x) * 3];
    return v3(t[0], t[1], t[2]);
}
// MIPMap::evalBilinear(0, uv), mipmap.h:575-596
inline Spec envBilinear(const EnvMap &e, Float uvx, Float uvy)
{
    if (!std::isfinite(uvx) || !std::isfinite(uvy)) return spec(0);
    Float u = uvx * e.w - 0.5f, v = uvy * e.h - 0.5f;
    int xPos = floorToInt(u), yPos = floorToInt(v);
    Float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
    return envTexel(e, xPos, yPos) * dx2 * dy2 + envTexel(e, xPos, yPos + 1) * dx2 * dy1
         + envTexel(e, xPos + 1, yPos) * dx1 * dy2 + envTexel(e, xPos + 1, yPos + 1) * dx1 * dy1;
}
// EnvironmentMap::evalEnvironment, envmap.cpp:385-409.  DEVIATION: the reference filters primary-ray lookups with
// ray differentials (EWA over the MIP pyramid); every lookup here is the differential-free branch (:393-396).
Spec evalEnvironment(const Scene &sc, V3 d)
{
    const EnvMap &e = sc.env;
    V3 v = xfVector(e.toObject, d);
    return envBilinear(e, std::atan2(v.x, -v.z) * INV_TWOPI, safe_acos(v.y) * INV_PI) * e.scale;
}
// envmap.cpp:660-665
inline uint32_t envSampleReuse(const float *cdf, uint32_t size, Float &sample)
{
    const float *entry = std::lower_bound(cdf, cdf + size + 1, (float)sample);
    uint32_t index = std::min((uint32_t)std::max((ptrdiff_t)0, entry - cdf - 1), size - 1);
    sample = (sample - (Float)cdf[index]) / (Float)(cdf[index + 1] - cdf[index]);
    return index;
}
inline Float intervalToTent(Float sample)                                            // warp.cpp:143-155
{
    Float sign;
    if (sample < 0.5f) { sign = 1; sample *= 2; } else { sign = -1; sample = 2 * (sample - 0.5f); }
    return sign * (1 - std::sqrt(sample));
}
// envmap.cpp:571-608
void envSampleDirection(const EnvMap &e, Float sx, Float sy, V3 &d, Spec &value, Float &pdf)
{
    uint32_t row = envSampleReuse(&e.cdfRows[0], e.h, sy), col = envSampleReuse(&e.cdfCols[(size_t)row * (e.w + 1)], e.w, sx);
    Float posx = (Float)col + intervalToTent(sx), posy = (Float)row + intervalToTent(sy);
    int xPos = floorToInt(posx), yPos = floorToInt(posy);
    Float dx1 = posx - xPos, dx2 = 1.0f - dx1, dy1 = posy - yPos, dy2 = 1.0f - dy1;
    Spec value1 = envTexel(e, xPos, yPos) * dx2 * dy2 + envTexel(e, xPos + 1, yPos) * dx1 * dy2;
    Spec value2 = envTexel(e, xPos, yPos + 1) * dx2 * dy1 + envTexel(e, xPos + 1, yPos + 1) * dx1 * dy1;
    value = (value1 + value2) * e.scale;
    pdf = (luminance(value1) * e.rowWeights[std::min(std::max(yPos, 0), e.h - 1)] + luminance(value2) * e.rowWeights[std::min(std::max(yPos + 1, 0), e.h - 1)]) * e.normalization;
    Float sinPhi = std::sin(e.pixelSizeX * (posx + 0.5f)), cosPhi = std::cos(e.pixelSizeX * (posx + 0.5f));
    Float sinTheta = std::sin(e.pixelSizeY * (posy + 0.5f)), cosTheta = std::cos(e.pixelSizeY * (posy + 0.5f));
    d = v3(sinPhi * sinTheta, cosTheta, -cosPhi * sinTheta);
    pdf /= std::max(std::abs(sinTheta), Epsilon);
}
// envmap.cpp:611-642
Float envPdfDirection(const EnvMap &e, V3 d)
{
    Float uvx = std::atan2(d.x, -d.z) * INV_TWOPI, uvy = safe_acos(d.y) * INV_PI;
    if (!std::isfinite(uvx) || !std::isfinite(uvy)) return 0.0;
    Float u = uvx * e.w - 0.5f, v = uvy * e.h - 0.5f;
    int xPos = floorToInt(u), yPos = floorToInt(v);
    Float dx1 = u - xPos, dx2 = 1.0f - dx1, dy1 = v - yPos, dy2 = 1.0f - dy1;
    Spec value1 = envTexel(e, xPos, yPos) * dx2 * dy2 + envTexel(e, xPos + 1, yPos) * dx1 * dy2;
    Spec value2 = envTexel(e, xPos, yPos + 1) * dx2 * dy1 + envTexel(e, xPos + 1, yPos + 1) * dx1 * dy1;
    Float sinTheta = std::sqrt(std::max(0.0, 1 - d.y * d.y));
    return (luminance(value1) * e.rowWeights[std::min(std::max(yPos, 0), e.h - 1)] + luminance(value2) * e.rowWeights[std::min(std::max(yPos + 1, 0), e.h - 1)])
        * e.normalization / std::max(std::abs(sinTheta), Epsilon);
}
// BSphere::rayIntersect, bsphere.h:88-95
inline bool bsphereIntersect(const EnvMap &e, V3 o, V3 d, Float &nearHit, Float &farHit)
{
    V3 oc = o - e.center;
    return solveQuadratic(lengthSquared(d), 2 * dot(oc, d), lengthSquared(oc) - e.radius * e.radius, nearHit, farHit);
}
// EnvironmentMap::fillDirectSamplingRecord, envmap.cpp:358-374
bool envFillDirectSamplingRecord(const Scene &sc, DRec &dRec, V3 o, V3 d)
{
    Float nearT, farT;
    if (!bsphereIntersect(sc.env, o, d, nearT, farT) || nearT > 0 || farT < 0) return false;
    dRec.p = o + d * farT;
    dRec.n = normalize(sc.env.center - dRec.p);
    dRec.d = d; dRec.dist = farT; dRec.emitter = sc.env.emitter;
    return true;
}

// SpotEmitter::falloffCurve, spot.cpp:105-125, for the constant "texture" (the default): d is the world direction
// leaving the light, brought into the light's frame by the inverse transform's 3x3 block (transform.h:175-183).
Float spotFalloff(const gdb200_emitter &em, V3 dWorld)
{
    const double *m = em.to_local;
    const V3 d = v3(m[0] * dWorld.x + m[1] * dWorld.y + m[2] * dWorld.z, m[3] * dWorld.x + m[4] * dWorld.y + m[5] * dWorld.z,
                    m[6] * dWorld.x + m[7] * dWorld.y + m[8] * dWorld.z);
    const Float cosTheta = d.z;
    const Float m_cosCutoffAngle = std::cos(em.cutoff_angle), m_cosBeamWidth = std::cos(em.beam_width);     // spot.cpp:89-94
    const Float m_invTransitionWidth = 1.0f / (em.cutoff_angle - em.beam_width);
    if (cosTheta <= m_cosCutoffAngle) return 0.0;
    if (cosTheta >= m_cosBeamWidth) return 1.0;
    return (em.cutoff_angle - std::acos(cosTheta)) * m_invTransitionWidth;
}

// Scene::sampleEmitterDirectVisible, scene.cpp:855-879 (+ pmf.h:124-188; area.cpp:158-176 over shape.cpp:102-114 with
// rectangle.cpp:210-216 or trimesh.cpp:412-423 / triangle.cpp:24-50; envmap.cpp:516-544).  Returns value = Le/pdf
// (0 when occluded).
Spec sampleEmitterDirectVisible(const Scene &sc, DRec &dRec, Float sx, Float sy, bool &visible)
{
    Float emPdf;
    size_t index = pmfSampleReuse(&sc.emCdf[0], sc.ems.size(), sx, emPdf);
    const gdb200_emitter &em = sc.ems[index];
    Spec value;
    dRec.emitter = (int)index;
    dRec.discrete = false;
    if (em.type == GDB200_EMITTER_ENVMAP) {                                         // envmap.cpp:516-544
        Spec v; V3 d; Float pdf;
        envSampleDirection(sc.env, sx, sy, d, v, pdf);
        V3 dw = xfVector(sc.env.toWorld, d);
        Float nearT = 0, farT = 0;
        if (isZero(v) || pdf == 0 || !bsphereIntersect(sc.env, dRec.ref, dw, nearT, farT) || nearT >= 0 || farT <= 0) {
            // the reference returns with p/d/dist unset and its caller still traces a shadow ray from them;
            // defined here as "no contribution" (unreachable for strictly positive maps seen from inside the sphere)
            dRec.pdf = 0.0; dRec.p = dRec.ref; dRec.n = v3(0, 0, 0); dRec.d = v3(0, 0, 1); dRec.dist = 0;
            visible = true;
            return spec(0);
        }
        dRec.pdf = pdf; dRec.p = dRec.ref + dw * farT; dRec.n = normalize(sc.env.center - dRec.p); dRec.dist = farT; dRec.d = dw;
        value = v / pdf;
    } else if (em.type == GDB200_EMITTER_POINT) {                                   // point.cpp:131-147
        dRec.p = specOf(em.position);
        dRec.pdf = 1.0f;
        dRec.d = dRec.p - dRec.ref;
        dRec.dist = length(dRec.d);
        Float invDist = 1.0f / dRec.dist;
        dRec.d = dRec.d * invDist;
        dRec.n = v3(0, 0, 0);
        dRec.discrete = true;
        value = specOf(em.radiance) * (invDist * invDist);
    } else if (em.type == GDB200_EMITTER_SPOT) {                                    // spot.cpp:184-200
        dRec.p = specOf(em.position);
        dRec.pdf = 1.0f;
        dRec.d = dRec.p - dRec.ref;
        dRec.dist = length(dRec.d);
        Float invDist = 1.0f / dRec.dist;
        dRec.d = dRec.d * invDist;
        dRec.n = v3(0, 0, 0);
        dRec.discrete = true;
        value = specOf(em.radiance) * spotFalloff(em, -dRec.d) * (invDist * invDist);
    } else {
        const Shape &s = sc.shapes[em.shape];
        if (s.d.type == GDB200_SHAPE_SPHERE) {                                      // Sphere::sampleDirect, sphere.cpp:283-355
            const V3 m_center = specOf(s.d.center); const Float m_radius = s.d.radius, m_invSurfaceArea = 1 / (4 * PI * m_radius * m_radius);
            const V3 refToCenter = m_center - dRec.ref;
            const Float refDist2 = lengthSquared(refToCenter);
            const Float invRefDist = static_cast<Float>(1) / std::sqrt(refDist2);
            const Float sinAlpha = m_radius * invRefDist;
            if (sinAlpha < 1 - Epsilon) {
                Float cosAlpha = std::sqrt(std::max(0.0, 1.0f - sinAlpha * sinAlpha));
                Float cosTheta = (1 - sx) + sx * cosAlpha, sinTheta = std::sqrt(std::max(0.0, 1.0f - cosTheta * cosTheta));   // warp.cpp:54-63
                Float sinPhi = std::sin(2.0f * PI * sy), cosPhi = std::cos(2.0f * PI * sy);
                Frame fr; fr.n = refToCenter * invRefDist; coordinateSystem(fr.n, fr.s, fr.t);
                dRec.d = toWorld(fr, v3(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta));
                dRec.pdf = INV_TWOPI / (1 - cosAlpha);
                const Float projDist = dot(refToCenter, dRec.d);
                const Float baseT = refDist2 / projDist;
                const V3 query = dRec.ref + dRec.d * baseT;
                const V3 queryToCenter = m_center - query;
                const Float queryDist2 = lengthSquared(queryToCenter), queryProjDist = dot(queryToCenter, dRec.d);
                Float A = 1.0f, B = -2 * queryProjDist, C = queryDist2 - m_radius * m_radius;
                double nearT, farT;
                if (!solveQuadratic(A, B, C, nearT, farT)) nearT = queryProjDist;
                dRec.dist = baseT + nearT;
                dRec.n = normalize(dRec.d * nearT - queryToCenter);
                dRec.p = m_center + dRec.n * m_radius;
            } else {
                Float z = 1.0f - 2.0f * sy, r = std::sqrt(std::max(0.0, 1.0f - z * z));           // squareToUniformSphere, warp.cpp:25-31
                V3 d = v3(r * std::cos(2.0f * PI * sx), r * std::sin(2.0f * PI * sx), z);
                dRec.p = m_center + d * m_radius; dRec.n = d;
                dRec.d = dRec.p - dRec.ref;
                Float dist2 = lengthSquared(dRec.d);
                dRec.dist = std::sqrt(dist2);
                dRec.d = dRec.d / dRec.dist;
                dRec.pdf = m_invSurfaceArea * dist2 / std::abs(dot(dRec.d, dRec.n));
            }
            if (s.d.flip_normals) dRec.n = dRec.n * -1.0;
        } else {
        if (s.d.type == GDB200_SHAPE_RECTANGLE) {                                   // rectangle.cpp:210-216
            dRec.p = xfPoint(s.d.to_world, v3(sx * 2 - 1, sy * 2 - 1, 0));
            dRec.n = s.frame.n;
            dRec.pdf = s.invArea;
        } else {                                                                    // trimesh.cpp:412-423, triangle.cpp:24-50
            const MeshSampling &ms = sc.meshSampling[em.shape];
            Float triPdf;
            size_t tri = pmfSampleReuse(&ms.cdf[0], ms.cdf.size() - 1, sy, triPdf);
            const V3 p0 = ms.verts[3 * tri], p1 = ms.verts[3 * tri + 1], p2 = ms.verts[3 * tri + 2];
            Float a = std::sqrt(std::max(0.0, 1.0f - sx));                           // warp.cpp:76-79
            Float bx = 1 - a, by = a * sy;
            V3 sideA = p1 - p0, sideB = p2 - p0;
            dRec.p = p0 + (sideA * bx) + (sideB * by);
            if (!ms.normals.empty())                                                 // triangle.cpp:33-42
                dRec.n = normalize(ms.normals[3 * tri] * (1.0f - bx - by) + ms.normals[3 * tri + 1] * bx + ms.normals[3 * tri + 2] * by);
            else dRec.n = normalize(cross(sideA, sideB));
            dRec.pdf = ms.invSurfaceArea;
        }
        dRec.d = dRec.p - dRec.ref;                                                 // shape.cpp:102-114
        Float distSquared = lengthSquared(dRec.d);
        dRec.dist = std::sqrt(distSquared);
        dRec.d = dRec.d / dRec.dist;
        Float dp = std::abs(dot(dRec.d, dRec.n));
        dRec.pdf *= dp != 0 ? (distSquared / dp) : 0.0;
        }
        if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) value = specOf(em.radiance) / dRec.pdf;   // area.cpp:158-176
        else { dRec.pdf = 0.0; value = spec(0); }
    }
    dRec.pdf *= emPdf;
    value = value / emPdf;
    Ray ray = {dRec.ref, dRec.d, Epsilon, dRec.dist * (1 - ShadowEpsilon)};
    if (rayOccluded(sc, ray)) { visible = false; return spec(0); }
    visible = true;
    return value;
}

// Scene::pdfEmitterDirect, scene.cpp:976-979 + area.cpp:178-186 + shape.cpp:116-126 / envmap.cpp:546-556
Float pdfEmitterDirect(const Scene &sc, const DRec &dRec)
{
    const gdb200_emitter &em = sc.ems[dRec.emitter];
    Float pdf = 0.0;
    if (em.type == GDB200_EMITTER_ENVMAP) pdf = envPdfDirection(sc.env, xfVector(sc.env.toObject, dRec.d));
    else if (em.type == GDB200_EMITTER_POINT || em.type == GDB200_EMITTER_SPOT) pdf = 0.0;   // point.cpp:149-151, spot.cpp:202-204: solid-angle query
    else if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0) {
        const Shape &s = sc.shapes[em.shape];
        if (s.d.type == GDB200_SHAPE_SPHERE) {                                      // Sphere::pdfDirect, sphere.cpp:357-387
            const Float m_radius = s.d.radius;
            const V3 refToCenter = specOf(s.d.center) - dRec.ref;
            const Float invRefDist = (Float)1.0f / length(refToCenter), sinAlpha = m_radius * invRefDist;
            if (sinAlpha < 1 - Epsilon) pdf = INV_TWOPI / (1 - std::sqrt(std::max(0.0, 1 - sinAlpha * sinAlpha)));
            else pdf = (1 / (4 * PI * m_radius * m_radius)) * dRec.dist * dRec.dist / std::abs(dot(dRec.d, dRec.n));
        } else {
        const Float pdfPos = s.d.type == GDB200_SHAPE_RECTANGLE ? s.invArea : sc.meshSampling[em.shape].invSurfaceArea;
        pdf = pdfPos * (dRec.dist * dRec.dist) / std::abs(dot(dRec.d, dRec.n));
        }
    }
    return pdf * (em.sampling_weight * sc.emNormalization);
}

// ---------------------------------------------------------------- G-PT
enum VertexType { VERTEX_TYPE_GLOSSY, VERTEX_TYPE_DIFFUSE };
enum RayConnection { RAY_NOT_CONNECTED, RAY_RECENTLY_CONNECTED, RAY_CONNECTED };

struct Config { int maxDepth, minDepth, rrDepth; bool strictNormals; Float shiftThreshold; bool uninitMeasureIsInvalid; };

struct RayState {                      // gpt.cpp:135-173
    Ray ray; Spec throughput; Float pdf; Spec radiance, gradient; Its its; Float eta; bool alive; RayConnection connection_status;
    RayState() : throughput(spec(0)), pdf(1.0), radiance(spec(0)), gradient(spec(0)), eta(1.0), alive(true), connection_status(RAY_NOT_CONNECTED) {}
    void addRadiance(Spec c, Float w) { radiance = radiance + c * w; }
    void addGradient(Spec c, Float w) { gradient = gradient + c * w; }
};

// gpt.cpp:176-226
VertexType getVertexType(const gdb200_material &m, const Config &cfg, unsigned bsdfTypeMask)
{
    Float lowest = INF;
    bool found_smooth = false, found_dirac = false;
    for (int i = 0, n = bsdfComponentCount(m); i < n; ++i) {
        Float r = bsdfRoughness(m, i);
        if (r == 0) { found_dirac = true; if (!(bsdfTypeMask & EDelta)) continue; } else found_smooth = true;
        if (r < lowest) lowest = r;
    }
    if (!found_smooth && found_dirac && !(bsdfTypeMask & EDelta)) lowest = 0;
    return lowest <= cfg.shiftThreshold ? VERTEX_TYPE_GLOSSY : VERTEX_TYPE_DIFFUSE;
}

// util.cpp:763-765, :774-792
inline V3 reflectAbout(V3 wi, V3 n) { return 2 * dot(wi, n) * n - wi; }
inline V3 refractAbout(V3 wi, V3 n, Float eta)
{
    if (eta == 1) return -wi;
    Float cosThetaI = dot(wi, n);
    if (cosThetaI > 0) eta = 1 / eta;
    Float cosThetaTSqr = 1 - (1 - cosThetaI * cosThetaI) * (eta * eta);
    if (cosThetaTSqr <= 0.0) return v3(0, 0, 0);
    return n * (cosThetaI * eta - std::copysign(1.0, cosThetaI) * std::sqrt(cosThetaTSqr)) - wi * eta;
}

struct ShiftResult { bool success; Float jacobian; V3 wo; };

// gpt.cpp:242-305
ShiftResult halfVectorShift(V3 mainWi, V3 mainWo, V3 shiftedWi, Float mainEta, Float shiftedEta)
{
    ShiftResult result; result.success = false; result.jacobian = 0; result.wo = v3(0, 0, 0);
    if (mainWi.z * mainWo.z < 0) {
        if (mainEta == 1 || shiftedEta == 1) return result;
        V3 hMain = (mainWi.z < 0) ? -(mainWi * mainEta + mainWo) : -(mainWi + mainWo * mainEta);
        V3 h = normalize(hMain);
        V3 shiftedWo = refractAbout(shiftedWi, h, shiftedEta);
        if (isZero(shiftedWo)) return result;
        V3 hShifted = (shiftedWi.z < 0) ? -(shiftedWi * shiftedEta + shiftedWo) : -(shiftedWi + shiftedWo * shiftedEta);
        Float hLengthSquared = lengthSquared(hShifted) / (D_EPSILON + lengthSquared(hMain));
        Float WoDotH = std::abs(dot(mainWo, h)) / (D_EPSILON + std::abs(dot(shiftedWo, h)));
        result.success = true; result.wo = shiftedWo; result.jacobian = hLengthSquared * WoDotH;
    } else {
        V3 h = normalize(mainWi + mainWo);
        V3 shiftedWo = reflectAbout(shiftedWi, h);
        Float WoDotH = dot(shiftedWo, h) / dot(mainWo, h);
        result.success = true; result.wo = shiftedWo; result.jacobian = std::abs(WoDotH);
    }
    return result;
}

// gpt.cpp:84-93
bool testVisibility(const Scene &sc, V3 p1, V3 p2)
{
    Ray r = {p1, p2 - p1, Epsilon, 1.0 - ShadowEpsilon};
    return !rayOccluded(sc, r);
}
// gpt.cpp:316-345
ShiftResult reconnectShift(const Scene &sc, V3 mainSource, V3 target, V3 shiftSource, V3 targetNormal)
{
    ShiftResult result; result.success = false; result.jacobian = 0; result.wo = v3(0, 0, 0);
    if (!testVisibility(sc, shiftSource, target)) return result;
    V3 mainEdge = mainSource - target, shiftedEdge = shiftSource - target;
    Float mainL2 = lengthSquared(mainEdge), shiftedL2 = lengthSquared(shiftedEdge);
    V3 shiftedWo = -shiftedEdge / std::sqrt(shiftedL2);
    Float mainOpposingCosine = dot(mainEdge, targetNormal) / std::sqrt(mainL2);
    Float shiftedOpposingCosine = dot(shiftedWo, targetNormal);
    result.jacobian = std::abs(shiftedOpposingCosine * mainL2) / (D_EPSILON + std::abs(mainOpposingCosine * shiftedL2));
    result.success = true; result.wo = shiftedWo;
    return result;
}

// gpt.cpp:96-114
bool testEnvironmentVisibility(const Scene &sc, const Ray &ray)
{
    if (!sc.env.present) return false;
    DRec dr; dr.dist = 0;
    envFillDirectSamplingRecord(sc, dr, ray.o, ray.d);
    Ray shadowRay = {ray.o, ray.d, Epsilon, ((Float)1.0 - ShadowEpsilon) * dr.dist};
    return !rayOccluded(sc, shadowRay);
}
// gpt.cpp:348-369
ShiftResult environmentShift(const Scene &sc, const Ray &mainRay, V3 shiftSourceVertex)
{
    ShiftResult result; result.success = false; result.jacobian = 0; result.wo = v3(0, 0, 0);
    Ray offsetRay = mainRay; offsetRay.o = shiftSourceVertex;
    if (!testEnvironmentVisibility(sc, offsetRay)) return result;
    result.success = true; result.jacobian = 1; result.wo = mainRay.d;
    return result;
}

struct Counters { double rays, vertices; };

// PerspectiveCamera::sampleRayDifferential, perspective.cpp:271-298 / ThinLensCamera, thinlens.cpp:289-318 (aperture_radius > 0);
// ray differentials are unused by the texture-free material subset
void sampleCameraRay(const Scene &sc, Float px, Float py, Float ax, Float ay, Ray &ray)
{
    V3 nearP = xfPoint(sc.cam.sample_to_camera, v3(px * sc.invResX, py * sc.invResY, 0.0));
    V3 d, origin = v3(0, 0, 0);
    if (sc.cam.aperture_radius > 0) {
        Float tx, ty; squareToUniformDiskConcentric(ax, ay, tx, ty);
        V3 apertureP = v3(tx * sc.cam.aperture_radius, ty * sc.cam.aperture_radius, 0.0);
        V3 focusP = nearP * (sc.cam.focus_distance / nearP.z);
        d = normalize(focusP - apertureP);
        origin = apertureP;
    } else d = normalize(nearP);
    Float invZ = 1.0 / d.z;
    ray.mint = sc.cam.near_clip * invZ;
    ray.maxt = sc.cam.far_clip * invZ;
    ray.o = xfAffine(sc.cam.camera_to_world, origin);
    ray.d = xfVector(sc.cam.camera_to_world, d);
}

// gpt.cpp:468-1180.  No environment emitter, no subsurface in this subset: those branches
// (gpt.cpp:486-488,501-504,786-803,908-915,1053-1074) reduce to "path leaves the scene".
void evaluate(const Scene &sc, const Config &cfg, Sampler &sampler, RayState &main, RayState *shiftedRays,
              int secondaryCount, Spec &out_veryDirect, Counters &cnt)
{
    rayIntersect(sc, main.ray, main.its); cnt.rays++;                               // :472
    main.ray.mint = Epsilon;
    for (int i = 0; i < secondaryCount; ++i) {
        rayIntersect(sc, shiftedRays[i].ray, shiftedRays[i].its); cnt.rays++;       // :476-480
        shiftedRays[i].ray.mint = Epsilon;
    }
    if (!main.its.valid()) {                                                        // :482-492
        if (sc.env.present) out_veryDirect = out_veryDirect + main.throughput * evalEnvironment(sc, main.ray.d);
        return;
    }
    if (isEmitter(sc, main.its)) out_veryDirect = out_veryDirect + main.throughput * emittedLe(sc, main.its, -main.ray.d);  // :497-499
    for (int i = 0; i < secondaryCount; ++i) if (!shiftedRays[i].its.valid()) shiftedRays[i].alive = false;              // :508-513
    if (cfg.strictNormals) {                                                        // :516-531
        if (dot(main.ray.d, main.its.geoN) * main.its.wi.z >= 0) return;
        for (int i = 0; i < secondaryCount; ++i) {
            RayState &s = shiftedRays[i];
            if (dot(s.ray.d, s.its.geoN) * s.its.wi.z >= 0) s.alive = false;
        }
    }

    int depth = 1;                                                                  // :535
    while (depth < cfg.maxDepth || cfg.maxDepth < 0) {                              // :537
        if (cfg.strictNormals) {                                                    // :541-555
            if (dot(main.ray.d, main.its.geoN) * main.its.wi.z >= 0) break;
            for (int i = 0; i < secondaryCount; ++i) {
                RayState &s = shiftedRays[i];
                if (dot(s.ray.d, s.its.geoN) * s.its.wi.z >= 0) s.alive = false;
            }
        }
        const bool lastSegment = (depth + 1 == cfg.maxDepth);                       // :558
        const gdb200_material &mainBSDF = matOf(sc, main.its);

        // ---- direct illumination sampling, :565-730
        if ((bsdfType(mainBSDF) & ESmooth) && depth + 1 >= cfg.minDepth) {          // :568
            DRec dRec; initDRec(sc, main.its, dRec);
            Float lsx, lsy; sampler.next2D(lsx, lsy);                               // :572
            bool mainEmitterVisible;
            Spec value = sampleEmitterDirectVisible(sc, dRec, lsx, lsy, mainEmitterVisible); cnt.rays++;
            Spec mainEmitterRadiance = value * dRec.pdf;                            // :575
            V3 mainWoLocal = toLocal(main.its.sh, dRec.d);
            Spec mainBSDFValue = bsdfEval(mainBSDF, main.its.wi, mainWoLocal, ESolidAngle);                  // :588
            const bool onSurfaceSolidAngle = sc.ems[dRec.emitter].type != GDB200_EMITTER_POINT && sc.ems[dRec.emitter].type != GDB200_EMITTER_SPOT && !dRec.discrete;   // emitter->isOnSurface() && dRec.measure == ESolidAngle
            const bool mainAtPointLight = dRec.discrete;                            // :670
            Float mainBsdfPdf = (onSurfaceSolidAngle && mainEmitterVisible) ? bsdfPdf(mainBSDF, main.its.wi, mainWoLocal, ESolidAngle) : 0;  // :592
            Float mainDistanceSquared = lengthSquared(main.its.p - dRec.p);         // :595-596
            Float mainOpposingCosine = dot(dRec.n, (main.its.p - dRec.p)) / std::sqrt(mainDistanceSquared);
            Float mainWeightNumerator = main.pdf * dRec.pdf;                        // :599-600
            Float mainWeightDenominator = (main.pdf * main.pdf) * ((dRec.pdf * dRec.pdf) + (mainBsdfPdf * mainBsdfPdf));
            if (!cfg.strictNormals || dot(main.its.geoN, dRec.d) * mainWoLocal.z > 0) {                      // :607
                for (int i = 0; i < secondaryCount; ++i) {
                    RayState &shifted = shiftedRays[i];
                    Spec mainContribution = spec(0), shiftedContribution = spec(0);
                    Float weight = 0;
                    bool shiftSuccessful = shifted.alive;
                    if (shiftSuccessful) {
                        if (shifted.connection_status == RAY_CONNECTED) {           // :622-637
                            Float shiftedBsdfPdf = mainBsdfPdf, shiftedDRecPdf = dRec.pdf, jacobian = 1;
                            Float den = (jacobian * shifted.pdf) * (jacobian * shifted.pdf) * ((shiftedDRecPdf * shiftedDRecPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                            weight = mainWeightNumerator / (D_EPSILON + den + mainWeightDenominator);
                            mainContribution = main.throughput * (mainBSDFValue * mainEmitterRadiance);
                            shiftedContribution = jacobian * shifted.throughput * (mainBSDFValue * mainEmitterRadiance);
                        } else if (shifted.connection_status == RAY_RECENTLY_CONNECTED) {   // :638-658
                            V3 incoming = normalize(shifted.its.p - main.its.p);
                            V3 wiL = toLocal(main.its.sh, incoming);
                            Float shiftedBsdfPdf = (onSurfaceSolidAngle && mainEmitterVisible) ? bsdfPdf(mainBSDF, wiL, mainWoLocal, ESolidAngle) : 0;
                            Float shiftedDRecPdf = dRec.pdf;
                            Spec shiftedBsdfValue = bsdfEval(mainBSDF, wiL, mainWoLocal, ESolidAngle);
                            Float jacobian = 1;
                            Float den = (jacobian * shifted.pdf) * (jacobian * shifted.pdf) * ((shiftedDRecPdf * shiftedDRecPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                            weight = mainWeightNumerator / (D_EPSILON + den + mainWeightDenominator);
                            mainContribution = main.throughput * (mainBSDFValue * mainEmitterRadiance);
                            shiftedContribution = jacobian * shifted.throughput * (shiftedBsdfValue * mainEmitterRadiance);
                        } else {                                                    // :659-705
                            const gdb200_material &shiftedBSDF = matOf(sc, shifted.its);
                            VertexType mainVT = getVertexType(mainBSDF, cfg, ESmooth), shiftedVT = getVertexType(shiftedBSDF, cfg, ESmooth);
                            if (mainAtPointLight || (mainVT == VERTEX_TYPE_DIFFUSE && shiftedVT == VERTEX_TYPE_DIFFUSE)) {   // :672
                                DRec sRec; initDRec(sc, shifted.its, sRec);
                                bool shiftedEmitterVisible;
                                Spec sv = sampleEmitterDirectVisible(sc, sRec, lsx, lsy, shiftedEmitterVisible); cnt.rays++;
                                Spec shiftedEmitterRadiance = sv * sRec.pdf;
                                Float shiftedDRecPdf = sRec.pdf;
                                Float shiftedDistanceSquared = lengthSquared(dRec.p - shifted.its.p);
                                V3 emitterDirection = (dRec.p - shifted.its.p) / std::sqrt(shiftedDistanceSquared);
                                Float shiftedOpposingCosine = -dot(dRec.n, emitterDirection);
                                V3 woL = toLocal(shifted.its.sh, emitterDirection);
                                if (cfg.strictNormals && dot(shifted.its.geoN, emitterDirection) * woL.z < 0) {
                                    shiftSuccessful = false;
                                } else {
                                    Spec shiftedBsdfValue = bsdfEval(shiftedBSDF, shifted.its.wi, woL, ESolidAngle);
                                    Float shiftedBsdfPdf = (onSurfaceSolidAngle && shiftedEmitterVisible) ? bsdfPdf(shiftedBSDF, shifted.its.wi, woL, ESolidAngle) : 0;
                                    Float jacobian = std::abs(shiftedOpposingCosine * mainDistanceSquared) / (Epsilon + std::abs(mainOpposingCosine * shiftedDistanceSquared));   // :695
                                    Float den = (jacobian * shifted.pdf) * (jacobian * shifted.pdf) * ((shiftedDRecPdf * shiftedDRecPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                                    weight = mainWeightNumerator / (D_EPSILON + den + mainWeightDenominator);
                                    mainContribution = main.throughput * (mainBSDFValue * mainEmitterRadiance);
                                    shiftedContribution = jacobian * shifted.throughput * (shiftedBsdfValue * shiftedEmitterRadiance);
                                }
                            }
                        }
                    }
                    if (!shiftSuccessful) {                                         // :708-717
                        weight = mainWeightNumerator / (D_EPSILON + mainWeightDenominator);
                        mainContribution = main.throughput * (mainBSDFValue * mainEmitterRadiance);
                        shiftedContribution = spec(0);
                    }
                    main.addRadiance(mainContribution, weight);                     // :723-726
                    shifted.addRadiance(shiftedContribution, weight);
                    shifted.addGradient(shiftedContribution - mainContribution, weight);
                }
            }
        }

        // ---- BSDF sampling and emitter hits, :737-826
        BSDFSample bs; bs.wi = main.its.wi;
        { Float sx, sy; sampler.next2D(sx, sy); bsdfSample(mainBSDF, bs, sx, sy, sampler); }  // :456-457 (bRec.sampler = rRec.sampler)
        if (bs.pdf <= 0.0) break;                                                   // :739
        const V3 mainWo = toWorld(main.its.sh, bs.wo);
        Float mainWoDotGeoN = dot(main.its.geoN, mainWo);
        if (cfg.strictNormals && mainWoDotGeoN * bs.wo.z <= 0) break;               // :748
        Its previousMainIts = main.its;                                             // :753
        bool mainHitEmitter = false;
        Spec mainEmitterRadiance = spec(0);
        DRec mainDRec; initDRec(sc, main.its, mainDRec);                            // :759
        VertexType mainVertexType = getVertexType(mainBSDF, cfg, bs.sampledType);   // :764
        VertexType mainNextVertexType;
        main.ray.o = main.its.p; main.ray.d = mainWo; main.ray.mint = Epsilon; main.ray.maxt = INF;   // :767
        cnt.rays++;
        if (rayIntersect(sc, main.ray, main.its)) {                                 // :769-784
            if (isEmitter(sc, main.its)) {
                mainEmitterRadiance = emittedLe(sc, main.its, -main.ray.d);
                mainDRec.p = main.its.p; mainDRec.n = main.its.sh.n; mainDRec.d = main.ray.d; mainDRec.dist = main.its.t;   // setQuery, records.inl:167-175
                mainDRec.emitter = sc.shapes[main.its.shape].d.emitter;
                mainHitEmitter = true;
            }
            mainNextVertexType = getVertexType(matOf(sc, main.its), cfg, bs.sampledType);
        } else {                                                                    // :786-803
            if (sc.env.present) {
                mainEmitterRadiance = evalEnvironment(sc, main.ray.d);
                if (!envFillDirectSamplingRecord(sc, mainDRec, main.ray.o, main.ray.d)) break;
                mainHitEmitter = true;
                mainNextVertexType = VERTEX_TYPE_DIFFUSE;
            } else break;
        }
        Float mainBsdfPdf = bs.pdf, mainPreviousPdf = main.pdf;                     // :807-812
        main.throughput = main.throughput * (bs.weight * bs.pdf);
        main.pdf *= bs.pdf;
        main.eta *= bs.eta;
        const Float mainLumPdf = (mainHitEmitter && depth + 1 >= cfg.minDepth && !(bs.sampledType & EDelta)) ? pdfEmitterDirect(sc, mainDRec) : 0;   // :815-816
        Float mainWeightNumerator = mainPreviousPdf * bs.pdf;                       // :819-820
        Float mainWeightDenominator = (mainPreviousPdf * mainPreviousPdf) * ((mainLumPdf * mainLumPdf) + (mainBsdfPdf * mainBsdfPdf));

        for (int i = 0; i < secondaryCount; ++i) {                                  // :830-1151
            RayState &shifted = shiftedRays[i];
            Spec mainContribution = spec(0), shiftedContribution = spec(0);
            Float weight = 0;
            bool postponedShiftEnd = false;
            if (shifted.alive) {
                Float shiftedPreviousPdf = shifted.pdf;
                if (shifted.connection_status == RAY_CONNECTED) {                   // :844-861
                    Spec shiftedBsdfValue = bs.weight * bs.pdf;
                    shifted.throughput = shifted.throughput * shiftedBsdfValue;
                    shifted.pdf *= mainBsdfPdf;
                    Float den = (shiftedPreviousPdf * shiftedPreviousPdf) * ((mainLumPdf * mainLumPdf) + (mainBsdfPdf * mainBsdfPdf));
                    weight = mainWeightNumerator / (D_EPSILON + den + mainWeightDenominator);
                    mainContribution = main.throughput * mainEmitterRadiance;
                    shiftedContribution = shifted.throughput * mainEmitterRadiance;
                } else if (shifted.connection_status == RAY_RECENTLY_CONNECTED) {   // :862-888
                    V3 incoming = normalize(shifted.its.p - main.ray.o);
                    V3 wiL = toLocal(previousMainIts.sh, incoming), woL = toLocal(previousMainIts.sh, main.ray.d);
                    Measure measure = (bs.sampledType & EDelta) ? EDiscrete : ESolidAngle;
                    Spec shiftedBsdfValue = bsdfEval(mainBSDF, wiL, woL, measure);
                    Float shiftedBsdfPdf = bsdfPdf(mainBSDF, wiL, woL, measure);
                    shifted.throughput = shifted.throughput * shiftedBsdfValue;
                    shifted.pdf *= shiftedBsdfPdf;
                    shifted.connection_status = RAY_CONNECTED;
                    Float den = (shiftedPreviousPdf * shiftedPreviousPdf) * ((mainLumPdf * mainLumPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                    weight = mainWeightNumerator / (D_EPSILON + den + mainWeightDenominator);
                    mainContribution = main.throughput * mainEmitterRadiance;
                    shiftedContribution = shifted.throughput * mainEmitterRadiance;
                } else {                                                            // :889-1126
                    const gdb200_material &shiftedBSDF = matOf(sc, shifted.its);
                    VertexType shiftedVertexType = getVertexType(shiftedBSDF, cfg, bs.sampledType);
                    if (mainVertexType == VERTEX_TYPE_DIFFUSE && mainNextVertexType == VERTEX_TYPE_DIFFUSE && shiftedVertexType == VERTEX_TYPE_DIFFUSE) {
                        if (!lastSegment || mainHitEmitter) {                       // :901
                            ShiftResult sr = main.its.valid() ? reconnectShift(sc, main.ray.o, main.its.p, shifted.its.p, main.its.geoN)   // :907
                                                              : environmentShift(sc, main.ray, shifted.its.p);                               // :908-915
                            cnt.rays++;
                            if (!sr.success) { shifted.alive = false; goto shift_failed; }
                            V3 incomingDirection = -shifted.ray.d, outgoingDirection = sr.wo;
                            V3 wiL = toLocal(shifted.its.sh, incomingDirection), woL = toLocal(shifted.its.sh, outgoingDirection);
                            if (cfg.strictNormals && dot(outgoingDirection, shifted.its.geoN) * woL.z <= 0) { shifted.alive = false; goto shift_failed; }
                            Spec shiftedBsdfValue = bsdfEval(shiftedBSDF, wiL, woL, ESolidAngle);    // :935-936
                            Float shiftedBsdfPdf = bsdfPdf(shiftedBSDF, wiL, woL, ESolidAngle);
                            shifted.throughput = shifted.throughput * (shiftedBsdfValue * sr.jacobian);
                            shifted.pdf *= shiftedBsdfPdf * sr.jacobian;
                            shifted.connection_status = RAY_RECENTLY_CONNECTED;
                            if (mainHitEmitter) {                                   // :944-985
                                Spec shiftedEmitterRadiance = spec(0); Float shiftedLumPdf = 0;
                                if (main.its.valid()) {
                                    shiftedEmitterRadiance = emittedLe(sc, main.its, -outgoingDirection);
                                    DRec sd;                                        // :957-964
                                    sd.p = mainDRec.p; sd.n = mainDRec.n;
                                    sd.dist = length(mainDRec.p - shifted.its.p);
                                    sd.d = (mainDRec.p - shifted.its.p) / sd.dist;
                                    sd.ref = mainDRec.ref; sd.refN = shifted.its.sh.n; sd.emitter = mainDRec.emitter;
                                    shiftedLumPdf = pdfEmitterDirect(sc, sd);
                                    // gpt.cpp:957 default-constructs shiftedDRec and never sets .measure, so Shape::pdfDirect
                                    // (shape.cpp:116-126) compares an indeterminate value with ESolidAngle.  The restatement uses
                                    // the intended ESolidAngle; gdb200_gpt_params.flags & GDB200_GPT_REF_UNINIT_MEASURE instead mimics a build in which
                                    // the stale value is not ESolidAngle (what g++ -O2 produces from the reference sources here:
                                    // an area emitter then reports density 0), for tests/test_ref_gpt.py.
                                    if (cfg.uninitMeasureIsInvalid && sc.ems[sd.emitter].type == GDB200_EMITTER_AREA) shiftedLumPdf = 0;   // also sphere.cpp:370-387
                                } else { shiftedEmitterRadiance = mainEmitterRadiance; shiftedLumPdf = mainLumPdf; }   // :973-977
                                Float den = (shiftedPreviousPdf * shiftedPreviousPdf) * ((shiftedLumPdf * shiftedLumPdf) + (shiftedBsdfPdf * shiftedBsdfPdf));
                                weight = mainWeightNumerator / (D_EPSILON + den + mainWeightDenominator);
                                mainContribution = main.throughput * mainEmitterRadiance;
                                shiftedContribution = shifted.throughput * shiftedEmitterRadiance;
                            }
                        }
                    } else {                                                        // half-vector shift, :987-1126
                        V3 tangentSpaceIncomingDirection = toLocal(shifted.its.sh, -shifted.ray.d);
                        V3 tangentSpaceOutgoingDirection = v3(0, 0, 0);
                        Spec shiftedEmitterRadiance = spec(0);
                        bool bothDelta = (bs.sampledType & EDelta) && (bsdfType(shiftedBSDF) & EDelta);
                        bool bothSmooth = (bs.sampledType & ESmooth) && (bsdfType(shiftedBSDF) & ESmooth);
                        if (!(bothDelta || bothSmooth)) { shifted.alive = false; goto half_vector_shift_failed; }
                        {
                            ShiftResult sr = halfVectorShift(bs.wi, bs.wo, toLocal(shifted.its.sh, -shifted.ray.d), bsdfEta(mainBSDF), bsdfEta(shiftedBSDF));   // :1006
                            if (bs.sampledType & EDelta) sr.jacobian = 1;           // :1008-1011
                            if (sr.success) {
                                shifted.throughput = shifted.throughput * sr.jacobian;
                                shifted.pdf *= sr.jacobian;
                                tangentSpaceOutgoingDirection = sr.wo;
                            } else { shifted.alive = false; goto half_vector_shift_failed; }
                            V3 outgoingDirection = toWorld(shifted.its.sh, tangentSpaceOutgoingDirection);
                            Measure measure = (bs.sampledType & EDelta) ? EDiscrete : ESolidAngle;
                            shifted.throughput = shifted.throughput * bsdfEval(shiftedBSDF, tangentSpaceIncomingDirection, tangentSpaceOutgoingDirection, measure);   // :1030-1031
                            shifted.pdf *= bsdfPdf(shiftedBSDF, tangentSpaceIncomingDirection, tangentSpaceOutgoingDirection, measure);
                            if (shifted.pdf == 0) { shifted.alive = false; goto half_vector_shift_failed; }
                            if (cfg.strictNormals && dot(outgoingDirection, shifted.its.geoN) * tangentSpaceOutgoingDirection.z <= 0) { shifted.alive = false; goto half_vector_shift_failed; }
                            VertexType shiftedVertexType2 = getVertexType(shiftedBSDF, cfg, bs.sampledType);   // :1047
                            shifted.ray.o = shifted.its.p; shifted.ray.d = outgoingDirection; shifted.ray.mint = Epsilon; shifted.ray.maxt = INF;   // :1050
                            cnt.rays++;
                            if (!rayIntersect(sc, shifted.ray, shifted.its)) {      // :1052-1074
                                if (!sc.env.present) { shifted.alive = false; goto half_vector_shift_failed; }
                                if (main.its.valid()) { shifted.alive = false; goto half_vector_shift_failed; }
                                if (mainVertexType == VERTEX_TYPE_DIFFUSE && shiftedVertexType2 == VERTEX_TYPE_DIFFUSE) { shifted.alive = false; goto half_vector_shift_failed; }
                                shiftedEmitterRadiance = evalEnvironment(sc, shifted.ray.d);
                                postponedShiftEnd = true;
                            } else {
                                if (!main.its.valid()) { shifted.alive = false; goto half_vector_shift_failed; }   // :1078-1082
                                VertexType shiftedNextVertexType = getVertexType(matOf(sc, shifted.its), cfg, bs.sampledType);
                                if (mainVertexType == VERTEX_TYPE_DIFFUSE && shiftedVertexType2 == VERTEX_TYPE_DIFFUSE && shiftedNextVertexType == VERTEX_TYPE_DIFFUSE) {   // :1089-1093
                                    shifted.alive = false; goto half_vector_shift_failed;
                                }
                                if (isEmitter(sc, shifted.its)) shiftedEmitterRadiance = emittedLe(sc, shifted.its, -shifted.ray.d);   // :1095-1098
                            }
                        }
half_vector_shift_failed:
                        if (shifted.alive) {                                        // :1107-1112
                            weight = main.pdf / (shifted.pdf * shifted.pdf + main.pdf * main.pdf);
                            mainContribution = main.throughput * mainEmitterRadiance;
                            shiftedContribution = shifted.throughput * shiftedEmitterRadiance;
                        } else {                                                    // :1113-1125
                            weight = (Float)1 / main.pdf;
                            mainContribution = main.throughput * mainEmitterRadiance;
                            shiftedContribution = spec(0);
                            shifted.alive = true;
                            postponedShiftEnd = true;
                        }
                    }
                }
            }
shift_failed:
            if (!shifted.alive) {                                                   // :1131-1136
                weight = mainWeightNumerator / (D_EPSILON + mainWeightDenominator);
                mainContribution = main.throughput * mainEmitterRadiance;
                shiftedContribution = spec(0);
            }
            if (depth + 1 >= cfg.minDepth) {                                        // :1140-1146
                main.addRadiance(mainContribution, weight);
                shifted.addRadiance(shiftedContribution, weight);
                shifted.addGradient(shiftedContribution - mainContribution, weight);
            }
            if (postponedShiftEnd) shifted.alive = false;                           // :1148-1150
        }

        if (!main.its.valid()) break;                                               // :1154-1157: the base path hit the environment
        if (depth++ >= cfg.rrDepth) {                                               // :1159-1174
            Float q = std::min(maxComp(main.throughput / main.pdf) * main.eta * main.eta, (Float)0.95f);
            if (sampler.next1D() >= q) break;
            main.pdf *= q;
            for (int i = 0; i < secondaryCount; ++i) shiftedRays[i].pdf *= q;
        }
    }
    cnt.vertices += depth;                                                          // :1178-1179
}

// ---------------------------------------------------------------- film
struct Film {
    int w, h; Float radius, values[32];    // ReconstructionFilter::m_values (rfilter.cpp:37-55); box: 1/(2r) (box.cpp:45-47)
    std::vector<Float> acc;                // [5][h][w][4]: R,G,B,weight   (the alpha channel of gpt_wr.h:60 is dropped)
    Float *px(int buf, int x, int y) { return &acc[(((size_t)buf * h + y) * w + x) * 4]; }
    // rfilter.h:76-77 with MTS_FILTER_RESOLUTION = 31
    Float evalDiscretized(Float x) const
    {
        return values[std::min((int)std::abs(x * (31 / radius)), 31)];
    }
    // GPTWorkResult::put (gpt_wr.h:56-64) -> ImageBlock::put (imageblock.h:150-195), in image coordinates
    void put(Float sx, Float sy, Spec v, Float weight, int buf, bool allowNegative)
    {
        const Float value[4] = {v.x, v.y, v.z, weight};
        for (int i = 0; i < 4; i++)
            if (!std::isfinite(value[i]) || (!allowNegative && value[i] < 0)) return;   // sample dropped with its weight
        const Float posx = sx - 0.5, posy = sy - 0.5;
        const int minx = std::max((int)std::ceil(posx - radius), 0), miny = std::max((int)std::ceil(posy - radius), 0);
        const int maxx = std::min((int)std::floor(posx + radius), w - 1), maxy = std::min((int)std::floor(posy + radius), h - 1);
        for (int y = miny; y <= maxy; ++y) {
            const Float weightY = evalDiscretized(y - posy);
            for (int x = minx; x <= maxx; ++x) {
                const Float wgt = evalDiscretized(x - posx) * weightY;
                Float *dst = px(buf, x, y);
                for (int k = 0; k < 4; k++) dst[k] += wgt * value[k];
            }
        }
    }
};

enum { BUF_FINAL = 0, BUF_THROUGHPUT = 1, BUF_DX = 2, BUF_DY = 3, BUF_DIRECT = 4 };

// GaussLobattoIntegrator (quad.cpp:287-403) as fresnelDiffuseReflectance(eta, false) uses it (util.cpp:855-859):
// maxEvals 1024, absError 0, relError 1e-5, useConvergenceEstimate = false.
struct GaussLobatto {
    Float eta; size_t maxEvals, evals;
    Float f(Float xi) const { Float unused; return fresnelDielectricExt(std::sqrt(xi), unused, eta); }   // util.cpp:808-811
    Float adaptiveStep(Float a, Float b, Float fa, Float fb, Float acc)
    {
        const Float m_alpha = (Float)std::sqrt(2.0 / 3.0), m_beta = (Float)(1.0 / std::sqrt(5.0));
        const Float h = (b - a) / 2, m = (a + b) / 2;
        const Float mll = m - m_alpha * h, ml = m - m_beta * h, mr = m + m_beta * h, mrr = m + m_alpha * h;
        const Float fmll = f(mll), fml = f(ml), fm = f(m), fmr = f(mr), fmrr = f(mrr);
        const Float integral2 = (h / 6) * (fa + fb + 5 * (fml + fmr));
        const Float integral1 = (h / 1470) * (77 * (fa + fb) + 432 * (fmll + fmrr) + 625 * (fml + fmr) + 672 * fm);
        evals += 5;
        if (evals >= maxEvals) return integral1;
        Float dist = acc + (integral1 - integral2);
        if (dist == acc || mll <= a || b <= mrr) return integral1;
        return adaptiveStep(a, mll, fa, fmll, acc) + adaptiveStep(mll, ml, fmll, fml, acc) + adaptiveStep(ml, m, fml, fm, acc)
             + adaptiveStep(m, mr, fm, fmr, acc) + adaptiveStep(mr, mrr, fmr, fmrr, acc) + adaptiveStep(mrr, b, fmrr, fb, acc);
    }
    Float integrate(Float a, Float b)
    {
        const Float m_alpha = (Float)std::sqrt(2.0 / 3.0), m_beta = (Float)(1.0 / std::sqrt(5.0));
        const Float m_x1 = (Float)0.94288241569547971906, m_x2 = (Float)0.64185334234578130578, m_x3 = (Float)0.23638319966214988028;
        evals = 0;
        const Float m = (a + b) / 2, h = (b - a) / 2;
        const Float y1 = f(a), y3 = f(m - m_alpha * h), y5 = f(m - m_beta * h), y7 = f(m), y9 = f(m + m_beta * h), y11 = f(m + m_alpha * h), y13 = f(b);
        Float acc = h * ((Float)0.0158271919734801831 * (y1 + y13) + (Float)0.0942738402188500455 * (f(m - m_x1 * h) + f(m + m_x1 * h))
                         + (Float)0.1550719873365853963 * (y3 + y11) + (Float)0.1888215739601824544 * (f(m - m_x2 * h) + f(m + m_x2 * h))
                         + (Float)0.1997734052268585268 * (y5 + y9) + (Float)0.2249264653333395270 * (f(m - m_x3 * h) + f(m + m_x3 * h))
                         + (Float)0.2426110719014077338 * y7);
        evals += 13;
        const Float r = 1.0, relError = 1e-5f;
        Float absTolerance = std::numeric_limits<Float>::infinity();
        if (acc != 0) absTolerance = acc * std::max(relError, std::numeric_limits<Float>::epsilon()) / (r * std::numeric_limits<Float>::epsilon());
        evals += 2;
        return adaptiveStep(a, b, f(a), f(b), absTolerance);
    }
};
inline Float fresnelDiffuseReflectance(Float eta) { GaussLobatto q; q.eta = eta; q.maxEvals = 1024; return q.integrate(0, 1); }

void buildScene(const gdb200_scene_desc *d, Scene &sc)
{
    sc.cam = d->camera;
    sc.invResX = 1.0 / d->camera.width; sc.invResY = 1.0 / d->camera.height;   // sensor.cpp: m_invResolution
    sc.filterRadius = d->rfilter_radius;
    sc.mats.assign(d->materials, d->materials + d->n_materials);
    sc.ems.assign(d->emitters, d->emitters + d->n_emitters);
    for (int i = 0; i < d->n_shapes; i++) {
        Shape s; s.d = d->shapes[i];
        if (s.d.type == GDB200_SHAPE_RECTANGLE) {                               // rectangle.cpp:100-110
            s.dpdu = xfVector(s.d.to_world, v3(2, 0, 0));
            s.dpdv = xfVector(s.d.to_world, v3(0, 2, 0));
            V3 normal = normalize(xfNormal(s.d.to_object, v3(0, 0, 1)));
            s.frame.s = normalize(s.dpdu); s.frame.t = normalize(s.dpdv); s.frame.n = normal;
            s.invArea = 1.0 / (length(s.dpdu) * length(s.dpdv));
        } else if (s.d.type == GDB200_SHAPE_MESH) {
            for (int t = s.d.first_tri; t < s.d.first_tri + s.d.tri_count; t++) {
                const int *ix = d->triangles + 3 * t;
                Tri T; T.shape = i;
                triLoad(T, specOf(d->vertices + 3 * ix[0]), specOf(d->vertices + 3 * ix[1]), specOf(d->vertices + 3 * ix[2]));
                T.hasNormals = s.d.has_vertex_normals && d->normals;
                if (T.hasNormals) { T.n0 = specOf(d->normals + 3 * ix[0]); T.n1 = specOf(d->normals + 3 * ix[1]); T.n2 = specOf(d->normals + 3 * ix[2]); }
                sc.tris.push_back(T);
            }
        }
        sc.shapes.push_back(s);
    }
    // TriMesh::prepareSamplingTable (trimesh.cpp:388-403) for emitting meshes
    sc.meshSampling.resize(d->n_shapes);
    for (int i = 0; i < d->n_shapes; i++) {
        const gdb200_shape &sh = d->shapes[i];
        if (sh.type != GDB200_SHAPE_MESH || sh.emitter < 0) continue;
        MeshSampling &ms = sc.meshSampling[i];
        ms.cdf.assign(1, 0.0);
        for (int t = sh.first_tri; t < sh.first_tri + sh.tri_count; t++) {
            const int *ix = d->triangles + 3 * t;
            V3 p0 = specOf(d->vertices + 3 * ix[0]), p1 = specOf(d->vertices + 3 * ix[1]), p2 = specOf(d->vertices + 3 * ix[2]);
            ms.verts.push_back(p0); ms.verts.push_back(p1); ms.verts.push_back(p2);
            if (sh.has_vertex_normals && d->normals) for (int k = 0; k < 3; k++) ms.normals.push_back(specOf(d->normals + 3 * ix[k]));
            ms.cdf.push_back(ms.cdf.back() + 0.5f * length(cross(p1 - p0, p2 - p0)));   // triangle.cpp:61-67, pmf.h:62-71
        }
        Float surfaceArea = ms.cdf.back();                                       // DiscreteDistribution::normalize, pmf.h:101-114
        if (surfaceArea > 0) {
            Float normalization = 1.0f / surfaceArea;
            for (size_t k = 1; k < ms.cdf.size(); ++k) ms.cdf[k] *= normalization;
            ms.cdf.back() = 1.0f;
        }
        ms.invSurfaceArea = 1.0f / surfaceArea;
    }
    // plastic.cpp:188-190
    sc.fdrInt.assign(d->n_materials, 0.0); sc.fdrExt.assign(d->n_materials, 0.0);
    for (int i = 0; i < d->n_materials; i++)
        if (d->materials[i].type == GDB200_BSDF_PLASTIC) {
            sc.fdrInt[i] = fresnelDiffuseReflectance(1 / d->materials[i].ior_ratio);
            sc.fdrExt[i] = fresnelDiffuseReflectance(d->materials[i].ior_ratio);
        }
    // EnvironmentMap::configure, envmap.cpp:263-320
    sc.env = EnvMap();
    for (int i = 0; i < d->n_emitters; i++) {
        if (d->emitters[i].type != GDB200_EMITTER_ENVMAP || !d->envmap) continue;
        const gdb200_envmap &src = *d->envmap;
        EnvMap &e = sc.env;
        e.present = true; e.emitter = i; e.w = src.width; e.h = src.height; e.scale = src.scale;
        memcpy(e.toWorld, src.to_world, sizeof(e.toWorld)); memcpy(e.toObject, src.to_object, sizeof(e.toObject));
        e.center = specOf(src.bsphere_center); e.radius = src.bsphere_radius;
        e.texels.resize((size_t)e.w * e.h * 3);
        for (size_t k = 0; k < e.texels.size(); k++) e.texels[k] = (Float)src.rgb[k];
        size_t nEntries = (size_t)(e.w + 1) * (size_t)e.h;
        e.cdfCols.assign(nEntries, 0.f); e.cdfRows.assign(e.h + 1, 0.f); e.rowWeights.assign(e.h, 0.0);
        size_t colPos = 0, rowPos = 0;
        Float rowSum = 0.0f;
        e.cdfRows[rowPos++] = 0;
        for (int y = 0; y < e.h; ++y) {
            Float colSum = 0;
            e.cdfCols[colPos++] = 0;
            for (int x = 0; x < e.w; ++x) {
                colSum += luminance(envTexel(e, x, y));
                e.cdfCols[colPos++] = (float)colSum;
            }
            float normalization = 1.0f / (float)colSum;
            for (int x = 1; x < e.w; ++x) e.cdfCols[colPos - x - 1] *= normalization;
            e.cdfCols[colPos - 1] = 1.0f;
            Float weight = std::sin((y + 0.5f) * PI / e.h);
            e.rowWeights[y] = weight;
            rowSum += colSum * weight;
            e.cdfRows[rowPos++] = (float)rowSum;
        }
        float normalization = 1.0f / (float)rowSum;
        for (int y = 1; y < e.h; ++y) e.cdfRows[rowPos - y - 1] *= normalization;
        e.cdfRows[rowPos - 1] = 1.0f;
        e.normalization = 1.0f / (rowSum * (2 * PI / e.w) * (PI / e.h));
        e.pixelSizeX = 2 * PI / e.w; e.pixelSizeY = PI / e.h;
    }
    // scene.cpp:357-380 + pmf.h:100-114: CDF over samplingWeight, normalised
    sc.emCdf.assign(1, 0.0);
    for (size_t i = 0; i < sc.ems.size(); i++) sc.emCdf.push_back(sc.emCdf.back() + sc.ems[i].sampling_weight);
    Float sum = sc.emCdf.back();
    sc.emNormalization = sum > 0 ? 1.0 / sum : 0.0;
    if (sum > 0) { for (size_t i = 1; i < sc.emCdf.size(); i++) sc.emCdf[i] *= sc.emNormalization; sc.emCdf.back() = 1.0; }
}

}  // namespace

extern "C" {

// renderBlock (gpt.cpp:1220-1355) over the whole image + developMulti (multifilm.cpp:366-416).
// out buffers: width*height*3 doubles each (any may be NULL); out_weights (optional): 5*h*w weights.
// counters (optional): [0] samples, [1] rays, [2] sum of base path depths.
int gdb200_oracle_gpt_render(const gdb200_scene_desc *desc, const gdb200_gpt_params *prm, gdb200_buffers *out,
                             double *out_weights, double *counters, int num_threads)
{
    if (!desc || !prm || desc->n_emitters < 1) return 1;
    Scene sc; buildScene(desc, sc);
    g_fdrInt = &sc.fdrInt; g_matBase = &sc.mats[0];
    Config cfg; cfg.maxDepth = prm->max_depth; cfg.minDepth = 1; cfg.rrDepth = prm->rr_depth;   // gpt.cpp:1368-1371
    cfg.strictNormals = prm->strict_normals != 0; cfg.shiftThreshold = prm->shift_threshold;
    cfg.uninitMeasureIsInvalid = (prm->flags & GDB200_GPT_REF_UNINIT_MEASURE) != 0;
    const int W = sc.cam.width, H = sc.cam.height;
    Film film; film.w = W; film.h = H; film.radius = sc.filterRadius;
    bool box = true;
    for (int i = 0; i < 32; i++) { film.values[i] = desc->rfilter_table[i]; if (film.values[i] != 0) box = false; }
    if (box) { for (int i = 0; i < 31; i++) film.values[i] = 1.0 / (2 * film.radius); film.values[31] = 0; }
    film.acc.assign((size_t)5 * W * H * 4, 0.0);
    const int y0 = (prm->y_begin == 0 && prm->y_end == 0) ? 0 : prm->y_begin, y1 = (prm->y_begin == 0 && prm->y_end == 0) ? H : prm->y_end;
    double totRays = 0, totVerts = 0;
    const int nChunks = std::max(1, prm->streams_per_pixel);
    // Row bands of 2*reach rows (4 for the box filter): a sample in row y writes rows y-reach..y+reach at most, so bands of equal parity
    // never touch the same film rows and can run concurrently without atomics.
    const int reach = 1 + (int)std::floor(film.radius + 0.5);        // rows a sample can touch above/below its pixel row
    const int band = 2 * reach, nBands = (y1 - y0 + band - 1) / band;
    for (int parity = 0; parity < 2; parity++) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(num_threads > 0 ? num_threads : 1) reduction(+ : totRays, totVerts)
        for (int b = parity; b < nBands; b += 2) {
            Counters cnt = {0, 0};
            Sampler sampler;
            for (int y = y0 + b * band; y < std::min(y1, y0 + (b + 1) * band); y++)
                for (int x = 0; x < W; x++) {
                  for (int chunk = 0; chunk < nChunks; chunk++) {                     // gdb200_gpt_params.streams_per_pixel (1 = the reference's single stream)
                    sampler.generate(prm->seed, x, y);                              // gpt.cpp:1250-1251
                    if (chunk > 0) sampler.key = Sampler::mix(sampler.key ^ ((uint64_t)chunk * 0xD1B54A32D192ED03ULL));
                    const int chunkSpp = prm->spp / nChunks + (chunk < prm->spp % nChunks ? 1 : 0);
                    for (int j = 0; j < chunkSpp; j++) {
                        Float u, v; sampler.next2D(u, v);                           // :1261
                        const Float spx = x + u, spy = y + v;
                        Float apx = 0.5f, apy = 0.5f;                               // :1235
                        if (sc.cam.aperture_radius > 0) sampler.next2D(apx, apy);   // needsApertureSample, :1263-1265
                        RayState main, shifted[4];
                        static const Float shiftX[4] = {1, 0, -1, 0}, shiftY[4] = {0, 1, 0, -1};   // :410-415
                        sampleCameraRay(sc, spx, spy, apx, apy, main.ray); main.throughput = spec(1);
                        for (int i = 0; i < 4; i++) { sampleCameraRay(sc, spx + shiftX[i], spy + shiftY[i], apx, apy, shifted[i].ray); shifted[i].throughput = spec(1); }
                        Spec veryDirect = spec(0);
                        evaluate(sc, cfg, sampler, main, shifted, 4, veryDirect, cnt);
                        const int RIGHT = 0, BOTTOM = 1, LEFT = 2, TOP = 3;         // :1283-1286
                        const Spec C = main.radiance;
                        // :1319-1324 preview/final
                        film.put(spx, spy, (8 * veryDirect) + (2 * C), 4.0, BUF_FINAL, false);
                        film.put(spx - 1, spy, 2 * shifted[LEFT].radiance, 1.0, BUF_FINAL, false);
                        film.put(spx + 1, spy, 2 * shifted[RIGHT].radiance, 1.0, BUF_FINAL, false);
                        film.put(spx, spy - 1, 2 * shifted[TOP].radiance, 1.0, BUF_FINAL, false);
                        film.put(spx, spy + 1, 2 * shifted[BOTTOM].radiance, 1.0, BUF_FINAL, false);
                        // :1334-1339 throughput
                        film.put(spx, spy, 2 * C, 4.0, BUF_THROUGHPUT, false);
                        film.put(spx - 1, spy, 2 * shifted[LEFT].radiance, 1.0, BUF_THROUGHPUT, false);
                        film.put(spx + 1, spy, 2 * shifted[RIGHT].radiance, 1.0, BUF_THROUGHPUT, false);
                        film.put(spx, spy - 1, 2 * shifted[TOP].radiance, 1.0, BUF_THROUGHPUT, false);
                        film.put(spx, spy + 1, 2 * shifted[BOTTOM].radiance, 1.0, BUF_THROUGHPUT, false);
                        // :1345-1348 gradients
                        film.put(spx - 1, spy, -(2 * shifted[LEFT].gradient), 1.0, BUF_DX, true);
                        film.put(spx, spy, 2 * shifted[RIGHT].gradient, 1.0, BUF_DX, true);
                        film.put(spx, spy - 1, -(2 * shifted[TOP].gradient), 1.0, BUF_DY, true);
                        film.put(spx, spy, 2 * shifted[BOTTOM].gradient, 1.0, BUF_DY, true);
                        // :1352 very direct
                        film.put(spx, spy, veryDirect, 1.0, BUF_DIRECT, false);
                    }
                  }
                }
            totRays += cnt.rays; totVerts += cnt.vertices;
        }
    }
    // develop: value * (1/weight), 0 where weight == 0 (fmtconv.cpp:1036-1045)
    double *dst[5] = {out ? out->preview_final : 0, out ? out->throughput : 0, out ? out->dx : 0, out ? out->dy : 0, out ? out->direct : 0};
    for (int buf = 0; buf < 5; buf++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                const Float *p = film.px(buf, x, y);
                const Float wgt = p[3], inv = (wgt != 0) ? 1 / wgt : wgt;
                if (dst[buf]) for (int c = 0; c < 3; c++) dst[buf][((size_t)y * W + x) * 3 + c] = p[c] * inv;
                if (out_weights) out_weights[((size_t)buf * H + y) * W + x] = wgt;
            }
    if (counters) { counters[0] = (double)W * (y1 - y0) * prm->spp; counters[1] = totRays; counters[2] = totVerts; }
    return 0;
}

// Plain MIS path tracer = GradientPathIntegrator::Li (gpt.cpp:1489-1662, a copy of
// src/integrators/path/path.cpp) driven by the same sampler; used only to cross-check the
// G-PT restatement (E[throughput + direct] == E[Li]).  out: width*height*3 doubles.
int gdb200_oracle_path_render(const gdb200_scene_desc *desc, const gdb200_gpt_params *prm, double *out, int num_threads)
{
    if (!desc || !prm || !out) return 1;
    Scene sc; buildScene(desc, sc);
    g_fdrInt = &sc.fdrInt; g_matBase = &sc.mats[0];
    const int W = sc.cam.width, H = sc.cam.height;
#pragma omp parallel for schedule(dynamic, 4) num_threads(num_threads > 0 ? num_threads : 1)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            Sampler sampler; sampler.generate(prm->seed ^ 0x5bd1e995u, x, y);
            Spec sum = spec(0);
            for (int j = 0; j < prm->spp; j++) {
                Float u, v; sampler.next2D(u, v);
                Float apx = 0.5f, apy = 0.5f;
                if (sc.cam.aperture_radius > 0) sampler.next2D(apx, apy);
                Ray ray; sampleCameraRay(sc, x + u, y + v, apx, apy, ray);
                Its its; rayIntersect(sc, ray, its);
                ray.mint = Epsilon;
                Spec Li = spec(0), throughput = spec(1);
                Float eta = 1.0; bool scattered = false; int depth = 1;
                while (depth <= prm->max_depth || prm->max_depth < 0) {
                    if (!its.valid()) {                                                                         // :1515-1521
                        if (!scattered && sc.env.present) Li = Li + throughput * evalEnvironment(sc, ray.d);
                        break;
                    }
                    const gdb200_material &bsdf = matOf(sc, its);
                    if (isEmitter(sc, its) && !scattered) Li = Li + throughput * emittedLe(sc, its, -ray.d);   // :1527-1529 (EEmittedRadiance only on the first vertex)
                    if (depth >= prm->max_depth && prm->max_depth > 0) break;
                    DRec dRec; initDRec(sc, its, dRec);
                    if (bsdfType(bsdf) & ESmooth) {                                                             // :1553-1585
                        Float sx, sy; sampler.next2D(sx, sy);
                        bool vis; Spec value = sampleEmitterDirectVisible(sc, dRec, sx, sy, vis);
                        if (vis && !(value.x == 0 && value.y == 0 && value.z == 0)) {
                            V3 woL = toLocal(its.sh, dRec.d);
                            Spec bsdfVal = bsdfEval(bsdf, its.wi, woL, ESolidAngle);
                            if (!(bsdfVal.x == 0 && bsdfVal.y == 0 && bsdfVal.z == 0)) {
                                Float bsdfPdfV = dRec.discrete ? 0.0 : bsdfPdf(bsdf, its.wi, woL, ESolidAngle);   // emitter->isOnSurface() && measure == ESolidAngle, :1571-1572
                                Float a = dRec.pdf * dRec.pdf, b = bsdfPdfV * bsdfPdfV;
                                Li = Li + throughput * value * bsdfVal * (a / (a + b));
                            }
                        }
                    }
                    BSDFSample bs; bs.wi = its.wi;
                    { Float sx, sy; sampler.next2D(sx, sy); bsdfSample(bsdf, bs, sx, sy, sampler); }
                    if (bs.pdf <= 0 || (bs.weight.x == 0 && bs.weight.y == 0 && bs.weight.z == 0)) break;
                    scattered = true;
                    const V3 wo = toWorld(its.sh, bs.wo);
                    Ray next = {its.p, wo, Epsilon, INF};
                    Its prev = its;
                    bool hit = rayIntersect(sc, next, its);
                    throughput = throughput * bs.weight;
                    eta *= bs.eta;
                    if (!hit) {                                                                                 // :1618-1630
                        if (sc.env.present) {
                            Spec value = evalEnvironment(sc, next.d);
                            DRec q; initDRec(sc, prev, q);
                            if (envFillDirectSamplingRecord(sc, q, next.o, next.d)) {
                                const Float lumPdf = !(bs.sampledType & EDelta) ? pdfEmitterDirect(sc, q) : 0;
                                Float a = bs.pdf * bs.pdf, b = lumPdf * lumPdf;
                                Li = Li + throughput * value * (a / (a + b));
                            }
                        }
                        break;
                    }
                    if (isEmitter(sc, its)) {                                                                   // :1611-1645
                        Spec value = emittedLe(sc, its, -next.d);
                        DRec q; initDRec(sc, prev, q);
                        q.p = its.p; q.n = its.sh.n; q.d = next.d; q.dist = its.t; q.emitter = sc.shapes[its.shape].d.emitter;
                        const Float lumPdf = !(bs.sampledType & EDelta) ? pdfEmitterDirect(sc, q) : 0;
                        Float a = bs.pdf * bs.pdf, b = lumPdf * lumPdf;
                        Li = Li + throughput * value * (a / (a + b));
                    }
                    ray = next;
                    if (depth++ >= prm->rr_depth) {                                                             // :1649-1660
                        Float q = std::min(maxComp(throughput) * eta * eta, (Float)0.95f);
                        if (sampler.next1D() >= q) break;
                        throughput = throughput / q;
                    }
                }
                sum = sum + Li;
            }
            for (int c = 0; c < 3; c++) out[((size_t)y * W + x) * 3 + c] = (&sum.x)[c] / prm->spp;
        }
    return 0;
}


// ---- batched single-plugin entry points for the chi-square tests (tests/test_chisquare.py), which restate the
// reference's own consistency test of these plugins (src/tests/test_chisquare.cpp on data/tests/test_bsdf.xml /
// test_emitter.xml): BSDF::sample / eval / pdf and Emitter::sampleDirect / pdfDirect.
int gdb200_oracle_bsdf_sample_batch(const gdb200_material *m, const double *wi, int n, const double *samples,
                                    double *wo, double *weight, double *pdf, int *sampledType)
{
    std::vector<Float> fdr(1, m->type == GDB200_BSDF_PLASTIC ? fresnelDiffuseReflectance(1 / m->ior_ratio) : 0.0);
    g_fdrInt = &fdr; g_matBase = m;
    for (int i = 0; i < n; i++) {
        BSDFSample bs; bs.wi = v3(wi[0], wi[1], wi[2]);
        Sampler fake; fake.key = 0; fake.n = 0; fake.forced = samples[3 * i + 2]; fake.useForced = true;   // FakeSampler of test_chisquare.cpp
        bsdfSample(*m, bs, samples[3 * i], samples[3 * i + 1], fake);
        wo[3 * i] = bs.wo.x; wo[3 * i + 1] = bs.wo.y; wo[3 * i + 2] = bs.wo.z;
        weight[3 * i] = bs.weight.x; weight[3 * i + 1] = bs.weight.y; weight[3 * i + 2] = bs.weight.z;
        pdf[i] = bs.pdf; sampledType[i] = (int)bs.sampledType;
    }
    return 0;
}
int gdb200_oracle_bsdf_eval_batch(const gdb200_material *m, const double *wi, int n, const double *wo, int measure,
                                  double *value, double *pdf)
{
    std::vector<Float> fdr(1, m->type == GDB200_BSDF_PLASTIC ? fresnelDiffuseReflectance(1 / m->ior_ratio) : 0.0);
    g_fdrInt = &fdr; g_matBase = m;
    const V3 w = v3(wi[0], wi[1], wi[2]);
    for (int i = 0; i < n; i++) {
        const V3 o = v3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]);
        const Spec f = bsdfEval(*m, w, o, measure ? EDiscrete : ESolidAngle);
        value[3 * i] = f.x; value[3 * i + 1] = f.y; value[3 * i + 2] = f.z;
        pdf[i] = bsdfPdf(*m, w, o, measure ? EDiscrete : ESolidAngle);
    }
    return 0;
}
// Scene::rayIntersect for n rays (origin, direction, mint = Epsilon, maxt = inf): hit distance (inf = miss), shape index,
// geometric normal, shading frame (s, t, n) -- for the comparison with the reference's kd-tree (tests/test_ref_gpt.py).
int gdb200_oracle_intersect_batch(const gdb200_scene_desc *desc, int n, const double *o, const double *d, double *t, int *shape,
                                  double *geoN, double *shFrame)
{
    Scene sc; buildScene(desc, sc);
    for (int i = 0; i < n; i++) {
        Ray ray = {v3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), v3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), Epsilon, INF};
        Its its;
        const bool hit = rayIntersect(sc, ray, its);
        t[i] = hit ? its.t : INF; shape[i] = hit ? its.shape : -1;
        const V3 g = hit ? its.geoN : v3(0, 0, 0), fs = hit ? its.sh.s : g, ft = hit ? its.sh.t : g, fn = hit ? its.sh.n : g;
        geoN[3 * i] = g.x; geoN[3 * i + 1] = g.y; geoN[3 * i + 2] = g.z;
        const V3 f[3] = {fs, ft, fn};
        for (int k = 0; k < 3; k++) { shFrame[9 * i + 3 * k] = f[k].x; shFrame[9 * i + 3 * k + 1] = f[k].y; shFrame[9 * i + 3 * k + 2] = f[k].z; }
    }
    return 0;
}
// EnvironmentMap::sampleDirect from the reference point `ref` (EmitterAdapter of test_chisquare.cpp:341-389): world
// direction and solid-angle density per sample; gdb200_oracle_envmap_pdf_batch is the matching pdfDirect.
int gdb200_oracle_envmap_sample_batch(const gdb200_scene_desc *desc, const double *ref, int n, const double *samples, double *d, double *pdf)
{
    Scene sc; buildScene(desc, sc);
    if (!sc.env.present) return 1;
    const gdb200_emitter env = sc.ems[sc.env.emitter];
    sc.ems.assign(1, env); sc.emCdf.assign(2, 0.0); sc.emCdf[1] = 1.0; sc.emNormalization = 1.0 / env.sampling_weight; sc.env.emitter = 0;
    for (int i = 0; i < n; i++) {
        DRec r; r.ref = v3(ref[0], ref[1], ref[2]); r.refN = v3(0, 0, 0);
        bool vis;
        sampleEmitterDirectVisible(sc, r, samples[2 * i], samples[2 * i + 1], vis);   // d and pdf do not depend on the visibility
        d[3 * i] = r.d.x; d[3 * i + 1] = r.d.y; d[3 * i + 2] = r.d.z; pdf[i] = r.pdf;
    }
    return 0;
}
int gdb200_oracle_envmap_pdf_batch(const gdb200_scene_desc *desc, int n, const double *d, double *pdf)
{
    Scene sc; buildScene(desc, sc);
    if (!sc.env.present) return 1;
    for (int i = 0; i < n; i++) pdf[i] = envPdfDirection(sc.env, xfVector(sc.env.toObject, v3(d[3 * i], d[3 * i + 1], d[3 * i + 2])));
    return 0;
}

}  // extern "C"
