// gdb200 G-PT tracer — device-side scene model, intersection and BSDF evaluation (fp64).
//
// What these functions compute is fixed by the reference (file:line cited per function);
// how they are organised is not: the scene is a flat constant-memory table, shapes are tested
// by an unrolled uniform loop (every lane reads the same primitive => constant-cache
// broadcast), materials are plain structs switched on a small enum, and per-material facts
// that the reference re-derives through virtual calls on every use (BSDF type flags, the
// glossy/diffuse vertex classification of gpt.cpp:176-226) are precomputed at scene upload.
#pragma once
#ifndef GDB200_EMU          // tests/emu/ compiles this source for the host through a shim (test infrastructure)
#include <cuda_runtime.h>
#include <math_constants.h>
#endif
#include <cstdint>
#include "../../include/gdb200.h"

namespace gdb200 {

typedef double Float;

#define GDB_HD __host__ __device__ __forceinline__
#define GDB_D __device__ __forceinline__
#define GDB_CALL __device__ __noinline__     // big shared routines: keep one copy so the kernels fit the instruction cache

constexpr Float kEpsilon = 1e-7, kShadowEpsilon = 1e-5;       // constants.h:25-26 (DOUBLE_PRECISION)
constexpr Float kDeltaEpsilon = (Float)1e-3f;                 // constants.h:31 (float literal)
constexpr Float kDEps = 1e-14;                                // gpt.cpp:63 D_EPSILON
constexpr Float kPi = 3.14159265358979323846, kInvPi = 0.31830988618379067154, kInvTwoPi = 0.15915494309189533577;

struct V3 { Float x, y, z; };
GDB_HD V3 mk(Float x, Float y, Float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
GDB_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
GDB_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
GDB_HD V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
GDB_HD V3 operator*(V3 a, Float f) { return mk(a.x * f, a.y * f, a.z * f); }
GDB_HD V3 operator*(Float f, V3 a) { return mk(a.x * f, a.y * f, a.z * f); }
GDB_HD V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
GDB_HD V3 operator/(V3 a, Float f) { Float r = (Float)1 / f; return mk(a.x * r, a.y * r, a.z * r); }   // vector.h:535-542
GDB_HD V3 cdiv(V3 a, V3 b) { return mk(a.x / b.x, a.y / b.y, a.z / b.z); }
GDB_HD Float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GDB_HD V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GDB_HD Float len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
GDB_HD Float len(V3 a) { return sqrt(len2(a)); }
GDB_HD V3 normalize(V3 a) { return a / len(a); }
GDB_HD bool isZero(V3 a) { return a.x == 0 && a.y == 0 && a.z == 0; }
GDB_HD Float maxComp(V3 a) { return fmax(a.x, fmax(a.y, a.z)); }
GDB_HD V3 splat(Float v) { return mk(v, v, v); }
GDB_HD V3 safeSqrt3(V3 s) { return mk(sqrt(fmax(0.0, s.x)), sqrt(fmax(0.0, s.y)), sqrt(fmax(0.0, s.z))); }
typedef V3 Spec;

struct Frame { V3 s, t, n; };
GDB_HD V3 toLocal(const Frame &f, V3 v) { return mk(dot(v, f.s), dot(v, f.t), dot(v, f.n)); }   // frame.h:74-80
GDB_HD V3 toWorld(const Frame &f, V3 v) { return f.s * v.x + f.t * v.y + f.n * v.z; }           // frame.h:83-85

GDB_HD V3 xfAffine(const Float *m, V3 p)     // transform.h:128-137
{
    return mk(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
              m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
GDB_HD V3 xfVector(const Float *m, V3 v)     // transform.h:175-183
{
    return mk(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z,
              m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
GDB_HD V3 xfPoint(const Float *m, V3 p)      // transform.h:108-125
{
    Float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    Float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    Float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    Float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w == 1.0) return mk(x, y, z);
    return mk(x, y, z) / w;
}
GDB_HD V3 xfNormal(const Float *inv, V3 v)   // transform.h:203-211
{
    return mk(inv[0] * v.x + inv[4] * v.y + inv[8] * v.z, inv[1] * v.x + inv[5] * v.y + inv[9] * v.z,
              inv[2] * v.x + inv[6] * v.y + inv[10] * v.z);
}
GDB_HD void coordinateSystem(V3 a, V3 &b, V3 &c)            // util.cpp:592-601
{
    if (fabs(a.x) > fabs(a.y)) { const Float invLen = 1.0 / sqrt(a.x * a.x + a.z * a.z); c = mk(a.z * invLen, 0.0, -a.x * invLen); }
    else { const Float invLen = 1.0 / sqrt(a.y * a.y + a.z * a.z); c = mk(0.0, a.z * invLen, -a.y * invLen); }
    b = cross(c, a);
}
GDB_HD void computeShadingFrame(V3 n, V3 dpdu, Frame &f)    // util.cpp:603-608
{
    f.n = n;
    f.s = normalize(dpdu - f.n * dot(f.n, dpdu));
    f.t = cross(f.n, f.s);
}

// ---------------------------------------------------------------- scene tables
enum : unsigned { EDiffuseReflection = 0x1, EGlossyReflection = 0x4, EGlossyTransmission = 0x8, EDeltaReflection = 0x10, EDeltaTransmission = 0x20,
                  ESmooth = 0xF, EDelta = 0x30, ETransmissionBits = 0x2 | 0x8 | 0x20, EBackSide = 0x20000, EFrontSide = 0x10000 };
enum Measure { ESolidAngle = 0, EDiscrete = 1 };
enum VertexType { VERTEX_TYPE_GLOSSY = 0, VERTEX_TYPE_DIFFUSE = 1 };

constexpr int kMaxRects = 24, kMaxSpheres = 8, kMaxTris = 192, kMaxMeshes = 16, kMaxMaterials = 32, kMaxEmitters = 8;

struct DRect   { Float toObject[12], toWorld[12]; V3 dpdu, n; Float invArea; int material, emitter; };
struct DSphere { V3 center; Float radius; int flip, material, emitter, pad; };
struct DTri    { Float n_u, n_v, n_d, a_u, a_v, b_nu, b_nv, c_nu, c_nv; V3 p0, p1, p2, faceNormal; int k, material, emitter, normals; };   // normals: first of 3 vertex normals in triNormals, or -1 (flat)
struct DMesh   { V3 lo, hi; int first, count; int kEnd[3], pad; };   // triangles of a mesh are stored grouped by projection axis k: [first,kEnd[0]) k=0, [kEnd[0],kEnd[1]) k=1, [kEnd[1],kEnd[2]) k=2   // conservative (enlarged) bounds of one TriMesh, used only to skip its triangles
struct DMaterial {
    int type, distribution; unsigned flags; int vtSmooth, vtDelta, refNFromShading, twosided, nonlinear;
    Spec reflectance, specR, specT, eta, k; Float alpha, iorRatio, bsdfEta;
    Float fdrInt, fdrExt, specSamplingWeight, invEta2;       // plastic.cpp:188-206
};
enum { EM_RECT = 0, EM_MESH = 1, EM_ENV = 2, EM_POINT = 3, EM_SPHERE = 4, EM_SPOT = 5 };
// pdfDiscrete = samplingWeight * normalization (scene.h:855-857).  Mesh emitters: triangles [triFirst, triFirst+triCount) of
// emTris in the mesh's own order, area CDF (triCount+1 entries) at emTriCdf[cdfFirst], invArea = 1 / surface area.
struct DEmitter {
    int kind, rect, triFirst, triCount, cdfFirst, pad; Spec radiance; Float pdfDiscrete, invArea; V3 position;
    Float toLocal[9], cutoffAngle, cosCutoff, cosBeam, invTransition;   // spot (spot.cpp:89-94): inverse rotation rows, cone angles
};
struct DEmTri { V3 p0, p1, p2; int normals, pad; };   // normals: as DTri
// Environment map (envmap.cpp): top-level texels as Float RGB, the float CDF tables of envmap.cpp:263-311.
struct DEnv {
    int present, width, height, emitter;
    Float toWorld[12], toObject[12]; V3 center; Float radius, scale, normalization, pixelSizeX, pixelSizeY;
    const float *cdfRows, *cdfCols; const Float *rowWeights, *texels;
};
// BVH2 over the triangles of large meshes (scenes beyond the constant-memory table).  Bounds are single precision and
// padded like DBounds (conservative); inner: a/b = children; leaf: a = first triangle, b = -count.
struct BvhNode { float lo[3]; int a; float hi[3]; int b; };

struct DScene {
    Float sampleToCamera[16], cameraToWorld[12];
    Float nearClip, farClip, invResX, invResY, filterRadius, filterScale, apertureRadius, focusDistance;
    Float filterTable[32];             // ReconstructionFilter::m_values (rfilter.cpp:37-55)
    int filterIsBox, padFilter;        // box: every tap below index 31 is filterTable[0] (no per-lane table read)
    int width, height, nRects, nSpheres, nTris, nMaterials, nEmitters, nMeshes;
    Float emCdf[kMaxEmitters + 1];
    DRect rects[kMaxRects];
    DSphere spheres[kMaxSpheres];
    DMaterial materials[kMaxMaterials];
    DEmitter emitters[kMaxEmitters];
    DMesh meshes[kMaxMeshes];
    DTri tris[kMaxTris];
    DEnv env;
    const DEmTri *emTris; const Float *emTriCdf;
    const V3 *triNormals;              // vertex normals of smooth-shaded triangles, 3 per triangle
    const BvhNode *bvh; const DTri *bvhTris; int nBvhNodes, nBvhTris;
};

__constant__ DScene c_scene;   // one per device; renders sharing a device are serialised on the host

// The same tables in global memory.  Constant memory serialises a warp whose lanes read different
// entries, so every access with a per-lane index (the hit primitive, the material of a vertex, a lane's
// candidate primitive) goes through this copy (L1-cached); c_scene serves the warp-uniform loops.
__constant__ const DScene *c_sceneG;

// Padded world-space bounds of every primitive, in test order: rectangles, spheres, triangles.
constexpr int kMaxPrims = kMaxRects + kMaxSpheres + kMaxTris;
struct DBounds { float lo[3], hi[3]; };   // single precision is enough for a conservative (padded) slab test
__constant__ DBounds c_bounds[kMaxPrims];

struct Config { int maxDepth, minDepth, rrDepth, strictNormals; Float shiftThreshold; int refUninitMeasure, pad; };   // refUninitMeasure: see gpt_host.h setupArgs

struct Its { Float t; V3 p, geoN; Frame sh; V3 wi; int material, emitter; };   // emitter: index or -1
struct Ray { V3 o, d; Float mint, maxt; };

// ---------------------------------------------------------------- intersection
GDB_D Float maxAbs3(V3 o) { return fmax(fmax(fabs(o.x), fabs(o.y)), fabs(o.z)); }

// util.cpp:487-525
GDB_D bool solveQuadratic(double a, double b, double c, double &x0, double &x1)
{
    if (a == 0) { if (b != 0) { x0 = x1 = -c / b; return true; } return false; }
    double discrim = b * b - 4.0 * a * c;
    if (discrim < 0) return false;
    double temp, sqrtDiscrim = sqrt(discrim);
    if (b < 0) temp = -0.5 * (b - sqrtDiscrim); else temp = -0.5 * (b + sqrtDiscrim);
    x0 = temp / a; x1 = c / temp;
    if (x0 > x1) { double t = x0; x0 = x1; x1 = t; }
    return true;
}

// Exact fp64 tests of the reference, one primitive each.
GDB_D bool rectHit(const DRect &r, const Ray &ray, Float mint, Float maxt, Float &t)          // rectangle.cpp:125-151
{
    const Float *m = r.toObject;
    const Float oz = m[8] * ray.o.x + m[9] * ray.o.y + m[10] * ray.o.z + m[11];
    const Float dz = m[8] * ray.d.x + m[9] * ray.d.y + m[10] * ray.d.z;
    const Float hit = -oz / dz;
    if (!(hit >= mint && hit <= maxt)) return false;
    const Float lx = (m[0] * ray.o.x + m[1] * ray.o.y + m[2] * ray.o.z + m[3]) + (m[0] * ray.d.x + m[1] * ray.d.y + m[2] * ray.d.z) * hit;
    const Float ly = (m[4] * ray.o.x + m[5] * ray.o.y + m[6] * ray.o.z + m[7]) + (m[4] * ray.d.x + m[5] * ray.d.y + m[6] * ray.d.z) * hit;
    if (fabs(lx) <= 1 && fabs(ly) <= 1) { t = hit; return true; }
    return false;
}
GDB_D bool sphereHit(const DSphere &s, const Ray &ray, Float mint, Float maxt, Float &t)      // sphere.cpp:163-187
{
    const V3 o = ray.o - s.center;
    const double A = len2(ray.d), B = 2 * dot(o, ray.d), C = len2(o) - s.radius * s.radius;
    double nearT, farT;
    if (!solveQuadratic(A, B, C, nearT, farT)) return false;
    if (!(nearT <= maxt && farT >= mint)) return false;
    if (nearT < mint) { if (farT > maxt) return false; t = farT; } else t = nearT;
    return true;
}
GDB_D bool triHit(const DTri &T, const Ray &ray, Float mint, Float maxt, Float &t, Float &u, Float &v)   // triaccel.h:97-158
{
    const int k = T.k;
    if (k > 2) return false;
    const Float o_u = k == 0 ? ray.o.y : (k == 1 ? ray.o.z : ray.o.x), o_v = k == 0 ? ray.o.z : (k == 1 ? ray.o.x : ray.o.y);
    const Float o_k = k == 0 ? ray.o.x : (k == 1 ? ray.o.y : ray.o.z);
    const Float d_u = k == 0 ? ray.d.y : (k == 1 ? ray.d.z : ray.d.x), d_v = k == 0 ? ray.d.z : (k == 1 ? ray.d.x : ray.d.y);
    const Float d_k = k == 0 ? ray.d.x : (k == 1 ? ray.d.y : ray.d.z);
    t = (T.n_d - o_u * T.n_u - o_v * T.n_v - o_k) / (d_u * T.n_u + d_v * T.n_v + d_k);
    if (t < mint || t > maxt) return false;
    const Float hu = o_u + t * d_u - T.a_u, hv = o_v + t * d_v - T.a_v;
    u = hv * T.b_nu + hu * T.b_nv;
    v = hu * T.c_nu + hv * T.c_nv;
    return u >= 0 && v >= 0 && u + v <= 1.0;
}

// Large meshes: per-lane BVH2 walk (fp32 padded node bounds, children visited near to far), exact fp64 triangle tests in
// the leaves.  Same answers as testing every triangle: the bounds are conservative.  Continues a search that has already
// narrowed [mint, maxt]; returns true only for an any-hit query that found an occluder.
template <bool AnyHit>
GDB_D bool walkBvh(const Ray &ray, Float mint, Float &maxt, bool &found, int &kind, int &index, Float &uOut, Float &vOut)
{
    const float ox = (float)ray.o.x, oy = (float)ray.o.y, oz = (float)ray.o.z;
    const float ix = 1.0f / (float)ray.d.x, iy = 1.0f / (float)ray.d.y, iz = 1.0f / (float)ray.d.z;
    const float tlo = (float)mint * 0.999f, thi0 = (float)maxt * 1.001f;
    const BvhNode *nodes = c_scene.bvh;
    auto slab = [&](const BvhNode &N, float thi, float &tnear) {
        const float ax = (N.lo[0] - ox) * ix, bx = (N.hi[0] - ox) * ix;
        const float ay = (N.lo[1] - oy) * iy, by = (N.hi[1] - oy) * iy;
        const float az = (N.lo[2] - oz) * iz, bz = (N.hi[2] - oz) * iz;
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
        const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        tnear = tn;
        return tn <= tf && tf >= tlo && tn <= thi;
    };
    int stack[64], sp = 0;
    float t0;
    if (slab(nodes[0], thi0, t0)) stack[sp++] = 0;
    while (sp > 0) {
        const BvhNode N = nodes[stack[--sp]];
        const float thi = (float)maxt * 1.001f;
        if (N.b < 0) {
            for (int p = N.a; p < N.a - N.b; p++) {
                Float t, u, v;
                if (triHit(c_scene.bvhTris[p], ray, mint, maxt, t, u, v)) { if (AnyHit) return true; maxt = t; found = true; kind = 3; index = p; uOut = u; vOut = v; }
            }
        } else {
            float ta, tb;
            const bool ha = slab(nodes[N.a], thi, ta), hb = slab(nodes[N.b], thi, tb);
            if (ha && hb) {
                const bool aFirst = ta <= tb;
                if (sp + 2 <= 64) { stack[sp++] = aFirst ? N.b : N.a; stack[sp++] = aFirst ? N.a : N.b; }
            } else if (ha) { if (sp < 64) stack[sp++] = N.a; }
            else if (hb) { if (sp < 64) stack[sp++] = N.b; }
        }
    }
    return false;
}

// Nearest (or any) hit in [mint, maxt] = the outcome of the reference's kd-tree traversal
// (sahkdtree3.h:179-308: every candidate is tested against the shrinking interval).
//   pass 1: a division-free slab test of the ray against the padded bounds of every primitive, as a
//           warp-uniform loop over constant memory, yields a per-lane candidate mask;
//   pass 2: each lane walks ITS OWN candidates (typically 2-4 of 32) through the reference's exact
//           fp64 test, in primitive order.  All lanes run the same test code on different primitives,
//           so incoherent rays no longer pay for the union of everything any lane might hit.
// The bounds are conservative, so the answers are those of testing every primitive.
#ifndef GDB_SLAB_NO_FMA
#define GDB_SLAB_FMA 1      // the cast kernels are issue-bound: one FFMA per plane instead of FADD + FMUL is worth 8 % of them
#endif
template <bool AnyHit>
GDB_D bool closestPrimitive(const Ray &ray, Float mint, Float maxt, Float &tOut, int &kind, int &index, Float &uOut, Float &vOut)
{
    const DScene *g = c_sceneG;
    const int nR = c_scene.nRects, nS = c_scene.nSpheres, nP = nR + nS + c_scene.nTris;
    const float ox = (float)ray.o.x, oy = (float)ray.o.y, oz = (float)ray.o.z;
#ifdef GDB_SLAB_FMA
    // (B - o) * i as one FFMA per plane: B * i + (-o * i).  1/d is clamped to +-1e30 so an axis-parallel ray gives
    // huge finite plane distances instead of inf - inf; the rounding of o*i (<= 2^-24 |o| |i|) is far inside the
    // padding of the bounds (1e-4 of the scene scale, times |i|).
    const float ix = fminf(fmaxf(1.0f / (float)ray.d.x, -1e30f), 1e30f), iy = fminf(fmaxf(1.0f / (float)ray.d.y, -1e30f), 1e30f),
                iz = fminf(fmaxf(1.0f / (float)ray.d.z, -1e30f), 1e30f);
    const float nx = -(ox * ix), ny = -(oy * iy), nz = -(oz * iz);
#else
    const float ix = 1.0f / (float)ray.d.x, iy = 1.0f / (float)ray.d.y, iz = 1.0f / (float)ray.d.z;
#endif
    const float tlo = (float)mint * 0.999f, thi0 = (float)maxt * 1.001f;      // (+inf stays +inf)
    bool found = false;
    for (int base = 0; base < nP; base += 32) {
        const int cnt = min(32, nP - base);
        const float thi = found ? (float)maxt * 1.001f : thi0;
        unsigned mask = 0;
        for (int j = 0; j < cnt; j++) {
            const DBounds &B = c_bounds[base + j];
#ifdef GDB_SLAB_FMA
            const float ax = __fmaf_rn(B.lo[0], ix, nx), bx = __fmaf_rn(B.hi[0], ix, nx);
            const float ay = __fmaf_rn(B.lo[1], iy, ny), by = __fmaf_rn(B.hi[1], iy, ny);
            const float az = __fmaf_rn(B.lo[2], iz, nz), bz = __fmaf_rn(B.hi[2], iz, nz);
            const float tn = fmaxf(fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)), tlo);
            const float tf = fminf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)), thi);
            if (tn <= tf) mask |= 1u << j;
#else
            const float ax = (B.lo[0] - ox) * ix, bx = (B.hi[0] - ox) * ix;
            const float ay = (B.lo[1] - oy) * iy, by = (B.hi[1] - oy) * iy;
            const float az = (B.lo[2] - oz) * iz, bz = (B.hi[2] - oz) * iz;
            const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
            const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
            if (tn <= tf && tf >= tlo && tn <= thi) mask |= 1u << j;
#endif
        }
        // rectangles
        const int rEnd = min(max(nR - base, 0), 32), sEnd = min(max(nR + nS - base, 0), 32);
        unsigned m = mask & (rEnd >= 32 ? 0xffffffffu : ((1u << rEnd) - 1));
        while (m) {
            const int p = base + __ffs(m) - 1; m &= m - 1;
            Float t;
            if (rectHit(g->rects[p], ray, mint, maxt, t)) { if (AnyHit) return true; maxt = t; found = true; kind = 0; index = p; }
        }
        m = mask & (sEnd >= 32 ? 0xffffffffu : ((1u << sEnd) - 1)) & ~(rEnd >= 32 ? 0xffffffffu : ((1u << rEnd) - 1));
        while (m) {
            const int p = base + __ffs(m) - 1; m &= m - 1;
            Float t;
            if (sphereHit(g->spheres[p - nR], ray, mint, maxt, t)) { if (AnyHit) return true; maxt = t; found = true; kind = 1; index = p - nR; }
        }
        m = mask & ~(sEnd >= 32 ? 0xffffffffu : ((1u << sEnd) - 1));
        while (m) {
            const int p = base + __ffs(m) - 1; m &= m - 1;
            Float t, u, v;
            if (triHit(g->tris[p - nR - nS], ray, mint, maxt, t, u, v)) { if (AnyHit) return true; maxt = t; found = true; kind = 2; index = p - nR - nS; uOut = u; vOut = v; }
        }
    }
    if (c_scene.nBvhNodes > 0 && walkBvh<AnyHit>(ray, mint, maxt, found, kind, index, uOut, vOut)) return true;
    tOut = maxt;
    return found;
}

// Reference-order brute force over every primitive: used by the self-check kernel only.
template <bool AnyHit>
GDB_D bool closestPrimitiveExhaustive(const Ray &ray, Float mint, Float maxt, Float &tOut, int &kind, int &index, Float &uOut, Float &vOut)
{
    bool found = false;
    for (int i = 0; i < c_scene.nRects; i++) { Float t; if (rectHit(c_scene.rects[i], ray, mint, maxt, t)) { if (AnyHit) return true; maxt = t; found = true; kind = 0; index = i; } }
    for (int i = 0; i < c_scene.nSpheres; i++) { Float t; if (sphereHit(c_scene.spheres[i], ray, mint, maxt, t)) { if (AnyHit) return true; maxt = t; found = true; kind = 1; index = i; } }
    for (int i = 0; i < c_scene.nTris; i++) { Float t, u, v; if (triHit(c_scene.tris[i], ray, mint, maxt, t, u, v)) { if (AnyHit) return true; maxt = t; found = true; kind = 2; index = i; uOut = u; vOut = v; } }
    for (int i = 0; i < c_scene.nBvhTris; i++) { Float t, u, v; if (triHit(c_scene.bvhTris[i], ray, mint, maxt, t, u, v)) { if (AnyHit) return true; maxt = t; found = true; kind = 3; index = i; uOut = u; vOut = v; } }
    tOut = maxt;
    return found;
}

// ShapeKDTree::rayIntersect(ray, its): skdtree.cpp:112-147, record fill skdtree.h:343-428.  Two halves, so that the
// staged wavefront (gpt_stages.cuh) can run the search in its own kernel (gpt_cast_kernel) and rebuild the intersection
// record where it is consumed: a hit travels as {t, u, v, primitive}.
struct Hit { Float t, u, v; int kind, index; };   // kind: 0 rectangle, 1 sphere, 2 table triangle, 3 BVH triangle; t = +inf: miss
GDB_D bool castClosest(const Ray &ray, Hit &h)
{
    h.t = CUDART_INF; h.u = 0; h.v = 0; h.kind = 0; h.index = 0;
    Float rayMinT = ray.mint;
    if (rayMinT == kEpsilon) rayMinT *= fmax(maxAbs3(ray.o), kEpsilon);
    if (!(ray.maxt > rayMinT)) return false;
    Float t;
    if (!closestPrimitive<false>(ray, rayMinT, ray.maxt, t, h.kind, h.index, h.u, h.v)) return false;
    h.t = t;
    return true;
}
GDB_D void fillIts(const Ray &ray, const Hit &h, Its &its)
{
    const Float t = h.t, u = h.u, v = h.v; const int kind = h.kind, index = h.index;
    its.t = t;
    V3 dpdu;
    if (kind >= 2) {                                                 // skdtree.h:348-419 (BarycentricPos)
        const DTri &T = kind == 2 ? c_sceneG->tris[index] : c_scene.bvhTris[index];
        const V3 b = mk(1 - u - v, u, v);
        its.p = T.p0 * b.x + T.p1 * b.y + T.p2 * b.z;
        dpdu = T.p1 - T.p0;
        its.sh.n = T.faceNormal; its.geoN = T.faceNormal;
        if (T.normals >= 0) {                                         // skdtree.h:383-394
            const V3 *vn = c_scene.triNormals + T.normals;
            its.sh.n = normalize(vn[0] * b.x + vn[1] * b.y + vn[2] * b.z);
            if (dot(its.geoN, its.sh.n) < 0) its.geoN = -its.geoN;
        }
        its.material = T.material; its.emitter = T.emitter;
    } else if (kind == 0) {                                          // rectangle.cpp:158-171
        const DRect &r = c_sceneG->rects[index];
        its.geoN = r.n; its.sh.n = r.n; dpdu = r.dpdu;
        its.p = ray.o + ray.d * t;
        its.material = r.material; its.emitter = r.emitter;
    } else {                                                         // sphere.cpp:197-240
        const DSphere &s = c_sceneG->spheres[index];
        its.p = ray.o + ray.d * t;
        const V3 local = its.p - s.center;
        dpdu = mk(-local.y, local.x, 0) * (2 * kPi);
        its.geoN = normalize(its.p - s.center);
        if (s.flip) its.geoN = its.geoN * -1.0;
        its.sh.n = its.geoN;
        its.material = s.material; its.emitter = s.emitter;
    }
    computeShadingFrame(its.sh.n, dpdu, its.sh);                     // skdtree.h:425
    its.wi = toLocal(its.sh, -ray.d);                                // skdtree.h:426
}
GDB_D bool rayIntersectImpl(const Ray &ray, Its &its)
{
    Hit h;
    if (!castClosest(ray, h)) { its.t = CUDART_INF; return false; }
    fillIts(ray, h, its);
    return true;
}

// ShapeKDTree::rayIntersect(ray) for shadow rays: skdtree.cpp:206-226
GDB_D bool rayOccludedImpl(const Ray &ray)
{
    Float rayMinT = ray.mint;
    if (rayMinT == kEpsilon) rayMinT *= maxAbs3(ray.o);
    if (!(ray.maxt > rayMinT)) return false;
    Float t, u, v; int a, b;
    return closestPrimitive<true>(ray, rayMinT, ray.maxt, t, a, b, u, v);
}

// The out-of-line entry points take and return their structs BY VALUE: the device ABI then passes them in registers,
// whereas a `const Ray &` / `Its &` parameter forces the caller's struct into local memory (stack traffic through
// L1/L2 on every call: 44 % of the bounce kernel's L2 sectors before this change, profiles/r01b_*).
// Measured (profiles/r01_gpt_history.md, rows 17-18): by-value pays in gpt_generate_kernel (five back-to-back
// intersections per sample, -13 % kernel time) and costs in gpt_bounce_kernel (the returned record raises the register
// pressure around the call, +7 %), so both forms of rayIntersect exist and each kernel calls the one that suits it.
// GDB_BYVAL switches the remaining routines for A/B builds (bit 1 rayOccluded, 2 bsdfEvalPdf, 3 bsdfSample, 4 sampleEmitterDirectVisible).
#ifndef GDB_BYVAL
#define GDB_BYVAL 0
#endif
GDB_CALL Its rayIntersectV(Ray ray) { Its its; rayIntersectImpl(ray, its); return its; }
GDB_D bool rayIntersectByValue(const Ray &ray, Its &its)
{
    const Its r = rayIntersectV(ray);
    if (r.t == CUDART_INF) { its.t = CUDART_INF; return false; }
    its = r;
    return true;
}
GDB_CALL bool rayIntersect(const Ray &ray, Its &its) { return rayIntersectImpl(ray, its); }
#if GDB_BYVAL & 2
GDB_CALL bool rayOccludedV(Ray ray) { return rayOccludedImpl(ray); }
GDB_D bool rayOccluded(const Ray &ray) { return rayOccludedV(ray); }
#else
GDB_CALL bool rayOccluded(const Ray &ray) { return rayOccludedImpl(ray); }
#endif

// ---------------------------------------------------------------- sampling helpers
GDB_D void squareToUniformDiskConcentric(Float sx, Float sy, Float &ox, Float &oy)   // warp.cpp:81-102
{
    Float r1 = 2.0 * sx - 1.0, r2 = 2.0 * sy - 1.0, phi, r;
    if (r1 == 0 && r2 == 0) { r = phi = 0; }
    else if (r1 * r1 > r2 * r2) { r = r1; phi = (kPi / 4.0) * (r2 / r1); }
    else { r = r2; phi = (kPi / 2.0) - (r1 / r2) * (kPi / 4.0); }
    ox = r * cos(phi); oy = r * sin(phi);
}
GDB_D V3 squareToCosineHemisphere(Float sx, Float sy)                                 // warp.cpp:43-52
{
    Float px, py;
    squareToUniformDiskConcentric(sx, sy, px, py);
    Float z = sqrt(fmax(0.0, 1.0 - px * px - py * py));
    if (z == 0) z = (Float)1e-10f;
    return mk(px, py, z);
}

// ---------------------------------------------------------------- microfacet.h (isotropic, visible normals)
GDB_D Float mfEval(int type, Float alpha, V3 m)                                       // :191-235
{
    if (m.z <= 0) return 0.0;
    const Float cosTheta2 = m.z * m.z;
    const Float beckmannExponent = ((m.x * m.x) / (alpha * alpha) + (m.y * m.y) / (alpha * alpha)) / cosTheta2;
    Float result;
    if (type == GDB200_MICROFACET_BECKMANN) result = exp(-beckmannExponent) / (kPi * alpha * alpha * cosTheta2 * cosTheta2);
    else { const Float root = ((Float)1 + beckmannExponent) * cosTheta2; result = (Float)1 / (kPi * alpha * alpha * root * root); }
    if (result * m.z < (Float)1e-20f) result = 0;
    return result;
}
GDB_D Float mfSmithG1(int type, Float alpha, V3 v, V3 m)                              // :470-508
{
    if (dot(v, m) * v.z <= 0) return 0.0;
    const Float temp = 1 - v.z * v.z;
    const Float tanTheta = temp <= 0.0 ? 0.0 : fabs(sqrt(temp) / v.z);
    if (tanTheta == 0.0) return 1.0;
    if (type == GDB200_MICROFACET_BECKMANN) {
        const Float a = 1.0 / (alpha * tanTheta);
        if (a >= (Float)1.6f) return 1.0;
        const Float aSqr = a * a;
        return ((Float)3.535f * a + (Float)2.181f * aSqr) / (1.0 + (Float)2.276f * a + (Float)2.577f * aSqr);
    }
    const Float root = alpha * tanTheta;                                             // math.cpp:89-101 hypot2(1, root)
    Float r;
    if (1.0 > fabs(root)) { r = root / 1.0; r = 1.0 * sqrt(1.0 + r * r); }
    else if (root != 0.0) { r = 1.0 / root; r = fabs(root) * sqrt(1.0 + r * r); }
    else r = 0.0;
    return 2.0 / (1.0 + r);
}
GDB_D Float mfPdfVisible(int type, Float alpha, V3 wi, V3 m)                          // :455-459
{
    if (wi.z == 0) return 0.0;
    return mfSmithG1(type, alpha, wi, m) * fabs(dot(wi, m)) * mfEval(type, alpha, m) / fabs(wi.z);
}
GDB_D Float mtsErf(Float x)                                                           // math.cpp:55-72
{
    const Float a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027, a5 = 1.061405429, p = 0.3275911;
    const Float sign = copysign(1.0, x);
    x = fabs(x);
    const Float t = 1.0 / (1.0 + p * x);
    const Float y = 1.0 - (((((a5 * t + a4) * t) + a3) * t + a2) * t + a1) * t * exp(-x * x);
    return sign * y;
}
GDB_D Float mtsErfinv(Float x)                                                        // math.cpp:25-53
{
    Float w = -log(((Float)1 - x) * ((Float)1 + x)), p;
    if (w < (Float)5) {
        w = w - (Float)2.5;
        p = (Float)2.81022636e-08; p = (Float)3.43273939e-07 + p * w; p = (Float)-3.5233877e-06 + p * w;
        p = (Float)-4.39150654e-06 + p * w; p = (Float)0.00021858087 + p * w; p = (Float)-0.00125372503 + p * w;
        p = (Float)-0.00417768164 + p * w; p = (Float)0.246640727 + p * w; p = (Float)1.50140941 + p * w;
    } else {
        w = sqrt(w) - (Float)3;
        p = (Float)-0.000200214257; p = (Float)0.000100950558 + p * w; p = (Float)0.00134934322 + p * w;
        p = (Float)-0.00367342844 + p * w; p = (Float)0.00573950773 + p * w; p = (Float)-0.0076224613 + p * w;
        p = (Float)0.00943887047 + p * w; p = (Float)1.00167406 + p * w; p = (Float)2.83297682 + p * w;
    }
    return p * x;
}
GDB_D void mfSampleVisible11(int type, Float thetaI, Float sx, Float sy, Float &slopeX, Float &slopeY)   // :573-696
{
    const Float SQRT_PI_INV = 1 / sqrt(kPi);
    if (type == GDB200_MICROFACET_BECKMANN) {
        if (thetaI < (Float)1e-4f) {
            const Float r = sqrt(-log(1.0 - sx));
            const Float sinPhi = sin(2 * kPi * sy), cosPhi = cos(2 * kPi * sy);
            slopeX = r * cosPhi; slopeY = r * sinPhi; return;
        }
        const Float tanThetaI = tan(thetaI), cotThetaI = 1 / tanThetaI;
        Float a = -1, c = mtsErf(cotThetaI);
        const Float sample_x = fmax(sx, (Float)1e-6f);
        const Float fit = 1 + thetaI * ((Float)-0.876f + thetaI * ((Float)0.4265f - (Float)0.0594f * thetaI));
        Float b = c - (1 + c) * pow(1 - sample_x, fit);
        const Float normalization = 1 / (1 + c + SQRT_PI_INV * tanThetaI * exp(-cotThetaI * cotThetaI));
        int it = 0;
        while (++it < 10) {
            if (!(b >= a && b <= c)) b = 0.5 * (a + c);
            const Float invErf = mtsErfinv(b);
            const Float value = normalization * (1 + b + SQRT_PI_INV * tanThetaI * exp(-invErf * invErf)) - sample_x;
            const Float derivative = normalization * (1 - invErf * tanThetaI);
            if (fabs(value) < (Float)1e-5f) break;
            if (value > 0) c = b; else a = b;
            b -= value / derivative;
        }
        slopeX = mtsErfinv(b);
        slopeY = mtsErfinv(2.0 * fmax(sy, (Float)1e-6f) - 1.0);
        return;
    }
    if (thetaI < (Float)1e-4f) {
        const Float r = sqrt(fmax(0.0, sx / (1 - sx)));
        const Float sinPhi = sin(2 * kPi * sy), cosPhi = cos(2 * kPi * sy);
        slopeX = r * cosPhi; slopeY = r * sinPhi; return;
    }
    const Float tanThetaI = tan(thetaI);
    const Float a = 1 / tanThetaI;
    const Float G1 = 2.0 / (1.0 + sqrt(fmax(0.0, 1.0 + 1.0 / (a * a))));
    Float A = 2.0 * sx / G1 - 1.0;
    if (fabs(A) == 1) A -= copysign(1.0, A) * kEpsilon;
    const Float tmp = 1.0 / (A * A - 1.0);
    const Float B = tanThetaI;
    const Float D = sqrt(fmax(0.0, B * B * tmp * tmp - (A * A - B * B) * tmp));
    const Float slope_x_1 = B * tmp - D, slope_x_2 = B * tmp + D;
    slopeX = (A < 0.0 || slope_x_2 > 1.0 / tanThetaI) ? slope_x_1 : slope_x_2;
    Float S;
    if (sy > 0.5) { S = 1.0; sy = 2.0 * (sy - 0.5); } else { S = -1.0; sy = 2.0 * (0.5 - sy); }
    const Float z = (sy * (sy * (sy * (-(Float)0.365728915865723) + (Float)0.790235037209296) - (Float)0.424965825137544) + (Float)0.000152998850436920) /
                    (sy * (sy * (sy * (sy * (Float)0.169507819808272 - (Float)0.397203533833404) - (Float)0.232500544458471) + (Float)1) - (Float)0.539825872510702);
    slopeY = S * z * sqrt(1.0 + slopeX * slopeX);
}
GDB_D V3 mfSampleVisible(int type, Float alpha, V3 _wi, Float sx, Float sy)           // :421-452
{
    const V3 wi = normalize(mk(alpha * _wi.x, alpha * _wi.y, _wi.z));
    Float theta = 0, phi = 0;
    if (wi.z < (Float)0.99999) { theta = acos(wi.z); phi = atan2(wi.y, wi.x); }
    const Float sinPhi = sin(phi), cosPhi = cos(phi);
    Float slx, sly;
    mfSampleVisible11(type, theta, sx, sy, slx, sly);
    Float rx = cosPhi * slx - sinPhi * sly, ry = sinPhi * slx + cosPhi * sly;
    rx *= alpha; ry *= alpha;
    const Float normalization = (Float)1 / sqrt(rx * rx + ry * ry + (Float)1.0);
    return mk(-rx * normalization, -ry * normalization, normalization);
}

GDB_D Spec fresnelConductorExact(Float cosThetaI, Spec eta, Spec k)                   // util.cpp:739-761
{
    const Float cosThetaI2 = cosThetaI * cosThetaI, sinThetaI2 = 1 - cosThetaI2, sinThetaI4 = sinThetaI2 * sinThetaI2;
    const Spec temp1 = eta * eta - k * k - splat(sinThetaI2);
    const Spec a2pb2 = safeSqrt3(temp1 * temp1 + k * k * eta * eta * 4.0);
    const Spec a = safeSqrt3((a2pb2 + temp1) * 0.5);
    const Spec term1 = a2pb2 + splat(cosThetaI2), term2 = a * (2 * cosThetaI);
    const Spec Rs2 = cdiv(term1 - term2, term1 + term2);
    const Spec term3 = a2pb2 * cosThetaI2 + splat(sinThetaI4), term4 = term2 * sinThetaI2;
    const Spec Rp2 = cdiv(Rs2 * (term3 - term4), term3 + term4);
    return 0.5 * (Rp2 + Rs2);
}
GDB_D Float fresnelDielectricExt(Float cosThetaI_, Float &cosThetaT_, Float eta)      // util.cpp:651-681
{
    if (eta == 1) { cosThetaT_ = -cosThetaI_; return 0.0; }
    const Float scale = (cosThetaI_ > 0) ? 1 / eta : eta, cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0) { cosThetaT_ = 0.0; return 1.0; }
    const Float cosThetaI = fabs(cosThetaI_), cosThetaT = sqrt(cosThetaTSqr);
    const Float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    const Float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5 * (Rs * Rs + Rp * Rp);
}
GDB_D Float fresnelDielectricExt(Float cosThetaI, Float eta) { Float cosThetaT; return fresnelDielectricExt(cosThetaI, cosThetaT, eta); }
GDB_D V3 reflectAboutLocal(V3 wi, V3 n) { return 2 * dot(wi, n) * n - wi; }                // util.cpp:763-765
GDB_D V3 reflectLocal(V3 wi) { return mk(-wi.x, -wi.y, wi.z); }
GDB_D V3 refractLocal(const DMaterial &m, V3 wi, Float cosThetaT)                     // dielectric.cpp:223-226
{
    const Float scale = -(cosThetaT < 0 ? 1.0 / m.iorRatio : m.iorRatio);
    return mk(scale * wi.x, scale * wi.y, cosThetaT);
}

// ---------------------------------------------------------------- BSDF eval / pdf / sample
// Evaluates f*cos and the solid-angle (or discrete) density together: every call site of the
// reference asks for both (gpt.cpp:588-592,645-647,693-694,871-872,935-936,1030-1031).
GDB_D void bsdfEvalPdfImpl(const DMaterial &m, V3 wi, V3 wo, int measure, Spec &value, Float &pdf)
{
    value = splat(0); pdf = 0;
    if (m.twosided && !(wi.z > 0)) { wi.z *= -1; wo.z *= -1; }                         // twosided.cpp:109-135
    switch (m.type) {
    case GDB200_BSDF_DIFFUSE:                                                          // diffuse.cpp:110-129
        if (measure != ESolidAngle || wi.z <= 0 || wo.z <= 0) return;
        value = m.reflectance * (kInvPi * wo.z);
        pdf = kInvPi * wo.z;
        return;
    case GDB200_BSDF_ROUGHCONDUCTOR: {                                                 // roughconductor.cpp:256-320
        if (measure != ESolidAngle || wi.z <= 0 || wo.z <= 0) return;
        const V3 H = normalize(wo + wi);
        const Float D = mfEval(m.distribution, m.alpha, H);
        const Float G1i = mfSmithG1(m.distribution, m.alpha, wi, H);
        pdf = D * G1i / (4.0 * wi.z);
        if (D == 0) return;
        const Spec F = fresnelConductorExact(dot(wi, H), m.eta, m.k) * m.specR;
        const Float G = G1i * mfSmithG1(m.distribution, m.alpha, wo, H);
        const Float model = D * G / (4.0 * wi.z);
        value = F * model;
        return;
    }
    case GDB200_BSDF_CONDUCTOR:                                                        // conductor.cpp:221-250
        if (measure != EDiscrete || wi.z <= 0 || wo.z <= 0 || fabs(dot(reflectLocal(wi), wo) - 1) > kDeltaEpsilon) return;
        value = m.specR * fresnelConductorExact(wi.z, m.eta, m.k);
        pdf = 1.0;
        return;
    case GDB200_BSDF_ROUGHDIELECTRIC: {                                                // roughdielectric.cpp:277-422 (mode ERadiance)
        if (measure != ESolidAngle || wi.z == 0) return;
        const bool reflect = wi.z * wo.z > 0;
        const Float eta = wi.z > 0 ? m.iorRatio : 1 / m.iorRatio;
        V3 H = reflect ? normalize(wo + wi) : normalize(wi + wo * eta);
        Float dwh_dwo;
        if (reflect) dwh_dwo = 1.0 / (4.0 * dot(wo, H));
        else { const Float sd = dot(wi, H) + eta * dot(wo, H); dwh_dwo = (eta * eta * dot(wo, H)) / (sd * sd); }
        H = H * copysign(1.0, H.z);
        const Float D = mfEval(m.distribution, m.alpha, H);
        const Float F = fresnelDielectricExt(dot(wi, H), m.iorRatio);
        Float prob = mfPdfVisible(m.distribution, m.alpha, wi * copysign(1.0, wi.z), H);
        prob *= reflect ? F : (1 - F);
        pdf = fabs(prob * dwh_dwo);
        if (D == 0) return;
        const Float G = mfSmithG1(m.distribution, m.alpha, wi, H) * mfSmithG1(m.distribution, m.alpha, wo, H);
        if (reflect) {
            const Float v = F * D * G / (4.0 * fabs(wi.z));
            value = m.specR * v;
        } else {
            const Float sqrtDenom = dot(wi, H) + eta * dot(wo, H);
            const Float v = ((1 - F) * D * G * eta * eta * dot(wi, H) * dot(wo, H)) / (wi.z * sqrtDenom * sqrtDenom);
            const Float factor = wi.z > 0 ? 1 / m.iorRatio : m.iorRatio;
            value = m.specT * fabs(v * factor * factor);
        }
        return;
    }
    case GDB200_BSDF_PLASTIC: {                                                        // plastic.cpp:243-302 (typeMask = EAll, component = -1)
        if (wo.z <= 0 || wi.z <= 0) return;
        Float dummy;
        const Float Fi = fresnelDielectricExt(wi.z, dummy, m.iorRatio);
        const Float probSpecular = (Fi * m.specSamplingWeight) / (Fi * m.specSamplingWeight + (1 - Fi) * (1 - m.specSamplingWeight));
        if (measure == EDiscrete) {
            if (fabs(dot(reflectLocal(wi), wo) - 1) < kDeltaEpsilon) { value = m.specR * Fi; pdf = probSpecular; }
        } else {
            const Float Fo = fresnelDielectricExt(wo.z, dummy, m.iorRatio);
            Spec diff = m.reflectance;
            if (m.nonlinear) diff = cdiv(diff, splat(1.0) - diff * m.fdrInt); else diff = diff / (1 - m.fdrInt);
            value = diff * ((kInvPi * wo.z) * m.invEta2 * (1 - Fi) * (1 - Fo));
            pdf = (kInvPi * wo.z) * (1 - probSpecular);
        }
        return;
    }
    default: {                                                                         // dielectric.cpp:228-275
        Float cosThetaT;
        const Float F = fresnelDielectricExt(wi.z, cosThetaT, m.iorRatio);
        if (wi.z * wo.z >= 0) {
            if (measure != EDiscrete || fabs(dot(reflectLocal(wi), wo) - 1) > kDeltaEpsilon) return;
            value = m.specR * F; pdf = F;
        } else {
            if (measure != EDiscrete || fabs(dot(refractLocal(m, wi, cosThetaT), wo) - 1) > kDeltaEpsilon) return;
            const Float factor = cosThetaT < 0 ? 1.0 / m.iorRatio : m.iorRatio;
            value = m.specT * factor * factor * (1 - F); pdf = 1 - F;
        }
        return;
    }
    }
}

struct EvalPdf { Spec value; Float pdf; };
#if GDB_BYVAL & 4
GDB_CALL EvalPdf bsdfEvalPdfV(const DMaterial &m, V3 wi, V3 wo, int measure) { EvalPdf r; bsdfEvalPdfImpl(m, wi, wo, measure, r.value, r.pdf); return r; }
GDB_D void bsdfEvalPdf(const DMaterial &m, V3 wi, V3 wo, int measure, Spec &value, Float &pdf)
{
    const EvalPdf r = bsdfEvalPdfV(m, wi, wo, measure);
    value = r.value; pdf = r.pdf;
}
#else
GDB_CALL void bsdfEvalPdf(const DMaterial &m, V3 wi, V3 wo, int measure, Spec &value, Float &pdf) { bsdfEvalPdfImpl(m, wi, wo, measure, value, pdf); }
#endif

struct BSDFSample { V3 wo; Float eta; unsigned sampledType; Spec weight; Float pdf; int extraDraws; };   // extraDraws: sampler values consumed inside sample() (EUsesSampler)

// BSDF::sample(bRec, pdf, sample), pdf pre-set to 0 by the caller (gpt.cpp:450-457)
GDB_D void bsdfSampleOneSided(const DMaterial &m, V3 wi, Float sx, Float sy, Float s3, BSDFSample &r);
#if GDB_BYVAL & 8
GDB_CALL
#else
GDB_D
#endif
BSDFSample bsdfSampleV(const DMaterial &m, V3 wi, Float sx, Float sy, Float s3)
{
    BSDFSample r;
    const bool flipped = m.twosided && wi.z < 0;                                       // twosided.cpp:160-183
    if (flipped) wi.z *= -1;
    bsdfSampleOneSided(m, wi, sx, sy, s3, r);
    if (flipped && !isZero(r.weight) && r.pdf != 0) r.wo.z *= -1;
    return r;
}
#if GDB_BYVAL & 8
// s3: the NEXT value of the pixel's sample stream, peeked by the caller; a BSDF that draws from the sampler inside
// sample() (EUsesSampler: roughdielectric.cpp:524-531) consumes it and reports r.extraDraws = 1 so the caller advances.
GDB_D void bsdfSample(const DMaterial &m, V3 wi, Float sx, Float sy, Float s3, BSDFSample &r) { r = bsdfSampleV(m, wi, sx, sy, s3); }
#else
GDB_CALL void bsdfSample(const DMaterial &m, V3 wi, Float sx, Float sy, Float s3, BSDFSample &r) { r = bsdfSampleV(m, wi, sx, sy, s3); }
#endif
GDB_D void bsdfSampleOneSided(const DMaterial &m, V3 wi, Float sx, Float sy, Float s3, BSDFSample &r)
{
    r.weight = splat(0); r.pdf = 0; r.eta = 1.0; r.sampledType = 0; r.wo = mk(0, 0, 0); r.extraDraws = 0;
    switch (m.type) {
    case GDB200_BSDF_DIFFUSE:                                                          // diffuse.cpp:143-153
        if (wi.z <= 0) return;
        r.wo = squareToCosineHemisphere(sx, sy);
        r.sampledType = EDiffuseReflection;
        r.pdf = kInvPi * r.wo.z;
        r.weight = m.reflectance;
        return;
    case GDB200_BSDF_ROUGHCONDUCTOR: {                                                 // roughconductor.cpp:369-419
        if (wi.z < 0) return;
        const V3 mm = mfSampleVisible(m.distribution, m.alpha, wi, sx, sy);
        const Float temporaryPdf = mfPdfVisible(m.distribution, m.alpha, wi, mm);
        if (temporaryPdf == 0) return;
        r.wo = 2 * dot(wi, mm) * mm - wi;
        r.sampledType = EGlossyReflection;
        if (r.wo.z <= 0) return;
        const Spec F = fresnelConductorExact(dot(wi, mm), m.eta, m.k) * m.specR;
        const Float weight = mfSmithG1(m.distribution, m.alpha, r.wo, mm);
        if (weight > 0) { r.pdf = temporaryPdf / (4.0 * dot(r.wo, mm)); r.weight = F * weight; }
        return;
    }
    case GDB200_BSDF_CONDUCTOR:                                                        // conductor.cpp:268-285
        if (wi.z <= 0) return;
        r.sampledType = EDeltaReflection;
        r.wo = reflectLocal(wi);
        r.pdf = 1;
        r.weight = m.specR * fresnelConductorExact(wi.z, m.eta, m.k);
        return;
    case GDB200_BSDF_ROUGHDIELECTRIC: {                                                // roughdielectric.cpp:505-614 (both components, sampleVisible)
        const V3 wiS = wi * copysign(1.0, wi.z);                                       // math::signum = copysign(1, x), math.h:269-278
        const V3 mm = mfSampleVisible(m.distribution, m.alpha, wiS, sx, sy);
        const Float microfacetPDF = mfPdfVisible(m.distribution, m.alpha, wiS, mm);
        if (microfacetPDF == 0) return;
        float temporaryPdf = (float)microfacetPDF;                                     // a `float` in the reference (:533)
        Float cosThetaT;
        const Float F = fresnelDielectricExt(dot(wi, mm), cosThetaT, m.iorRatio);
        r.extraDraws = 1;                                                              // bRec.sampler->next1D()
        const bool sampleReflection = !(s3 > F);
        temporaryPdf = (float)((Float)temporaryPdf * (sampleReflection ? F : 1 - F));
        Spec weight = splat(1.0);
        Float dwh_dwo;
        if (sampleReflection) {
            r.wo = reflectAboutLocal(wi, mm); r.eta = 1.0; r.sampledType = EGlossyReflection;
            if (wi.z * r.wo.z <= 0) return;
            weight = weight * m.specR;
            dwh_dwo = 1.0 / (4.0 * dot(r.wo, mm));
        } else {
            if (cosThetaT == 0) return;
            const Float e = cosThetaT < 0 ? 1 / m.iorRatio : m.iorRatio;               // util.cpp:767-772
            r.wo = mm * (dot(wi, mm) * e + cosThetaT) - wi * e;
            r.eta = cosThetaT < 0 ? m.iorRatio : 1 / m.iorRatio; r.sampledType = EGlossyTransmission;
            if (wi.z * r.wo.z >= 0) return;
            const Float factor = cosThetaT < 0 ? 1 / m.iorRatio : m.iorRatio;
            weight = weight * (m.specT * (factor * factor));
            const Float sqrtDenom = dot(wi, mm) + r.eta * dot(r.wo, mm);
            dwh_dwo = (r.eta * r.eta * dot(r.wo, mm)) / (sqrtDenom * sqrtDenom);
        }
        weight = weight * mfSmithG1(m.distribution, m.alpha, r.wo, mm);
        temporaryPdf = (float)((Float)temporaryPdf * fabs(dwh_dwo));
        r.pdf = temporaryPdf; r.weight = weight;
        return;
    }
    case GDB200_BSDF_PLASTIC: {                                                        // plastic.cpp:372-414
        if (wi.z <= 0) return;
        Float dummy;
        const Float Fi = fresnelDielectricExt(wi.z, dummy, m.iorRatio);
        const Float probSpecular = (Fi * m.specSamplingWeight) / (Fi * m.specSamplingWeight + (1 - Fi) * (1 - m.specSamplingWeight));
        if (sx < probSpecular) {
            r.sampledType = EDeltaReflection; r.wo = reflectLocal(wi); r.pdf = probSpecular;
            r.weight = m.specR * Fi / probSpecular;
        } else {
            r.sampledType = EDiffuseReflection;
            r.wo = squareToCosineHemisphere((sx - probSpecular) / (1 - probSpecular), sy);
            const Float Fo = fresnelDielectricExt(r.wo.z, dummy, m.iorRatio);
            Spec diff = m.reflectance;
            if (m.nonlinear) diff = cdiv(diff, splat(1.0) - diff * m.fdrInt); else diff = diff / (1 - m.fdrInt);
            r.pdf = (1 - probSpecular) * (kInvPi * r.wo.z);
            r.weight = diff * (m.invEta2 * (1 - Fi) * (1 - Fo) / (1 - probSpecular));
        }
        return;
    }
    default: {                                                                         // dielectric.cpp:277-305
        Float cosThetaT;
        const Float F = fresnelDielectricExt(wi.z, cosThetaT, m.iorRatio);
        if (sx <= F) {
            r.sampledType = EDeltaReflection; r.wo = reflectLocal(wi); r.eta = 1.0; r.pdf = F;
            r.weight = m.specR;
        } else {
            r.sampledType = EDeltaTransmission; r.wo = refractLocal(m, wi, cosThetaT);
            r.eta = cosThetaT < 0 ? m.iorRatio : 1.0 / m.iorRatio; r.pdf = 1 - F;
            const Float factor = cosThetaT < 0 ? 1.0 / m.iorRatio : m.iorRatio;
            r.weight = m.specT * (factor * factor);
        }
        return;
    }
    }
}

GDB_D int vertexType(const DMaterial &m, unsigned bsdfTypeMask)                       // gpt.cpp:176-226, precomputed per material
{
    return (bsdfTypeMask & EDelta) ? m.vtDelta : m.vtSmooth;
}

// ---------------------------------------------------------------- emitters
struct DRec { V3 ref, refN, p, n, d; Float dist, pdf; int emitter; };

GDB_D Spec emittedLe(const Its &its, V3 d)                                            // area.cpp:104-109
{
    if (dot(its.sh.n, d) <= 0) return splat(0);
    return c_sceneG->emitters[its.emitter].radiance;
}
GDB_D void initDRec(const Its &ref, DRec &r)                                          // records.inl:160-165
{
    r.ref = ref.p;
    r.refN = c_sceneG->materials[ref.material].refNFromShading ? ref.sh.n : mk(0, 0, 0);
}

// DiscreteDistribution::sampleReuse (pmf.h:124-188): std::lower_bound over the n+1 CDF entries, zero-probability
// entries skipped, the sample rescaled into the chosen bin.
GDB_D int cdfSampleReuse(const Float *cdf, int n, Float &s, Float &pdf)
{
    int lo = 0, len = n + 1;                                                           // first entry with !(cdf[i] < s)
    while (len > 0) { const int half = len >> 1; if (cdf[lo + half] < s) { lo += half + 1; len -= half + 1; } else len = half; }
    int index = min(n - 1, max(0, lo - 1));
    while (cdf[index + 1] - cdf[index] == 0 && index < n) ++index;
    pdf = cdf[index + 1] - cdf[index];
    s = (s - cdf[index]) / (cdf[index + 1] - cdf[index]);
    return index;
}

// ---- environment map (envmap.cpp); lookups on the top MIP level, u repeats / v clamps (mipmap.h:503-566)
GDB_D Spec envTexel(int x, int y)
{
    const DEnv &e = c_scene.env;
    if (x < 0 || x >= e.width) { x = x % e.width; if (x < 0) x += e.width; }
    y = min(max(y, 0), e.height - 1);
    const Float *t = e.texels + ((size_t)y * e.width + x) * 3;
    return mk(t[0], t[1], t[2]);
}
GDB_HD Float luminance(Spec s) { return s.x * (Float)0.212671f + s.y * (Float)0.715160f + s.z * (Float)0.072169f; }   // spectrum.h:725-727
GDB_D Float safeAcos(Float v) { return acos(fmin(1.0, fmax(-1.0, v))); }
// EnvironmentMap::evalEnvironment without ray differentials (envmap.cpp:385-409, mipmap.h:575-596)
GDB_CALL Spec envEval(V3 dWorld)
{
    const DEnv &e = c_scene.env;
    const V3 v = xfVector(e.toObject, dWorld);
    const Float uvx = atan2(v.x, -v.z) * kInvTwoPi, uvy = safeAcos(v.y) * kInvPi;
    if (!isfinite(uvx) || !isfinite(uvy)) return splat(0);
    const Float u = uvx * e.width - (Float)0.5f, w = uvy * e.height - (Float)0.5f;
    const int xPos = (int)floor(u), yPos = (int)floor(w);
    const Float dx1 = u - xPos, dx2 = (Float)1.0f - dx1, dy1 = w - yPos, dy2 = (Float)1.0f - dy1;
    const Spec value = envTexel(xPos, yPos) * dx2 * dy2 + envTexel(xPos, yPos + 1) * dx2 * dy1
                     + envTexel(xPos + 1, yPos) * dx1 * dy2 + envTexel(xPos + 1, yPos + 1) * dx1 * dy1;
    return value * e.scale;
}
// envmap.cpp:611-642
GDB_D Float envPdfDirection(V3 d)
{
    const DEnv &e = c_scene.env;
    const Float uvx = atan2(d.x, -d.z) * kInvTwoPi, uvy = safeAcos(d.y) * kInvPi;
    if (!isfinite(uvx) || !isfinite(uvy)) return 0.0;
    const Float u = uvx * e.width - (Float)0.5f, v = uvy * e.height - (Float)0.5f;
    const int xPos = (int)floor(u), yPos = (int)floor(v);
    const Float dx1 = u - xPos, dx2 = (Float)1.0f - dx1, dy1 = v - yPos, dy2 = (Float)1.0f - dy1;
    const Spec value1 = envTexel(xPos, yPos) * dx2 * dy2 + envTexel(xPos + 1, yPos) * dx1 * dy2;
    const Spec value2 = envTexel(xPos, yPos + 1) * dx2 * dy1 + envTexel(xPos + 1, yPos + 1) * dx1 * dy1;
    const Float sinTheta = sqrt(fmax(0.0, 1 - d.y * d.y));
    return (luminance(value1) * e.rowWeights[min(max(yPos, 0), e.height - 1)] + luminance(value2) * e.rowWeights[min(max(yPos + 1, 0), e.height - 1)])
           * e.normalization / fmax(fabs(sinTheta), kEpsilon);
}
// sampleReuse over a float CDF (envmap.cpp:660-665): the comparison is made in single precision
GDB_D int envSampleReuse(const float *cdf, int size, Float &sample)
{
    const float key = (float)sample;
    int lo = 0, len = size + 1;
    while (len > 0) { const int half = len >> 1; if (cdf[lo + half] < key) { lo += half + 1; len -= half + 1; } else len = half; }
    const int index = min(max(0, lo - 1), size - 1);
    sample = (sample - (Float)cdf[index]) / (Float)(cdf[index + 1] - cdf[index]);
    return index;
}
GDB_D Float intervalToTent(Float sample)                                              // warp.cpp:143-155
{
    Float sign;
    if (sample < (Float)0.5f) { sign = 1; sample *= 2; } else { sign = -1; sample = 2 * (sample - (Float)0.5f); }
    return sign * (1 - sqrt(sample));
}
// envmap.cpp:571-608
GDB_D void envSampleDirection(Float sx, Float sy, V3 &d, Spec &value, Float &pdf)
{
    const DEnv &e = c_scene.env;
    const int row = envSampleReuse(e.cdfRows, e.height, sy);
    const int col = envSampleReuse(e.cdfCols + (size_t)row * (e.width + 1), e.width, sx);
    const Float posx = (Float)col + intervalToTent(sx), posy = (Float)row + intervalToTent(sy);
    const int xPos = (int)floor(posx), yPos = (int)floor(posy);
    const Float dx1 = posx - xPos, dx2 = (Float)1.0f - dx1, dy1 = posy - yPos, dy2 = (Float)1.0f - dy1;
    const Spec value1 = envTexel(xPos, yPos) * dx2 * dy2 + envTexel(xPos + 1, yPos) * dx1 * dy2;
    const Spec value2 = envTexel(xPos, yPos + 1) * dx2 * dy1 + envTexel(xPos + 1, yPos + 1) * dx1 * dy1;
    value = (value1 + value2) * e.scale;
    pdf = (luminance(value1) * e.rowWeights[min(max(yPos, 0), e.height - 1)] + luminance(value2) * e.rowWeights[min(max(yPos + 1, 0), e.height - 1)]) * e.normalization;
    const Float phi = e.pixelSizeX * (posx + (Float)0.5f), theta = e.pixelSizeY * (posy + (Float)0.5f);
    const Float sinPhi = sin(phi), cosPhi = cos(phi), sinTheta = sin(theta), cosTheta = cos(theta);
    d = mk(sinPhi * sinTheta, cosTheta, -cosPhi * sinTheta);
    pdf /= fmax(fabs(sinTheta), kEpsilon);
}
// BSphere::rayIntersect (bsphere.h:88-95)
GDB_D bool envSphereIntersect(V3 o, V3 d, Float &nearT, Float &farT)
{
    const V3 oc = o - c_scene.env.center;
    return solveQuadratic(len2(d), 2 * dot(oc, d), len2(oc) - c_scene.env.radius * c_scene.env.radius, nearT, farT);
}
// EnvironmentMap::fillDirectSamplingRecord (envmap.cpp:358-374)
GDB_D bool envFillDRec(DRec &dRec, V3 o, V3 d)
{
    Float nearT, farT;
    if (!envSphereIntersect(o, d, nearT, farT) || nearT > 0 || farT < 0) return false;
    dRec.p = o + d * farT;
    dRec.n = normalize(c_scene.env.center - dRec.p);
    dRec.d = d; dRec.dist = farT; dRec.emitter = c_scene.env.emitter;
    return true;
}

// Scene::sampleEmitterDirectVisible, scene.cpp:855-879: emitter pick (pmf.h:124-188), Emitter::sampleDirect
// (area.cpp:158-176 over shape.cpp:102-114 with rectangle.cpp:210-216 / trimesh.cpp:412-423 + triangle.cpp:24-50;
// envmap.cpp:516-544), then the shadow ray.
// sampleEmitterDirect: everything up to the shadow ray.  needsRay == false: the environment's early return below (the
// sample counts as visible and nothing is traced); otherwise the caller tests `shadowRay` and, if it is blocked, the
// sample is invisible with value 0 (scene.cpp:869-876).
GDB_D Spec sampleEmitterDirect(DRec &dRec, Float sx, Float sy, bool &needsRay, Ray &shadowRay)
{
    needsRay = false;
    Float emPdf;
    const int index = cdfSampleReuse(c_scene.emCdf, c_scene.nEmitters, sx, emPdf);
    const DEmitter &em = c_sceneG->emitters[index];
    Spec value;
    dRec.emitter = index;
    if (em.kind == EM_ENV) {
        V3 dl; Float pdf; Spec v;
        envSampleDirection(sx, sy, dl, v, pdf);
        const V3 dw = xfVector(c_scene.env.toWorld, dl);
        Float nearT = 0, farT = 0;
        if (isZero(v) || pdf == 0 || !envSphereIntersect(dRec.ref, dw, nearT, farT) || nearT >= 0 || farT <= 0) {
            // The reference leaves p/d/dist unset here and still traces its shadow ray with them; unreachable for
            // maps without black texels seen from inside the bounding sphere.  Defined as: no contribution.
            dRec.pdf = 0.0; dRec.p = dRec.ref; dRec.n = mk(0, 0, 0); dRec.d = mk(0, 0, 1); dRec.dist = 0;
            return splat(0);
        }
        dRec.pdf = pdf; dRec.p = dRec.ref + dw * farT; dRec.n = normalize(c_scene.env.center - dRec.p); dRec.dist = farT; dRec.d = dw;
        value = v / pdf;
    } else if (em.kind == EM_POINT) {                                                // point.cpp:131-147 (measure: discrete)
        dRec.p = em.position; dRec.n = mk(0, 0, 0);
        dRec.d = dRec.p - dRec.ref;
        dRec.dist = len(dRec.d);
        const Float invDist = (Float)1.0f / dRec.dist;
        dRec.d = dRec.d * invDist;
        dRec.pdf = 1;
        value = em.radiance * (invDist * invDist);
    } else if (em.kind == EM_SPOT) {                                                 // spot.cpp:184-200 with falloffCurve :105-125 (constant texture)
        dRec.p = em.position; dRec.n = mk(0, 0, 0);
        dRec.d = dRec.p - dRec.ref;
        dRec.dist = len(dRec.d);
        const Float invDist = (Float)1.0f / dRec.dist;
        dRec.d = dRec.d * invDist;
        dRec.pdf = 1;
        const V3 dw = -dRec.d;
        const Float cosTheta = em.toLocal[6] * dw.x + em.toLocal[7] * dw.y + em.toLocal[8] * dw.z;             // Frame::cosTheta of trafo.inverse()(-d), transform.h:180-181
        Float falloff = 1;
        if (cosTheta <= em.cosCutoff) falloff = 0;
        else if (cosTheta < em.cosBeam) falloff = (em.cutoffAngle - acos(cosTheta)) * em.invTransition;
        value = em.radiance * falloff * (invDist * invDist);
    } else if (em.kind == EM_SPHERE) {                                               // sphere.cpp:283-355 (Shirley et al. cone sampling), then area.cpp:158-176
        const DSphere &sp = c_sceneG->spheres[em.rect];
        const V3 refToCenter = sp.center - dRec.ref;
        const Float refDist2 = len2(refToCenter), invRefDist = (Float)1 / sqrt(refDist2);
        const Float sinAlpha = sp.radius * invRefDist;
        if (sinAlpha < 1 - kEpsilon) {
            const Float cosAlpha = sqrt(fmax(0.0, (Float)1.0f - sinAlpha * sinAlpha));
            const Float cosTheta = (1 - sx) + sx * cosAlpha, sinTheta = sqrt(fmax(0.0, (Float)1.0f - cosTheta * cosTheta));   // warp.cpp:54-63
            const Float phi = (Float)2.0f * kPi * sy, sinPhi = sin(phi), cosPhi = cos(phi);
            Frame fr; fr.n = refToCenter * invRefDist; coordinateSystem(fr.n, fr.s, fr.t);
            dRec.d = toWorld(fr, mk(cosPhi * sinTheta, sinPhi * sinTheta, cosTheta));
            dRec.pdf = kInvTwoPi / (1 - cosAlpha);
            const Float projDist = dot(refToCenter, dRec.d);
            const Float baseT = refDist2 / projDist;
            const V3 query = dRec.ref + dRec.d * baseT;
            const V3 queryToCenter = sp.center - query;
            const Float queryDist2 = len2(queryToCenter), queryProjDist = dot(queryToCenter, dRec.d);
            double nearT, farT;
            if (!solveQuadratic(1.0, -2 * queryProjDist, queryDist2 - sp.radius * sp.radius, nearT, farT)) nearT = queryProjDist;
            dRec.dist = baseT + nearT;
            dRec.n = normalize(dRec.d * nearT - queryToCenter);
            dRec.p = sp.center + dRec.n * sp.radius;
        } else {
            const Float z = (Float)1.0f - (Float)2.0f * sy, r = sqrt(fmax(0.0, (Float)1.0f - z * z));                       // warp.cpp:25-31
            const Float phi = (Float)2.0f * kPi * sx;
            const V3 d = mk(r * cos(phi), r * sin(phi), z);
            dRec.p = sp.center + d * sp.radius; dRec.n = d;
            dRec.d = dRec.p - dRec.ref;
            const Float dist2 = len2(dRec.d);
            dRec.dist = sqrt(dist2);
            dRec.d = dRec.d / dRec.dist;
            dRec.pdf = em.invArea * dist2 / fabs(dot(dRec.d, dRec.n));
        }
        if (sp.flip) dRec.n = dRec.n * -1.0;
        if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) value = em.radiance / dRec.pdf;   // area.cpp:158-176
        else { dRec.pdf = 0.0; value = splat(0); }
    } else {
        if (em.kind == EM_RECT) {
            const DRect &s = c_sceneG->rects[em.rect];
            dRec.p = xfAffine(s.toWorld, mk(sx * 2 - 1, sy * 2 - 1, 0));
            dRec.n = s.n;
            dRec.pdf = s.invArea;
        } else {                                                                     // trimesh.cpp:412-423
            Float triPdf;
            const int tri = cdfSampleReuse(c_scene.emTriCdf + em.cdfFirst, em.triCount, sy, triPdf);
            const DEmTri &T = c_scene.emTris[em.triFirst + tri];
            const Float a = sqrt(fmax(0.0, (Float)1.0f - sx));                           // warp.cpp:76-79
            const Float bx = 1 - a, by = a * sy;
            const V3 sideA = T.p1 - T.p0, sideB = T.p2 - T.p0;
            dRec.p = T.p0 + (sideA * bx) + (sideB * by);
            if (T.normals >= 0) {                                                     // triangle.cpp:33-42
                const V3 *vn = c_scene.triNormals + T.normals;
                dRec.n = normalize(vn[0] * ((Float)1.0f - bx - by) + vn[1] * bx + vn[2] * by);
            } else dRec.n = normalize(cross(sideA, sideB));
            dRec.pdf = em.invArea;
        }
        dRec.d = dRec.p - dRec.ref;                                                  // shape.cpp:102-114
        const Float distSquared = len2(dRec.d);
        dRec.dist = sqrt(distSquared);
        dRec.d = dRec.d / dRec.dist;
        const Float dp = fabs(dot(dRec.d, dRec.n));
        dRec.pdf *= dp != 0 ? (distSquared / dp) : 0.0;
        if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0 && dRec.pdf != 0) value = em.radiance / dRec.pdf;   // area.cpp:158-176
        else { dRec.pdf = 0.0; value = splat(0); }
    }
    dRec.pdf *= emPdf;
    value = value / emPdf;
    shadowRay.o = dRec.ref; shadowRay.d = dRec.d; shadowRay.mint = kEpsilon; shadowRay.maxt = dRec.dist * (1 - kShadowEpsilon);
    needsRay = true;
    return value;
}
GDB_D Spec sampleEmitterDirectVisibleImpl(DRec &dRec, Float sx, Float sy, bool &visible)
{
    bool needsRay; Ray ray;
    const Spec value = sampleEmitterDirect(dRec, sx, sy, needsRay, ray);
    visible = true;
    if (needsRay && rayOccluded(ray)) { visible = false; return splat(0); }
    return value;
}

struct EmitterSample { DRec dRec; Spec value; int visible; };
#if GDB_BYVAL & 16
GDB_CALL EmitterSample sampleEmitterDirectVisibleV(V3 ref, V3 refN, Float sx, Float sy)
{
    EmitterSample r; bool vis;
    r.dRec.ref = ref; r.dRec.refN = refN;
    r.value = sampleEmitterDirectVisibleImpl(r.dRec, sx, sy, vis);
    r.visible = vis;
    return r;
}
GDB_D Spec sampleEmitterDirectVisible(DRec &dRec, Float sx, Float sy, bool &visible)
{
    const EmitterSample r = sampleEmitterDirectVisibleV(dRec.ref, dRec.refN, sx, sy);
    dRec = r.dRec; visible = r.visible != 0;
    return r.value;
}
#else
GDB_CALL Spec sampleEmitterDirectVisible(DRec &dRec, Float sx, Float sy, bool &visible) { return sampleEmitterDirectVisibleImpl(dRec, sx, sy, visible); }
#endif

// Scene::pdfEmitterDirect, scene.cpp:976-979 + area.cpp:178-186 + shape.cpp:116-126 / envmap.cpp:546-556
GDB_D Float pdfEmitterDirect(const DRec &dRec)
{
    const DEmitter &em = c_sceneG->emitters[dRec.emitter];
    Float pdf = 0.0;
    if (em.kind == EM_ENV) pdf = envPdfDirection(xfVector(c_scene.env.toObject, dRec.d));
    else if (em.kind == EM_POINT || em.kind == EM_SPOT) pdf = 0.0;                   // point.cpp:149-151, spot.cpp:202-204 for a solid-angle query
    else if (em.kind == EM_SPHERE) {                                                 // area.cpp:178-186 + sphere.cpp:357-387
        if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0) {
            const DSphere &sp = c_sceneG->spheres[em.rect];
            const Float invRefDist = (Float)1.0f / len(sp.center - dRec.ref), sinAlpha = sp.radius * invRefDist;
            if (sinAlpha < 1 - kEpsilon) pdf = kInvTwoPi / (1 - sqrt(fmax(0.0, 1 - sinAlpha * sinAlpha)));
            else pdf = em.invArea * dRec.dist * dRec.dist / fabs(dot(dRec.d, dRec.n));
        }
    }
    else if (dot(dRec.d, dRec.refN) >= 0 && dot(dRec.d, dRec.n) < 0) {
        const Float invArea = em.kind == EM_RECT ? c_sceneG->rects[em.rect].invArea : em.invArea;
        pdf = invArea * (dRec.dist * dRec.dist) / fabs(dot(dRec.d, dRec.n));
    }
    return pdf * em.pdfDiscrete;
}

// ---------------------------------------------------------------- shifts
GDB_D V3 reflectAbout(V3 wi, V3 n) { return 2 * dot(wi, n) * n - wi; }                // util.cpp:763-765
GDB_D V3 refractAbout(V3 wi, V3 n, Float eta)                                         // util.cpp:774-792
{
    if (eta == 1) return -wi;
    const Float cosThetaI = dot(wi, n);
    if (cosThetaI > 0) eta = 1 / eta;
    const Float cosThetaTSqr = 1 - (1 - cosThetaI * cosThetaI) * (eta * eta);
    if (cosThetaTSqr <= 0.0) return mk(0, 0, 0);
    return n * (cosThetaI * eta - copysign(1.0, cosThetaI) * sqrt(cosThetaTSqr)) - wi * eta;
}

struct ShiftResult { bool success; Float jacobian; V3 wo; };

GDB_D ShiftResult halfVectorShift(V3 mainWi, V3 mainWo, V3 shiftedWi, Float mainEta, Float shiftedEta)   // gpt.cpp:242-305
{
    ShiftResult result; result.success = false; result.jacobian = 0; result.wo = mk(0, 0, 0);
    if (mainWi.z * mainWo.z < 0) {
        if (mainEta == 1 || shiftedEta == 1) return result;
        const V3 hMain = (mainWi.z < 0) ? -(mainWi * mainEta + mainWo) : -(mainWi + mainWo * mainEta);
        const V3 h = normalize(hMain);
        const V3 shiftedWo = refractAbout(shiftedWi, h, shiftedEta);
        if (isZero(shiftedWo)) return result;
        const V3 hShifted = (shiftedWi.z < 0) ? -(shiftedWi * shiftedEta + shiftedWo) : -(shiftedWi + shiftedWo * shiftedEta);
        const Float hLengthSquared = len2(hShifted) / (kDEps + len2(hMain));
        const Float WoDotH = fabs(dot(mainWo, h)) / (kDEps + fabs(dot(shiftedWo, h)));
        result.success = true; result.wo = shiftedWo; result.jacobian = hLengthSquared * WoDotH;
    } else {
        const V3 h = normalize(mainWi + mainWo);
        const V3 shiftedWo = reflectAbout(shiftedWi, h);
        const Float WoDotH = dot(shiftedWo, h) / dot(mainWo, h);
        result.success = true; result.wo = shiftedWo; result.jacobian = fabs(WoDotH);
    }
    return result;
}

// The reconnection as it comes out when `r` is unobstructed (success = true); the caller tests r (testVisibility, gpt.cpp:84-93).
GDB_D ShiftResult reconnectShiftUnoccluded(V3 mainSource, V3 target, V3 shiftSource, V3 targetNormal, Ray &r)   // gpt.cpp:316-345
{
    ShiftResult result; result.success = false; result.jacobian = 0; result.wo = mk(0, 0, 0);
    r.o = shiftSource; r.d = target - shiftSource; r.mint = kEpsilon; r.maxt = 1.0 - kShadowEpsilon;
    const V3 mainEdge = mainSource - target, shiftedEdge = shiftSource - target;
    const Float mainL2 = len2(mainEdge), shiftedL2 = len2(shiftedEdge);
    const V3 shiftedWo = -shiftedEdge / sqrt(shiftedL2);
    const Float mainOpposingCosine = dot(mainEdge, targetNormal) / sqrt(mainL2);
    const Float shiftedOpposingCosine = dot(shiftedWo, targetNormal);
    result.jacobian = fabs(shiftedOpposingCosine * mainL2) / (kDEps + fabs(mainOpposingCosine * shiftedL2));
    result.success = true; result.wo = shiftedWo;
    return result;
}
GDB_D ShiftResult reconnectShift(V3 mainSource, V3 target, V3 shiftSource, V3 targetNormal)   // gpt.cpp:84-93, 316-345
{
    Ray r;
    ShiftResult result = reconnectShiftUnoccluded(mainSource, target, shiftSource, targetNormal, r);
    if (rayOccluded(r)) { result.success = false; result.jacobian = 0; result.wo = mk(0, 0, 0); }
    return result;
}

// environmentShift + testEnvironmentVisibility (gpt.cpp:96-114, 348-369): the offset vertex must see the environment
// in the base path's direction; J = 1.
GDB_D ShiftResult environmentShiftUnoccluded(V3 mainD, V3 shiftSource, Ray &r)
{
    ShiftResult result;
    DRec dr; dr.dist = 0;
    envFillDRec(dr, shiftSource, mainD);
    r.o = shiftSource; r.d = mainD; r.mint = kEpsilon; r.maxt = (1.0 - kShadowEpsilon) * dr.dist;
    result.success = true; result.jacobian = 1; result.wo = mainD;
    return result;
}
GDB_D ShiftResult environmentShift(V3 mainD, V3 shiftSource)
{
    Ray r;
    ShiftResult result = environmentShiftUnoccluded(mainD, shiftSource, r);
    if (rayOccluded(r)) { result.success = false; result.jacobian = 0; result.wo = mk(0, 0, 0); }
    return result;
}

// perspective.cpp:271-298; with an aperture (apertureRadius > 0) thinlens.cpp:289-318
GDB_D void sampleCameraRay(Float px, Float py, Float ax, Float ay, Ray &ray)
{
    const V3 nearP = xfPoint(c_scene.sampleToCamera, mk(px * c_scene.invResX, py * c_scene.invResY, 0.0));
    if (c_scene.apertureRadius > 0) {
        Float tx, ty;
        squareToUniformDiskConcentric(ax, ay, tx, ty);
        const V3 apertureP = mk(tx * c_scene.apertureRadius, ty * c_scene.apertureRadius, 0.0);
        const V3 focusP = nearP * (c_scene.focusDistance / nearP.z);
        const V3 d = normalize(focusP - apertureP);
        const Float invZ = 1.0 / d.z;
        ray.mint = c_scene.nearClip * invZ;
        ray.maxt = c_scene.farClip * invZ;
        ray.o = xfAffine(c_scene.cameraToWorld, apertureP);
        ray.d = xfVector(c_scene.cameraToWorld, d);
        return;
    }
    const V3 d = normalize(nearP);
    const Float invZ = 1.0 / d.z;
    ray.mint = c_scene.nearClip * invZ;
    ray.maxt = c_scene.farClip * invZ;
    ray.o = xfAffine(c_scene.cameraToWorld, mk(0, 0, 0));
    ray.d = xfVector(c_scene.cameraToWorld, d);
}

// ---------------------------------------------------------------- sampler (gdb200_counter)
GDB_HD uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
GDB_HD uint64_t samplerKey(uint64_t seed, int px, int py)
{
    return mix64(mix64(seed + 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)(uint32_t)px | ((uint64_t)(uint32_t)py << 32)));
}
struct Sampler {
    uint64_t key; uint32_t n;
    GDB_D Float next1D()
    {
        n++;
        return (Float)(mix64(key + (uint64_t)n * 0x9E3779B97F4A7C15ULL) >> 11) * (1.0 / 9007199254740992.0);
    }
    GDB_D Float peek1D() const     // the value next1D() would return, without consuming it
    {
        return (Float)(mix64(key + (uint64_t)(n + 1) * 0x9E3779B97F4A7C15ULL) >> 11) * (1.0 / 9007199254740992.0);
    }
};

}  // namespace gdb200
