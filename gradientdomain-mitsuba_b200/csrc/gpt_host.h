// gdb200 G-PT tracer — host-side flattening of a gdb200_scene_desc into the device tables of
// gpt_device.cuh (pure C++: no CUDA runtime calls; gpt.cu uploads the result).  Requires
// gdb200::set_error to be declared by the including translation unit (common.h in the library).
#pragma once
#include "gpt_kernels.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace gdb200 {

struct HostScene {
    DScene host;                       // flattened tables (vertex classification filled per render); device pointers are patched in by the uploader
    DBounds bounds[kMaxPrims];         // padded per-primitive bounds (candidate selection)
    std::vector<gdb200_material> mats;
    int width = 0, height = 0;
    // variable-size tables (global memory on the device)
    std::vector<Float> envTexels, envRowWeights, emTriCdf;
    std::vector<float> envCdfRows, envCdfCols;
    std::vector<DEmTri> emTris;
    std::vector<V3> triNormals;
    std::vector<BvhNode> bvh;
    std::vector<DTri> bvhTris;
};

// TriAccel::load (triaccel.h:61-95) + the flat face normal of skdtree.h:367-371.  Returns false for a degenerate
// triangle (k = 3: never hit, triaccel.h:75-78).
inline bool makeTri(V3 A, V3 B, V3 C, int material, int emitter, DTri &T)
{
    memset(&T, 0, sizeof(T));
    T.normals = -1;
    static const int waldModulo[4] = {1, 2, 0, 1};
    const V3 b = C - A, cc = B - A, N = cross(cc, b);
    const double Nv[3] = {N.x, N.y, N.z}, bv[3] = {b.x, b.y, b.z}, cv[3] = {cc.x, cc.y, cc.z}, Av[3] = {A.x, A.y, A.z};
    int k = 0;
    for (int j = 0; j < 3; j++) if (std::abs(Nv[j]) > std::abs(Nv[k])) k = j;
    const int u = waldModulo[k], v = waldModulo[k + 1];
    const double n_k = Nv[k], denom = bv[u] * cv[v] - bv[v] * cv[u];
    T.p0 = A; T.p1 = B; T.p2 = C; T.material = material; T.emitter = emitter;
    if (denom == 0) { T.k = 3; return false; }
    T.k = k;
    T.n_u = Nv[u] / n_k; T.n_v = Nv[v] / n_k; T.n_d = dot(A, N) / n_k;
    T.b_nu = bv[u] / denom; T.b_nv = -bv[v] / denom; T.a_u = Av[u]; T.a_v = Av[v];
    T.c_nu = cv[v] / denom; T.c_nv = -cv[u] / denom;
    V3 faceNormal = cross(B - A, C - A);
    const double l = len(faceNormal);
    if (!isZero(faceNormal)) faceNormal = faceNormal / l;
    T.faceNormal = faceNormal;
    return true;
}

// BVH2 over `tris` (reordered in place into leaf order): binned-SAH splits on the centroid bounds, leaves of <= 4
// triangles, node bounds rounded outwards to single precision and padded by `pad` (conservative for the fp32 slab test).
inline void buildBvh(std::vector<DTri> &tris, std::vector<BvhNode> &nodes, double pad)
{
    const int n = (int)tris.size();
    struct Box { double lo[3], hi[3]; };
    auto emptyBox = [] { Box b; for (int k = 0; k < 3; k++) { b.lo[k] = std::numeric_limits<double>::infinity(); b.hi[k] = -b.lo[k]; } return b; };
    auto growP = [](Box &b, const double *p) { for (int k = 0; k < 3; k++) { b.lo[k] = std::min(b.lo[k], p[k]); b.hi[k] = std::max(b.hi[k], p[k]); } };
    auto growB = [](Box &b, const Box &o) { for (int k = 0; k < 3; k++) { b.lo[k] = std::min(b.lo[k], o.lo[k]); b.hi[k] = std::max(b.hi[k], o.hi[k]); } };
    auto area = [](const Box &b) { const double x = b.hi[0] - b.lo[0], y = b.hi[1] - b.lo[1], z = b.hi[2] - b.lo[2]; return x < 0 ? 0.0 : 2 * (x * y + y * z + z * x); };
    std::vector<Box> tb(n); std::vector<double> cen((size_t)3 * n); std::vector<int> order(n);
    for (int i = 0; i < n; i++) {
        tb[i] = emptyBox();
        const V3 P[3] = {tris[i].p0, tris[i].p1, tris[i].p2};
        for (const V3 &q : P) { const double v[3] = {q.x, q.y, q.z}; growP(tb[i], v); }
        for (int k = 0; k < 3; k++) cen[3 * (size_t)i + k] = 0.5 * (tb[i].lo[k] + tb[i].hi[k]);
        order[i] = i;
    }
    nodes.clear();
    struct Task { int node, first, count; };
    std::vector<Task> todo;
    auto setBounds = [&](BvhNode &N, const Box &b) {
        for (int k = 0; k < 3; k++) {
            N.lo[k] = std::nextafterf((float)(b.lo[k] - pad), -std::numeric_limits<float>::infinity());
            N.hi[k] = std::nextafterf((float)(b.hi[k] + pad), std::numeric_limits<float>::infinity());
        }
    };
    nodes.push_back(BvhNode());
    todo.push_back({0, 0, n});
    while (!todo.empty()) {
        const Task t = todo.back(); todo.pop_back();
        Box bb = emptyBox(), cb = emptyBox();
        for (int i = t.first; i < t.first + t.count; i++) { growB(bb, tb[order[i]]); growP(cb, &cen[3 * (size_t)order[i]]); }
        setBounds(nodes[t.node], bb);
        if (t.count <= 4) { nodes[t.node].a = t.first; nodes[t.node].b = -t.count; continue; }
        int axis = 0;
        for (int k = 1; k < 3; k++) if (cb.hi[k] - cb.lo[k] > cb.hi[axis] - cb.lo[axis]) axis = k;
        int mid = t.first + t.count / 2;
        const double ext = cb.hi[axis] - cb.lo[axis];
        bool split = false;
        if (ext > 0) {
            const int NB = 16;
            Box binBox[NB]; int binCnt[NB];
            for (int b = 0; b < NB; b++) { binBox[b] = emptyBox(); binCnt[b] = 0; }
            auto binOf = [&](int tri) { return std::min(NB - 1, (int)((cen[3 * (size_t)tri + axis] - cb.lo[axis]) / ext * NB)); };
            for (int i = t.first; i < t.first + t.count; i++) { const int b = binOf(order[i]); growB(binBox[b], tb[order[i]]); binCnt[b]++; }
            double best = std::numeric_limits<double>::infinity(); int bestSplit = -1;
            Box rightAcc[NB]; int rightCnt[NB]; Box acc = emptyBox(); int cnt = 0;
            for (int b = NB - 1; b > 0; b--) { growB(acc, binBox[b]); cnt += binCnt[b]; rightAcc[b] = acc; rightCnt[b] = cnt; }
            acc = emptyBox(); cnt = 0;
            for (int b = 0; b < NB - 1; b++) {
                growB(acc, binBox[b]); cnt += binCnt[b];
                if (cnt == 0 || rightCnt[b + 1] == 0) continue;
                const double cost = area(acc) * cnt + area(rightAcc[b + 1]) * rightCnt[b + 1];
                if (cost < best) { best = cost; bestSplit = b; }
            }
            if (bestSplit >= 0) {
                int *beg = order.data() + t.first, *end = beg + t.count;
                int *m = std::partition(beg, end, [&](int tri) { return binOf(tri) <= bestSplit; });
                mid = (int)(m - order.data());
                split = mid > t.first && mid < t.first + t.count;
            }
        }
        if (!split) {      // all centroids coincide (or SAH found nothing): split the list in half
            mid = t.first + t.count / 2;
            std::nth_element(order.begin() + t.first, order.begin() + mid, order.begin() + t.first + t.count,
                             [&](int x, int y) { return cen[3 * (size_t)x + axis] < cen[3 * (size_t)y + axis]; });
        }
        const int left = (int)nodes.size();
        nodes.push_back(BvhNode()); nodes.push_back(BvhNode());
        nodes[t.node].a = left; nodes[t.node].b = left + 1;
        todo.push_back({left, t.first, mid - t.first});
        todo.push_back({left + 1, mid, t.first + t.count - mid});
    }
    std::vector<DTri> sorted(n);
    for (int i = 0; i < n; i++) sorted[i] = tris[order[i]];
    tris.swap(sorted);
}

// EnvironmentMap::configure (envmap.cpp:263-320): marginal / conditional CDFs over luminance * sin(theta), in the
// reference's mixed precision (float tables, Float sums).
inline int buildEnvTables(const gdb200_envmap *e, int emitterIndex, HostScene *s)
{
    DEnv &o = s->host.env;
    if (e->width < 1 || e->height < 1 || !e->rgb) return set_error(GDB200_ERR_ARGUMENT, "envmap: empty image");
    if (std::max(e->width, e->height) > 0xFFFF) return set_error(GDB200_ERR_ARGUMENT, "Environment maps images must be smaller than 65536 pixels in width and height");
    const int W = e->width, H = e->height;
    o.present = 1; o.width = W; o.height = H; o.emitter = emitterIndex;
    if (e->to_world[12] != 0 || e->to_world[13] != 0 || e->to_world[14] != 0 || e->to_world[15] != 1) return set_error(GDB200_ERR_ARGUMENT, "envmap: toWorld must be affine");
    memcpy(o.toWorld, e->to_world, sizeof(o.toWorld)); memcpy(o.toObject, e->to_object, sizeof(o.toObject));
    o.center = mk(e->bsphere_center[0], e->bsphere_center[1], e->bsphere_center[2]); o.radius = e->bsphere_radius; o.scale = e->scale;
    s->envTexels.resize((size_t)W * H * 3);
    for (size_t i = 0; i < s->envTexels.size(); i++) s->envTexels[i] = (Float)e->rgb[i];
    s->envCdfCols.assign((size_t)(W + 1) * H, 0.f); s->envCdfRows.assign(H + 1, 0.f); s->envRowWeights.assign(H, 0.0);
    size_t colPos = 0, rowPos = 0;
    Float rowSum = 0.0;
    s->envCdfRows[rowPos++] = 0;
    for (int y = 0; y < H; ++y) {
        Float colSum = 0;
        s->envCdfCols[colPos++] = 0;
        for (int x = 0; x < W; ++x) {
            const Float *t = &s->envTexels[((size_t)y * W + x) * 3];
            colSum += t[0] * (Float)0.212671f + t[1] * (Float)0.715160f + t[2] * (Float)0.072169f;
            s->envCdfCols[colPos++] = (float)colSum;
        }
        const float normalization = 1.0f / (float)colSum;
        for (int x = 1; x < W; ++x) s->envCdfCols[colPos - x - 1] *= normalization;
        s->envCdfCols[colPos - 1] = 1.0f;
        const Float weight = std::sin((y + (Float)0.5f) * kPi / H);
        s->envRowWeights[y] = weight;
        rowSum += colSum * weight;
        s->envCdfRows[rowPos++] = (float)rowSum;
    }
    const float normalization = 1.0f / (float)rowSum;
    for (int y = 1; y < H; ++y) s->envCdfRows[rowPos - y - 1] *= normalization;
    s->envCdfRows[rowPos - 1] = 1.0f;
    if (rowSum == 0) return set_error(GDB200_ERR_ARGUMENT, "The environment map is completely black -- this is not allowed.");
    if (!std::isfinite(rowSum)) return set_error(GDB200_ERR_ARGUMENT, "The environment map contains an invalid floating point value (nan/inf) -- giving up.");
    o.normalization = (Float)1.0f / (rowSum * (2 * kPi / W) * (kPi / H));
    o.pixelSizeX = 2 * kPi / W; o.pixelSizeY = kPi / H;
    return GDB200_OK;
}

inline int flattenScene(const gdb200_scene_desc *d, HostScene *s)
{
    DScene &h = s->host;
    memset(&h, 0, sizeof(h));
    s->envTexels.clear(); s->envRowWeights.clear(); s->emTriCdf.clear(); s->envCdfRows.clear(); s->envCdfCols.clear();
    s->emTris.clear(); s->bvh.clear(); s->bvhTris.clear(); s->triNormals.clear();
    const gdb200_camera &c = d->camera;
    if (c.width <= 0 || c.height <= 0) return set_error(GDB200_ERR_ARGUMENT, "invalid film size %dx%d", c.width, c.height);
    if ((long long)c.width * c.height > (1LL << 28)) return set_error(GDB200_ERR_ARGUMENT, "film size %dx%d exceeds 2^28 pixels", c.width, c.height);   // 5*n*4 accumulators are indexed with int products elsewhere
    if (d->n_shapes < 0 || d->n_materials < 1 || d->n_vertices < 0 || d->n_triangles < 0)
        return set_error(GDB200_ERR_ARGUMENT, "invalid scene counts (%d shapes, %d materials, %d vertices, %d triangles)", d->n_shapes, d->n_materials, d->n_vertices, d->n_triangles);
    if ((d->n_shapes && !d->shapes) || !d->materials || !d->emitters || (d->n_triangles && (!d->triangles || !d->vertices)))
        return set_error(GDB200_ERR_ARGUMENT, "scene description has a NULL table");
    if (d->n_emitters < 1) return set_error(GDB200_ERR_ARGUMENT, "scene has no emitter");
    if (d->n_materials > kMaxMaterials || d->n_emitters > kMaxEmitters)
        return set_error(GDB200_ERR_ARGUMENT, "too many materials/emitters (%d/%d, limits %d/%d)", d->n_materials, d->n_emitters, kMaxMaterials, kMaxEmitters);
    memcpy(h.sampleToCamera, c.sample_to_camera, sizeof(h.sampleToCamera));
    memcpy(h.cameraToWorld, c.camera_to_world, sizeof(h.cameraToWorld));
    h.nearClip = c.near_clip; h.farClip = c.far_clip; h.width = c.width; h.height = c.height;
    if (c.aperture_radius < 0 || (c.aperture_radius > 0 && !(c.focus_distance > 0))) return set_error(GDB200_ERR_ARGUMENT, "invalid aperture radius / focus distance");
    h.apertureRadius = c.aperture_radius; h.focusDistance = c.focus_distance;
    h.invResX = 1.0 / c.width; h.invResY = 1.0 / c.height;
    if (!(d->rfilter_radius > 0)) return set_error(GDB200_ERR_ARGUMENT, "invalid reconstruction filter radius %g", d->rfilter_radius);
    h.filterRadius = d->rfilter_radius; h.filterScale = 31 / d->rfilter_radius;
    bool boxFilter = true;
    for (int i = 0; i < 32; i++) { h.filterTable[i] = d->rfilter_table[i]; if (d->rfilter_table[i] != 0) boxFilter = false; }
    if (boxFilter) { for (int i = 0; i < 31; i++) h.filterTable[i] = 1.0 / (2 * d->rfilter_radius); h.filterTable[31] = 0; }   // box.cpp:45-47 through rfilter.cpp:37-55
    h.filterIsBox = h.filterTable[31] == 0;                 // also when the caller passed the box filter's own table
    for (int i = 1; i < 31; i++) if (h.filterTable[i] != h.filterTable[0]) h.filterIsBox = 0;
    s->width = c.width; s->height = c.height;
    s->mats.assign(d->materials, d->materials + d->n_materials);
    h.nMaterials = d->n_materials;
    for (int i = 0; i < d->n_materials; i++) {
        const gdb200_material &m = d->materials[i];
        static_assert(kBsdfTypes == GDB200_BSDF_ROUGHDIELECTRIC + 1, "the compaction queues are keyed by BSDF type");
        if (m.type < GDB200_BSDF_DIFFUSE || m.type >= kBsdfTypes) return set_error(GDB200_ERR_ARGUMENT, "material %d: unknown BSDF type %d", i, m.type);
        if (m.twosided && (m.type == GDB200_BSDF_DIELECTRIC || m.type == GDB200_BSDF_ROUGHDIELECTRIC))
            return set_error(GDB200_ERR_ARGUMENT, "material %d: Only materials without a transmission component can be nested!", i);   // twosided.cpp:103-105
    }
    // Triangles of all meshes: in the constant-memory table while they fit, else (or with GDB200_FORCE_BVH) behind a BVH.
    long long meshTriTotal = 0;
    for (int i = 0; i < d->n_shapes; i++) if (d->shapes[i].type == GDB200_SHAPE_MESH) meshTriTotal += d->shapes[i].tri_count;
    const bool useBvh = meshTriTotal > kMaxTris || getenv("GDB200_FORCE_BVH") != nullptr;
    double scale = 0;                                   // scene scale for the padding of the fp32 bounds
    for (int k = 0; k < 3; k++) scale = std::max(scale, std::abs(c.camera_to_world[4 * k + 3]));
    std::vector<int> rectOfShape(d->n_shapes, -1), emTriFirstOfShape(d->n_shapes, -1), sphereOfShape(d->n_shapes, -1);
    std::vector<double> areaOfShape(d->n_shapes, 0.0);
    for (int i = 0; i < d->n_shapes; i++) {
        const gdb200_shape &sh = d->shapes[i];
        if (sh.material < 0 || sh.material >= d->n_materials) return set_error(GDB200_ERR_ARGUMENT, "shape %d: bad material index", i);
        if (sh.emitter >= d->n_emitters) return set_error(GDB200_ERR_ARGUMENT, "shape %d: bad emitter index", i);
        if (sh.emitter >= 0 && (d->emitters[sh.emitter].type != GDB200_EMITTER_AREA || d->emitters[sh.emitter].shape != i))
            return set_error(GDB200_ERR_ARGUMENT, "shape %d: emitter %d belongs to another shape", i, sh.emitter);
        if (sh.type == GDB200_SHAPE_RECTANGLE) {                                     // rectangle.cpp:100-110
            if (h.nRects >= kMaxRects) return set_error(GDB200_ERR_ARGUMENT, "too many rectangles (limit %d)", kMaxRects);
            if (sh.to_world[12] != 0 || sh.to_world[13] != 0 || sh.to_world[14] != 0 || sh.to_world[15] != 1)
                return set_error(GDB200_ERR_ARGUMENT, "shape %d: toWorld must be affine", i);
            DRect &r = h.rects[h.nRects];
            memcpy(r.toObject, sh.to_object, sizeof(r.toObject)); memcpy(r.toWorld, sh.to_world, sizeof(r.toWorld));
            r.dpdu = xfVector(sh.to_world, mk(2, 0, 0));
            const V3 dpdv = xfVector(sh.to_world, mk(0, 2, 0));
            r.n = normalize(xfNormal(sh.to_object, mk(0, 0, 1)));
            r.invArea = 1.0 / (len(r.dpdu) * len(dpdv));
            if (std::fabs(dot(normalize(r.dpdu), normalize(dpdv))) > kEpsilon)           // rectangle.cpp:108-109
                return set_error(GDB200_ERR_ARGUMENT, "shape %d: Error: 'toWorld' transformation contains shear!", i);
            r.material = sh.material; r.emitter = sh.emitter;
            rectOfShape[i] = h.nRects++;
        } else if (sh.type == GDB200_SHAPE_SPHERE) {
            if (h.nSpheres >= kMaxSpheres) return set_error(GDB200_ERR_ARGUMENT, "too many spheres (limit %d)", kMaxSpheres);
            if (!(sh.radius > 0)) return set_error(GDB200_ERR_ARGUMENT, "shape %d: sphere radius must be positive", i);
            sphereOfShape[i] = h.nSpheres;
            DSphere &sp = h.spheres[h.nSpheres++];
            sp.center = mk(sh.center[0], sh.center[1], sh.center[2]); sp.radius = sh.radius; sp.flip = sh.flip_normals;
            sp.material = sh.material; sp.emitter = sh.emitter;
        } else if (sh.type == GDB200_SHAPE_MESH) {
            if (sh.first_tri < 0 || sh.tri_count < 0 || sh.first_tri + (long long)sh.tri_count > d->n_triangles)
                return set_error(GDB200_ERR_ARGUMENT, "shape %d: triangle range out of bounds", i);
            if (!useBvh && h.nMeshes >= kMaxMeshes) return set_error(GDB200_ERR_ARGUMENT, "too many meshes (limit %d)", kMaxMeshes);
            std::vector<DTri> meshTris;
            const double big = std::numeric_limits<double>::infinity();
            V3 lo = mk(big, big, big), hi = mk(-big, -big, -big);
            if (sh.emitter >= 0) emTriFirstOfShape[i] = (int)s->emTris.size();
            for (int t = sh.first_tri; t < sh.first_tri + sh.tri_count; t++) {
                const int *ix = d->triangles + 3 * t;
                for (int k = 0; k < 3; k++) if (ix[k] < 0 || ix[k] >= d->n_vertices) return set_error(GDB200_ERR_ARGUMENT, "shape %d: vertex index out of bounds", i);
                const double *va = d->vertices + 3 * ix[0], *vb = d->vertices + 3 * ix[1], *vc = d->vertices + 3 * ix[2];
                const V3 A = mk(va[0], va[1], va[2]), B = mk(vb[0], vb[1], vb[2]), C = mk(vc[0], vc[1], vc[2]);
                for (const V3 &P : {A, B, C}) {
                    lo = mk(std::min(lo.x, P.x), std::min(lo.y, P.y), std::min(lo.z, P.z));
                    hi = mk(std::max(hi.x, P.x), std::max(hi.y, P.y), std::max(hi.z, P.z));
                    scale = std::max(scale, std::max(std::abs(P.x), std::max(std::abs(P.y), std::abs(P.z))));
                }
                int normalIndex = -1;
                if (sh.has_vertex_normals) {
                    if (!d->normals) return set_error(GDB200_ERR_ARGUMENT, "shape %d: has_vertex_normals without gdb200_scene_desc.normals", i);
                    normalIndex = (int)s->triNormals.size();
                    for (int k = 0; k < 3; k++) { const double *vn = d->normals + 3 * ix[k]; s->triNormals.push_back(mk(vn[0], vn[1], vn[2])); }
                }
                if (sh.emitter >= 0) {                                               // trimesh.cpp:388-403, triangle.cpp:61-67
                    DEmTri E; E.p0 = A; E.p1 = B; E.p2 = C; E.normals = normalIndex; E.pad = 0;
                    s->emTris.push_back(E);
                    areaOfShape[i] += (Float)0.5f * len(cross(B - A, C - A));
                }
                DTri T;
                if (!makeTri(A, B, C, sh.material, sh.emitter, T)) continue;
                T.normals = normalIndex;
                meshTris.push_back(T);
            }
            if (useBvh) s->bvhTris.insert(s->bvhTris.end(), meshTris.begin(), meshTris.end());
            else {
                if (h.nTris + (int)meshTris.size() > kMaxTris) return set_error(GDB200_ERR_ARGUMENT, "too many triangles for the constant-memory scene table (limit %d)", kMaxTris);
                DMesh &M = h.meshes[h.nMeshes++];
                M.first = h.nTris; M.lo = lo; M.hi = hi;
                for (int k = 0; k < 3; k++) {          // store grouped by projection axis (order inside a group is kept)
                    for (const DTri &T : meshTris) if (T.k == k) h.tris[h.nTris++] = T;
                    M.kEnd[k] = h.nTris;
                }
                M.count = h.nTris - M.first;
            }
        } else return set_error(GDB200_ERR_ARGUMENT, "shape %d: unknown type %d", i, sh.type);
    }
    for (int mi = 0; mi < h.nMeshes; mi++) {   // enlarge the skip-bounds far beyond any rounding of the slab test
        DMesh &M = h.meshes[mi];
        const V3 ext = M.hi - M.lo;
        const double pad = 1e-6 * std::max(1.0, std::max(ext.x, std::max(ext.y, ext.z))) + 1e-9 * std::max(maxComp(M.hi), -std::min(M.lo.x, std::min(M.lo.y, M.lo.z)));
        M.lo = M.lo - splat(pad); M.hi = M.hi + splat(pad);
    }
    // padded bounds of every table primitive for the candidate pass of closestPrimitive
    {
        int np = 0;
        auto grow = [&](DBounds &B, V3 P) {
            const double v[3] = {P.x, P.y, P.z};
            for (int k = 0; k < 3; k++) { B.lo[k] = std::min(B.lo[k], (float)v[k]); B.hi[k] = std::max(B.hi[k], (float)v[k]); scale = std::max(scale, std::abs(v[k])); }
        };
        auto reset = [](DBounds &B) { for (int k = 0; k < 3; k++) { B.lo[k] = std::numeric_limits<float>::infinity(); B.hi[k] = -B.lo[k]; } };
        for (int i = 0; i < h.nRects; i++) {
            DBounds &B = s->bounds[np++]; reset(B);
            for (int sx = -1; sx <= 1; sx += 2) for (int sy = -1; sy <= 1; sy += 2) grow(B, xfAffine(h.rects[i].toWorld, mk(sx, sy, 0)));
        }
        for (int i = 0; i < h.nSpheres; i++) {
            DBounds &B = s->bounds[np++]; reset(B);
            grow(B, h.spheres[i].center - splat(h.spheres[i].radius)); grow(B, h.spheres[i].center + splat(h.spheres[i].radius));
        }
        for (int i = 0; i < h.nTris; i++) {
            DBounds &B = s->bounds[np++]; reset(B);
            grow(B, h.tris[i].p0); grow(B, h.tris[i].p1); grow(B, h.tris[i].p2);
        }
        for (int i = 0; i < np; i++)        // pad by 1e-4 of the scene scale: >100x the fp32 rounding of the slab test (errors and
            for (int k = 0; k < 3; k++) {   // padding both scale with |1/d| per axis, so the margin holds for any ray direction)
                const float pad = (float)(1e-4 * (scale + (s->bounds[i].hi[k] - s->bounds[i].lo[k])) + 1e-6);
                s->bounds[i].lo[k] -= pad; s->bounds[i].hi[k] += pad;
            }
    }
    if (useBvh && !s->bvhTris.empty()) {
        // node padding: 2e-5 of the scene scale, >25x the worst-case rounding of the fp32 slab arithmetic (~4 * 2^-23 * scale),
        // small against the triangles of a >= 1e5-triangle scene so leaves stay tight
        buildBvh(s->bvhTris, s->bvh, 2e-5 * scale + 1e-7);
        h.nBvhNodes = (int)s->bvh.size(); h.nBvhTris = (int)s->bvhTris.size();
    }
    // emitters: DiscreteDistribution over samplingWeight (scene.cpp:357-380, pmf.h:100-114)
    h.nEmitters = d->n_emitters;
    h.emCdf[0] = 0.0;
    for (int i = 0; i < d->n_emitters; i++) h.emCdf[i + 1] = h.emCdf[i] + d->emitters[i].sampling_weight;
    const double sum = h.emCdf[d->n_emitters], norm = sum > 0 ? 1.0 / sum : 0.0;
    if (sum > 0) { for (int i = 1; i <= d->n_emitters; i++) h.emCdf[i] *= norm; h.emCdf[d->n_emitters] = 1.0; }
    for (int i = 0; i < d->n_emitters; i++) {
        const gdb200_emitter &e = d->emitters[i];
        DEmitter &o = h.emitters[i];
        o.radiance = mk(e.radiance[0], e.radiance[1], e.radiance[2]);
        o.pdfDiscrete = e.sampling_weight * norm;
        if (e.type == GDB200_EMITTER_ENVMAP) {
            if (h.env.present) return set_error(GDB200_ERR_ARGUMENT, "emitter %d: only one environment emitter is allowed", i);
            if (!d->envmap) return set_error(GDB200_ERR_ARGUMENT, "emitter %d: envmap emitter without gdb200_scene_desc.envmap", i);
            o.kind = EM_ENV; o.rect = -1;
            if (int rc = buildEnvTables(d->envmap, i, s)) return rc;
            continue;
        }
        if (e.type == GDB200_EMITTER_POINT) {
            o.kind = EM_POINT; o.rect = -1; o.position = mk(e.position[0], e.position[1], e.position[2]);
            continue;
        }
        if (e.type == GDB200_EMITTER_SPOT) {                                         // spot.cpp:70-94
            if (!(e.cutoff_angle >= e.beam_width)) return set_error(GDB200_ERR_ARGUMENT, "emitter %d: spot cutoffAngle must be >= beamWidth", i);
            o.kind = EM_SPOT; o.rect = -1; o.position = mk(e.position[0], e.position[1], e.position[2]);
            for (int k = 0; k < 9; k++) o.toLocal[k] = e.to_local[k];
            o.cutoffAngle = e.cutoff_angle; o.cosCutoff = std::cos(e.cutoff_angle); o.cosBeam = std::cos(e.beam_width);
            o.invTransition = 1.0f / (e.cutoff_angle - e.beam_width);
            continue;
        }
        if (e.type != GDB200_EMITTER_AREA) return set_error(GDB200_ERR_ARGUMENT, "emitter %d: unknown type %d", i, e.type);
        if (e.shape < 0 || e.shape >= d->n_shapes || d->shapes[e.shape].emitter != i)
            return set_error(GDB200_ERR_ARGUMENT, "emitter %d: area emitter and its shape must reference each other", i);
        if (rectOfShape[e.shape] >= 0) { o.kind = EM_RECT; o.rect = rectOfShape[e.shape]; }
        else if (sphereOfShape[e.shape] >= 0) {                                      // sphere.cpp:126-128: m_invSurfaceArea = 1 / (4 pi r^2)
            o.kind = EM_SPHERE; o.rect = sphereOfShape[e.shape];
            const double r = d->shapes[e.shape].radius;
            o.invArea = 1 / (4 * kPi * r * r);
        }
        else if (emTriFirstOfShape[e.shape] >= 0) {
            const gdb200_shape &sh = d->shapes[e.shape];
            if (sh.tri_count < 1) return set_error(GDB200_ERR_ARGUMENT, "emitter %d: Encountered an empty triangle mesh!", i);
            o.kind = EM_MESH; o.rect = -1; o.triFirst = emTriFirstOfShape[e.shape]; o.triCount = sh.tri_count;
            o.cdfFirst = (int)s->emTriCdf.size();
            s->emTriCdf.push_back(0.0);                                              // DiscreteDistribution::append / normalize, pmf.h:62-71,101-114
            for (int t = 0; t < sh.tri_count; t++) {
                const DEmTri &E = s->emTris[o.triFirst + t];
                s->emTriCdf.push_back(s->emTriCdf.back() + (Float)0.5f * len(cross(E.p1 - E.p0, E.p2 - E.p0)));
            }
            const Float areaSum = s->emTriCdf.back();
            if (areaSum > 0) {
                const Float normalization = (Float)1.0f / areaSum;
                for (int t = 1; t <= sh.tri_count; t++) s->emTriCdf[o.cdfFirst + t] *= normalization;
                s->emTriCdf[o.cdfFirst + sh.tri_count] = 1.0;
            }
            o.invArea = (Float)1.0f / areaSum;
        } else return set_error(GDB200_ERR_ARGUMENT, "emitter %d: area emitters are supported on rectangles, spheres and triangle meshes only", i);
    }
    return GDB200_OK;
}

// fresnelDiffuseReflectance(eta, fast = false) (util.cpp:814-862): adaptive Gauss-Lobatto quadrature
// (quad.cpp:287-403, maxEvals 1024, absError 0, relError 1e-5, no convergence estimate) of
// fresnelDielectricExt(sqrt(xi), eta) over xi in [0,1].  Host-side: two constants per plastic material.
struct DiffuseFresnelQuad {
    Float eta; size_t evals = 0; Float acc = 0;
    static Float fresnel(Float cosThetaI_, Float eta)                                // util.cpp:651-681
    {
        if (eta == 1) return 0.0;
        const Float scale = (cosThetaI_ > 0) ? 1 / eta : eta, cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
        if (cosThetaTSqr <= 0.0) return 1.0;
        const Float cosThetaI = std::abs(cosThetaI_), cosThetaT = std::sqrt(cosThetaTSqr);
        const Float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
        const Float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
        return 0.5 * (Rs * Rs + Rp * Rp);
    }
    Float f(Float xi) const { return fresnel(std::sqrt(xi), eta); }
    Float step(Float a, Float b, Float fa, Float fb)
    {
        const Float alpha = (Float)std::sqrt(2.0 / 3.0), beta = (Float)(1.0 / std::sqrt(5.0));
        const Float h = (b - a) / 2, m = (a + b) / 2;
        const Float mll = m - alpha * h, ml = m - beta * h, mr = m + beta * h, mrr = m + alpha * h;
        const Float fmll = f(mll), fml = f(ml), fm = f(m), fmr = f(mr), fmrr = f(mrr);
        const Float integral2 = (h / 6) * (fa + fb + 5 * (fml + fmr));
        const Float integral1 = (h / 1470) * (77 * (fa + fb) + 432 * (fmll + fmrr) + 625 * (fml + fmr) + 672 * fm);
        evals += 5;
        if (evals >= 1024) return integral1;
        const Float dist = acc + (integral1 - integral2);
        if (dist == acc || mll <= a || b <= mrr) return integral1;
        return step(a, mll, fa, fmll) + step(mll, ml, fmll, fml) + step(ml, m, fml, fm) + step(m, mr, fm, fmr) + step(mr, mrr, fmr, fmrr) + step(mrr, b, fmrr, fb);
    }
    Float integrate()
    {
        const Float a = 0, b = 1, alpha = (Float)std::sqrt(2.0 / 3.0), beta = (Float)(1.0 / std::sqrt(5.0));
        const Float x1 = 0.94288241569547971906, x2 = 0.64185334234578130578, x3 = 0.23638319966214988028;
        const Float m = (a + b) / 2, h = (b - a) / 2;
        const Float y1 = f(a), y3 = f(m - alpha * h), y5 = f(m - beta * h), y7 = f(m), y9 = f(m + beta * h), y11 = f(m + alpha * h), y13 = f(b);
        const Float q = h * ((Float)0.0158271919734801831 * (y1 + y13) + (Float)0.0942738402188500455 * (f(m - x1 * h) + f(m + x1 * h))
                             + (Float)0.1550719873365853963 * (y3 + y11) + (Float)0.1888215739601824544 * (f(m - x2 * h) + f(m + x2 * h))
                             + (Float)0.1997734052268585268 * (y5 + y9) + (Float)0.2249264653333395270 * (f(m - x3 * h) + f(m + x3 * h))
                             + (Float)0.2426110719014077338 * y7);
        evals = 13;
        const Float eps = std::numeric_limits<Float>::epsilon(), r = 1.0;
        acc = std::numeric_limits<Float>::infinity();
        if (q != 0) acc = q * std::max((Float)1e-5f, eps) / (r * eps);
        evals += 2;
        return step(a, b, f(a), f(b));
    }
};
inline Float fresnelDiffuseReflectance(Float eta) { DiffuseFresnelQuad q; q.eta = eta; return q.integrate(); }

// Per-material facts incl. the vertex classification of gpt.cpp:176-226 for this shiftThreshold.
inline void classifyMaterials(HostScene *s, double shiftThreshold)
{
    for (size_t i = 0; i < s->mats.size(); i++) {
        const gdb200_material &m = s->mats[i];
        DMaterial &o = s->host.materials[i];
        o.type = m.type; o.distribution = m.distribution;
        o.reflectance = mk(m.reflectance[0], m.reflectance[1], m.reflectance[2]);
        o.specR = mk(m.specular_reflectance[0], m.specular_reflectance[1], m.specular_reflectance[2]);
        o.specT = mk(m.specular_transmittance[0], m.specular_transmittance[1], m.specular_transmittance[2]);
        o.eta = mk(m.eta[0], m.eta[1], m.eta[2]); o.k = mk(m.k[0], m.k[1], m.k[2]);
        // The plugins store the clamped roughness (microfacet.h:67-71) in a ConstantFloatTexture and read it back through
        // Spectrum::average() (roughconductor.cpp:196,273; spectrum.h:481-486): result * (1.0f / N) with a SINGLE-precision
        // quotient, i.e. alpha * (1 + 3e-8) -- then the distribution clamps again.
        double alphaAvg = 0.0f;
        for (int c = 0; c < 3; c++) alphaAvg += std::max(m.alpha, (double)1e-4f);
        alphaAvg = alphaAvg * (1.0f / 3);
        o.alpha = std::max(alphaAvg, (double)1e-4f);
        o.iorRatio = m.ior_ratio;
        o.bsdfEta = (m.type == GDB200_BSDF_DIELECTRIC || m.type == GDB200_BSDF_ROUGHDIELECTRIC) ? m.ior_ratio : 1.0;   // bsdf.cpp:62-64, dielectric.cpp:389, roughdielectric.cpp:631; plastic and twosided inherit 1
        o.twosided = m.twosided != 0; o.nonlinear = m.nonlinear != 0;
        int nComp = 1; double rough[2] = {0, 0};
        const double inf = std::numeric_limits<double>::infinity();
        switch (m.type) {
            case GDB200_BSDF_DIFFUSE:                                                // diffuse.cpp:97-101,167-169
                o.flags = std::max(m.reflectance[0], std::max(m.reflectance[1], m.reflectance[2])) > 0 ? (EDiffuseReflection | EFrontSide) : 0;
                nComp = o.flags ? 1 : 0; rough[0] = inf; break;
            case GDB200_BSDF_ROUGHCONDUCTOR: o.flags = EGlossyReflection | EFrontSide; rough[0] = 0.5 * (alphaAvg + alphaAvg); break;   // roughconductor.cpp:437-440
            case GDB200_BSDF_CONDUCTOR: o.flags = EDeltaReflection | EFrontSide; rough[0] = 0; break;
            case GDB200_BSDF_ROUGHDIELECTRIC:                                        // roughdielectric.cpp:246-252,642-645
                o.flags = EGlossyReflection | EGlossyTransmission | EFrontSide | EBackSide; nComp = 2; rough[0] = rough[1] = 0.5 * (alphaAvg + alphaAvg); break;
            case GDB200_BSDF_PLASTIC: {                                              // plastic.cpp:186-217,442-449
                o.flags = EDeltaReflection | EDiffuseReflection | EFrontSide; nComp = 2; rough[0] = 0; rough[1] = inf;
                o.fdrInt = fresnelDiffuseReflectance(1 / m.ior_ratio); o.fdrExt = fresnelDiffuseReflectance(m.ior_ratio);
                const Float dAvg = luminance(o.reflectance), sAvg = luminance(o.specR);
                o.specSamplingWeight = sAvg / (dAvg + sAvg);
                o.invEta2 = 1 / (m.ior_ratio * m.ior_ratio);
                break;
            }
            default: o.flags = EDeltaReflection | EDeltaTransmission | EFrontSide | EBackSide; nComp = 2; break;
        }
        if (o.twosided && o.flags) o.flags |= EBackSide;                             // twosided.cpp:95-101: front components + the same set as back components
        o.refNFromShading = (o.flags & (ETransmissionBits | EBackSide)) == 0;        // records.inl:160-165
        for (int deltaQuery = 0; deltaQuery < 2; deltaQuery++) {                     // gpt.cpp:194-226
            double lowest = inf; bool found_smooth = false, found_dirac = false;
            for (int c = 0; c < nComp; c++) {
                const double r = rough[c];
                if (r == 0) { found_dirac = true; if (!deltaQuery) continue; } else found_smooth = true;
                if (r < lowest) lowest = r;
            }
            if (!found_smooth && found_dirac && !deltaQuery) lowest = 0;
            (deltaQuery ? o.vtDelta : o.vtSmooth) = lowest <= shiftThreshold ? VERTEX_TYPE_GLOSSY : VERTEX_TYPE_DIFFUSE;
        }
    }
}

// Parameter validation of gpt.cpp:1194-1210 / integrator.cpp:190-225 and the launch geometry of one render:
// which pixels this call owns (row strip or interleaved bands), how many sample streams they form and how
// many path slots are resident.  Fills every non-pointer field of GptArgs.
inline int setupArgs(const HostScene &s, const gdb200_gpt_params *p, GptArgs &a, int maxSlots = 1 << 23)
{
    if (p->max_depth <= 0 && p->max_depth != -1) return set_error(GDB200_ERR_ARGUMENT, "'maxDepth' must be set to -1 (infinite) or a value greater than zero!");
    if (p->rr_depth <= 0) return set_error(GDB200_ERR_ARGUMENT, "'rrDepth' must be set to a value greater than zero!");
    if (p->spp <= 0) return set_error(GDB200_ERR_ARGUMENT, "sampleCount must be positive");
    if (p->streams_per_pixel < 0 || p->streams_per_pixel > 4096) return set_error(GDB200_ERR_ARGUMENT, "streams_per_pixel must be in [0, 4096]");
    const bool banded = p->band_count > 1;
    const bool all = banded || (p->y_begin == 0 && p->y_end == 0);
    const int y0 = all ? 0 : p->y_begin, y1 = all ? s.height : p->y_end;
    if (y0 < 0 || y1 > s.height || y0 >= y1) return set_error(GDB200_ERR_ARGUMENT, "invalid row range [%d,%d)", y0, y1);
    int ownedRows = y1 - y0;
    if (banded) {
        if (p->band_rows <= 0 || p->band_index < 0 || p->band_index >= p->band_count)
            return set_error(GDB200_ERR_ARGUMENT, "invalid band sharding (%d rows, index %d of %d)", p->band_rows, p->band_index, p->band_count);
        ownedRows = 0;
        for (int y = 0; y < s.height; y++) ownedRows += ((y / p->band_rows) % p->band_count) == p->band_index;
        if (ownedRows == 0) return set_error(GDB200_ERR_ARGUMENT, "band sharding leaves rank %d without rows", p->band_index);
    }
    memset(&a, 0, sizeof(a));
    a.width = s.width; a.height = s.height; a.yBegin = y0;
    a.nPixels = s.width * ownedRows;
    a.streamsPerPixel = std::max(1, p->streams_per_pixel);
    const long long nStreams = (long long)a.nPixels * a.streamsPerPixel;
    if (nStreams > 0x7fffffffLL) return set_error(GDB200_ERR_ARGUMENT, "too many sample streams (%lld)", nStreams);
    a.nStreams = (int)nStreams;
    if (p->max_slots < 0) return set_error(GDB200_ERR_ARGUMENT, "max_slots must not be negative");
    if (p->max_slots > 0) maxSlots = p->max_slots;
    a.nSlots = (int)std::min<long long>(nStreams, std::max(1, maxSlots));
    a.spp = p->spp; a.seed = p->seed; a.skipPreview = p->skip_preview != 0;
    a.bandRows = banded ? p->band_rows : 0; a.bandCount = banded ? p->band_count : 0; a.bandIndex = banded ? p->band_index : 0;
    a.cfg.maxDepth = p->max_depth; a.cfg.minDepth = 1; a.cfg.rrDepth = p->rr_depth;         // gpt.cpp:1368-1371
    a.cfg.strictNormals = p->strict_normals; a.cfg.shiftThreshold = p->shift_threshold;
    // gpt.cpp:957 reads an uninitialised DirectSamplingRecord::measure; GDB200_GPT_REF_UNINIT_MEASURE (include/gdb200.h)
    // selects what a g++ -O2 build of the reference does there instead of the intended ESolidAngle.
    a.cfg.refUninitMeasure = (p->flags & GDB200_GPT_REF_UNINIT_MEASURE) ? 1 : 0; a.cfg.pad = 0;
    return GDB200_OK;
}

}  // namespace gdb200
