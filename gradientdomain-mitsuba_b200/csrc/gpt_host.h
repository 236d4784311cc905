// gdb200 G-PT tracer — host-side flattening of a gdb200_scene_desc into the device tables of
// gpt_device.cuh (pure C++: no CUDA runtime calls; gpt.cu uploads the result).  Requires
// gdb200::set_error to be declared by the including translation unit (common.h in the library).
#pragma once
#include "gpt_kernels.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace gdb200 {

struct HostScene {
    DScene host;                       // flattened tables (vertex classification filled per render)
    DBounds bounds[kMaxPrims];         // padded per-primitive bounds (candidate selection)
    std::vector<gdb200_material> mats;
    int width = 0, height = 0;
};

inline int flattenScene(const gdb200_scene_desc *d, HostScene *s)
{
    DScene &h = s->host;
    memset(&h, 0, sizeof(h));
    const gdb200_camera &c = d->camera;
    if (c.width <= 0 || c.height <= 0) return set_error(GDB200_ERR_ARGUMENT, "invalid film size %dx%d", c.width, c.height);
    if (d->n_emitters < 1) return set_error(GDB200_ERR_ARGUMENT, "scene has no emitter");
    if (d->n_materials > kMaxMaterials || d->n_emitters > kMaxEmitters)
        return set_error(GDB200_ERR_ARGUMENT, "too many materials/emitters (%d/%d, limits %d/%d)", d->n_materials, d->n_emitters, kMaxMaterials, kMaxEmitters);
    memcpy(h.sampleToCamera, c.sample_to_camera, sizeof(h.sampleToCamera));
    memcpy(h.cameraToWorld, c.camera_to_world, sizeof(h.cameraToWorld));
    h.nearClip = c.near_clip; h.farClip = c.far_clip; h.width = c.width; h.height = c.height;
    h.invResX = 1.0 / c.width; h.invResY = 1.0 / c.height;
    h.filterRadius = d->rfilter_radius; h.filterTap = 1.0 / (2 * d->rfilter_radius); h.filterScale = 31 / d->rfilter_radius;
    s->width = c.width; s->height = c.height;
    s->mats.assign(d->materials, d->materials + d->n_materials);
    h.nMaterials = d->n_materials;
    std::vector<int> rectOfShape(d->n_shapes, -1);
    for (int i = 0; i < d->n_shapes; i++) {
        const gdb200_shape &sh = d->shapes[i];
        if (sh.material < 0 || sh.material >= d->n_materials) return set_error(GDB200_ERR_ARGUMENT, "shape %d: bad material index", i);
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
            r.material = sh.material; r.emitter = sh.emitter;
            rectOfShape[i] = h.nRects++;
        } else if (sh.type == GDB200_SHAPE_SPHERE) {
            if (h.nSpheres >= kMaxSpheres) return set_error(GDB200_ERR_ARGUMENT, "too many spheres (limit %d)", kMaxSpheres);
            if (sh.emitter >= 0) return set_error(GDB200_ERR_ARGUMENT, "shape %d: sphere emitters are not supported yet", i);
            DSphere &sp = h.spheres[h.nSpheres++];
            sp.center = mk(sh.center[0], sh.center[1], sh.center[2]); sp.radius = sh.radius; sp.flip = sh.flip_normals;
            sp.material = sh.material; sp.emitter = -1;
        } else if (sh.type == GDB200_SHAPE_MESH) {
            if (sh.emitter >= 0) return set_error(GDB200_ERR_ARGUMENT, "shape %d: mesh emitters are not supported yet", i);
            if (h.nMeshes >= kMaxMeshes) return set_error(GDB200_ERR_ARGUMENT, "too many meshes (limit %d)", kMaxMeshes);
            DMesh &M = h.meshes[h.nMeshes++];
            M.first = h.nTris; M.count = 0;
            const double big = std::numeric_limits<double>::infinity();
            M.lo = mk(big, big, big); M.hi = mk(-big, -big, -big);
            std::vector<DTri> meshTris;
            for (int t = sh.first_tri; t < sh.first_tri + sh.tri_count; t++) {
                if (h.nTris + (int)meshTris.size() >= kMaxTris) return set_error(GDB200_ERR_ARGUMENT, "too many triangles for the constant-memory scene table (limit %d); the BVH path is not built yet", kMaxTris);
                if (t < 0 || t >= d->n_triangles) return set_error(GDB200_ERR_ARGUMENT, "shape %d: triangle range out of bounds", i);
                const int *ix = d->triangles + 3 * t;
                const double *va = d->vertices + 3 * ix[0], *vb = d->vertices + 3 * ix[1], *vc = d->vertices + 3 * ix[2];
                const V3 A = mk(va[0], va[1], va[2]), B = mk(vb[0], vb[1], vb[2]), C = mk(vc[0], vc[1], vc[2]);
                for (const V3 &P : {A, B, C}) {
                    M.lo = mk(std::min(M.lo.x, P.x), std::min(M.lo.y, P.y), std::min(M.lo.z, P.z));
                    M.hi = mk(std::max(M.hi.x, P.x), std::max(M.hi.y, P.y), std::max(M.hi.z, P.z));
                }
                DTri T;                                                              // TriAccel::load, triaccel.h:61-95
                memset(&T, 0, sizeof(T));
                static const int waldModulo[4] = {1, 2, 0, 1};
                const V3 b = C - A, cc = B - A, N = cross(cc, b);
                const double Nv[3] = {N.x, N.y, N.z}, bv[3] = {b.x, b.y, b.z}, cv[3] = {cc.x, cc.y, cc.z}, Av[3] = {A.x, A.y, A.z};
                int k = 0;
                for (int j = 0; j < 3; j++) if (std::abs(Nv[j]) > std::abs(Nv[k])) k = j;
                const int u = waldModulo[k], v = waldModulo[k + 1];
                const double n_k = Nv[k], denom = bv[u] * cv[v] - bv[v] * cv[u];
                T.p0 = A; T.p1 = B; T.p2 = C; T.material = sh.material; T.emitter = -1;
                if (denom == 0) continue;                                            // degenerate: k = 3, never hit (triaccel.h:75-78)
                T.k = k;
                T.n_u = Nv[u] / n_k; T.n_v = Nv[v] / n_k; T.n_d = dot(A, N) / n_k;
                T.b_nu = bv[u] / denom; T.b_nv = -bv[v] / denom; T.a_u = Av[u]; T.a_v = Av[v];
                T.c_nu = cv[v] / denom; T.c_nv = -cv[u] / denom;
                V3 faceNormal = cross(B - A, C - A);                                 // skdtree.h:367-371
                const double l = len(faceNormal);
                if (!isZero(faceNormal)) faceNormal = faceNormal / l;
                T.faceNormal = faceNormal;
                meshTris.push_back(T);
            }
            for (int k = 0; k < 3; k++) {          // store grouped by projection axis (order inside a group is kept)
                for (const DTri &T : meshTris) if (T.k == k) h.tris[h.nTris++] = T;
                M.kEnd[k] = h.nTris;
            }
            M.count = h.nTris - M.first;
        } else return set_error(GDB200_ERR_ARGUMENT, "shape %d: unknown type %d", i, sh.type);
    }
    for (int mi = 0; mi < h.nMeshes; mi++) {   // enlarge the skip-bounds far beyond any rounding of the slab test
        DMesh &M = h.meshes[mi];
        const V3 ext = M.hi - M.lo;
        const double pad = 1e-6 * std::max(1.0, std::max(ext.x, std::max(ext.y, ext.z))) + 1e-9 * std::max(maxComp(M.hi), -std::min(M.lo.x, std::min(M.lo.y, M.lo.z)));
        M.lo = M.lo - splat(pad); M.hi = M.hi + splat(pad);
    }
    // padded bounds of every primitive for the candidate pass of closestPrimitive
    {
        double scale = 0;
        for (int k = 0; k < 3; k++) scale = std::max(scale, std::abs(c.camera_to_world[4 * k + 3]));
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
    // emitters: DiscreteDistribution over samplingWeight (scene.cpp:357-380, pmf.h:100-114)
    h.nEmitters = d->n_emitters;
    h.emCdf[0] = 0.0;
    for (int i = 0; i < d->n_emitters; i++) h.emCdf[i + 1] = h.emCdf[i] + d->emitters[i].sampling_weight;
    const double sum = h.emCdf[d->n_emitters], norm = sum > 0 ? 1.0 / sum : 0.0;
    if (sum > 0) { for (int i = 1; i <= d->n_emitters; i++) h.emCdf[i] *= norm; h.emCdf[d->n_emitters] = 1.0; }
    for (int i = 0; i < d->n_emitters; i++) {
        const gdb200_emitter &e = d->emitters[i];
        if (e.shape < 0 || e.shape >= d->n_shapes || rectOfShape[e.shape] < 0)
            return set_error(GDB200_ERR_ARGUMENT, "emitter %d: only rectangle area emitters are supported", i);
        h.emitters[i].rect = rectOfShape[e.shape];
        h.emitters[i].radiance = mk(e.radiance[0], e.radiance[1], e.radiance[2]);
        h.emitters[i].pdfDiscrete = e.sampling_weight * norm;
    }
    return GDB200_OK;
}

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
        o.alpha = std::max(m.alpha, (double)1e-4f);                                  // microfacet.h:67-71
        o.iorRatio = m.ior_ratio;
        o.bsdfEta = m.type == GDB200_BSDF_DIELECTRIC ? m.ior_ratio : 1.0;            // bsdf.cpp:62-64, dielectric.cpp:389
        int nComp = 1; double rough[2] = {0, 0};
        const double inf = std::numeric_limits<double>::infinity();
        switch (m.type) {
            case GDB200_BSDF_DIFFUSE:                                                // diffuse.cpp:97-101,167-169
                o.flags = std::max(m.reflectance[0], std::max(m.reflectance[1], m.reflectance[2])) > 0 ? (EDiffuseReflection | EFrontSide) : 0;
                nComp = o.flags ? 1 : 0; rough[0] = inf; break;
            case GDB200_BSDF_ROUGHCONDUCTOR: o.flags = EGlossyReflection | EFrontSide; rough[0] = 0.5 * (m.alpha + m.alpha); break;   // roughconductor.cpp:437-440
            case GDB200_BSDF_CONDUCTOR: o.flags = EDeltaReflection | EFrontSide; rough[0] = 0; break;
            default: o.flags = EDeltaReflection | EDeltaTransmission | EFrontSide | EBackSide; nComp = 2; break;
        }
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
inline int setupArgs(const HostScene &s, const gdb200_gpt_params *p, GptArgs &a, int maxSlots = 1 << 20)
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
    a.nSlots = (int)std::min<long long>(nStreams, std::max(1, maxSlots));
    a.spp = p->spp; a.seed = p->seed; a.skipPreview = p->skip_preview != 0;
    a.bandRows = banded ? p->band_rows : 0; a.bandCount = banded ? p->band_count : 0; a.bandIndex = banded ? p->band_index : 0;
    a.cfg.maxDepth = p->max_depth; a.cfg.minDepth = 1; a.cfg.rrDepth = p->rr_depth;         // gpt.cpp:1368-1371
    a.cfg.strictNormals = p->strict_normals; a.cfg.shiftThreshold = p->shift_threshold;
    return GDB200_OK;
}

}  // namespace gdb200
