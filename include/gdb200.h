/* gdb200 — C ABI of the B200-native G-PT + screened-Poisson hot path.
 *
 * This is the drop-in boundary: plain C, caller-owned buffers,
 * status codes, no exceptions, no torch/C++ types.  A Mitsuba `gpt` integrator
 * plugin shim calls these instead of
 *   - GradientPathIntegrator::render's block scheduler + renderBlock
 *     (reference src/integrators/gpt/gpt.cpp:1358-1413, 1220-1355), and
 *   - poisson::Solver::{importImagesMTS,setupBackend,solveIndirect,
 *     exportImagesMTS} (reference src/integrators/poisson_solver/Solver.hpp:113-117,
 *     call site gpt.cpp:1445-1462).
 *
 * Image buffers: row-major, top-left origin, interleaved RGB (= Vec3f /
 * Spectrum AoS, Solver.cpp:224-227).  Every function returns 0 on success;
 * on failure the message is available from gdb200_last_error() (thread-local).
 * All functions fail (never fall back to a CPU path) when no CUDA device or
 * kernel image is available.
 */
#ifndef GDB200_H
#define GDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDB200_OK            0
#define GDB200_ERR_ARGUMENT  1
#define GDB200_ERR_CUDA      2
#define GDB200_ERR_NO_DEVICE 3
#define GDB200_ERR_CANCELLED 4

/* ------------------------------------------------------------------ misc */

int         gdb200_version(void);                 /* 100*major + minor */
const char *gdb200_last_error(void);              /* thread-local, never NULL */
int         gdb200_device_count(int *out_count);
int         gdb200_set_device(int device);
/* Pinned host allocation helpers so callers can stage film buffers for full
 * PCIe/NVLink-C2C copy bandwidth. */
int         gdb200_host_alloc(void **out_ptr, size_t bytes);
int         gdb200_host_free(void *ptr);
/* Plain device buffers for callers without their own CUDA runtime (the Mitsuba shim). */
int         gdb200_device_alloc(void **out_ptr, size_t bytes);
int         gdb200_device_free(void *ptr);
int         gdb200_device_download(void *host_dst, const void *device_src, size_t bytes);

typedef struct gdb200_stats {
    double device_ms;       /* CUDA-event time of the device work of this call   */
    double h2d_ms, d2h_ms;  /* copies performed by host-pointer entry points     */
    int    launches;        /* kernels of this library launched by this call     */
    int    irls_iters;      /* IRLS iterations executed (solver)                 */
    int    cg_iters;        /* total CG iterations executed (solver)             */
    int    reserved0;
    double samples;         /* evaluatePoint() samples traced (tracer)           */
    double rays;            /* rays cast (tracer)                                */
    double path_vertices;   /* sum of base-path depths (tracer; "Average path length", gpt.cpp:1178-1179) */
    double state_bytes;     /* algorithmic wavefront state bytes of the bounce kernel, SURVEY.md §8d record sizes (tracer) */
    double bounce_ms;       /* summed CUDA-event time of the gpt_bounce_kernel launches   */
    double generate_ms;     /* ... of the gpt_generate_kernel launches                    */
    double compact_ms;      /* ... of the gpt_compact_kernel launches                     */
    double path_bounces;    /* base-path bounce iterations executed (tracer)              */
    int    bounce_launches; /* wavefront steps (staged wavefront: gpt_stage_kernel<shade> launches) */
    int    reserved1;
    /* staged wavefront (csrc/gpt_stages.cuh): time per kernel family, summed over the render from the GPU's nanosecond timer
     * (the kernel that opens a phase stamps %globaltimer; the fused A/B path uses CUDA events); bounce_ms = shade stage */
    double cast_ms;         /* gpt_cast_kernel (nearest-hit + any-hit queues)             */
    double prepare_ms;      /* gpt_stage_kernel<prepare>                                  */
    double resolve_ms;      /* gpt_stage_kernel<resolve>                                  */
    double primary_ms;      /* gpt_stage_kernel<primary>                                  */
    double rays_cast;       /* rays answered by gpt_cast_kernel                           */
} gdb200_stats;

/* ------------------------------------------------ screened Poisson solver */

/* Mirrors poisson::Solver::Params' solver configuration (Solver.hpp:82-90). */
typedef struct gdb200_poisson_config {
    int   irlsIterMax;
    float irlsRegInit;
    float irlsRegIter;
    int   cgIterMax;
    int   cgIterCheck;
    float cgTolerance;
} gdb200_poisson_config;

/* Solver::Params::setConfigPreset (Solver.cpp:90-164): "L1D","L1Q","L1L","L2D","L2Q". */
int gdb200_poisson_preset(const char *preset, gdb200_poisson_config *out_cfg);

typedef struct gdb200_poisson_plan gdb200_poisson_plan;

/* Allocates the device workspace for a w x h solve on the current device
 * (replaces Solver::setupBackend's allocVector calls, Solver.cpp:296-313). */
int  gdb200_poisson_plan_create(int w, int h, gdb200_poisson_plan **out_plan);
void gdb200_poisson_plan_destroy(gdb200_poisson_plan *plan);

/* Sharded solve (SURVEY.md §8e "row-shard >= 4K"; no reference counterpart -- the reference's solver is single-device): n_ranks
 * GPUs, one process (or host thread) each, solve ONE image together.  Rank r owns the row band [y0, y1); bands cover the image
 * in rank order.  Inside the one persistent kernel each GPU pushes the first / last row of the CG vectors it updates into its
 * neighbours' halo rows with stores over NVLink peer memory, and the two reductions per CG iteration travel as 32-byte
 * messages into every GPU's mailbox -- no host round trip, no separate collective.  Wiring: every rank creates its shard,
 * exports a handle (CUDA IPC handles of its plane array and mailbox), the host layer all-gathers the n_ranks handles
 * (torch.distributed in gdb200.poisson.ShardedPoissonSolver) and every rank connects.  Then all ranks call
 * gdb200_poisson_solve_device on their shard with the WHOLE input images resident on their own GPU; each writes its band's rows
 * of out_final.  The sums are added in rank order, so every rank takes the same branches; results differ from the one-GPU
 * solve by reduction order only.  A rank whose peers do not show up gives up after 8 s with GDB200_ERR_CUDA
 * (the shards have lost step then: destroy and re-create them). */
typedef struct gdb200_shard_handle {
    unsigned char planes[64], mail[64];            /* cudaIpcMemHandle_t */
    unsigned long long planes_ptr, mail_ptr;       /* the same allocations as plain pointers, for peers in the same process */
    long long pid;
    int device, rank, y0, y1, w, h;
} gdb200_shard_handle;
int  gdb200_poisson_shard_create(int w, int h, int y0, int y1, int rank, int n_ranks, gdb200_poisson_plan **out_plan);
int  gdb200_poisson_shard_export(gdb200_poisson_plan *plan, void *out_handle /* gdb200_shard_handle */);
int  gdb200_poisson_shard_connect(gdb200_poisson_plan *plan, const void *handles /* n gdb200_shard_handle, rank order */, int n);

/* Tuning knob without a reference counterpart: the solver kernel has variants that differ in where the CG vectors live, not in
 * arithmetic -- 0: everything streams through L2; 1: x and Ap of every CTA's tiles stay in shared memory (images up to 2 tiles
 * per CTA, ~1.2 Mpixel on a B200); 2: x stays in shared memory (up to 4 tiles per CTA, ~2.4 Mpixel); 3: as 0 with the search
 * direction exchanged through a shared tile.  plan_create picks the fastest that fits (1, else 2, else 0).  All variants
 * return the same bits (tests/test_poisson_gpu.py); set_variant fails with GDB200_ERR_ARGUMENT if the image does not fit. */
int  gdb200_poisson_plan_set_variant(gdb200_poisson_plan *plan, int variant);
int  gdb200_poisson_plan_variant(const gdb200_poisson_plan *plan);

/* Device-pointer entry: inputs/outputs are resident in HBM (w*h*3 floats each).
 * d_throughput may be NULL (alpha := 0, x0 := 0; Solver.cpp:319,334-337) and
 * d_direct may be NULL (final := x; Solver.cpp:561-562).  `stream` is a
 * cudaStream_t (NULL = default stream).  Asynchronous unless stats != NULL. */
int gdb200_poisson_solve_device(gdb200_poisson_plan *plan,
                                const float *d_dx, const float *d_dy,
                                const float *d_throughput, const float *d_direct,
                                float alpha, const gdb200_poisson_config *cfg,
                                float *d_out_final, void *stream, gdb200_stats *stats);

/* Solver::evaluateMetricsMTS (Solver.cpp:511-541; Solver.hpp:116) on the result of the last solve of `plan`: e = b - P*x with
 * b = [alpha*throughput; dx; dy]; d_err (device, w*h*3 floats) receives the primal block of e, *out_errL1 / *out_errL2 the mean
 * of |e_i| and of |e_i|^2 over the 3*w*h RGB elements of e.  Synchronous. */
int gdb200_poisson_metrics_device(gdb200_poisson_plan *plan, float *d_err, float *out_errL1, float *out_errL2, void *stream);
/* The same after this thread's last gdb200_poisson_solve (host buffers): err = w*h*3 floats. */
int gdb200_poisson_metrics(float *err, float *out_errL1, float *out_errL2);

/* Host-pointer entry = importImagesMTS + setupBackend + solveIndirect +
 * exportImagesMTS (gpt.cpp:1456-1462) in one call.  Copies in, solves, copies out. */
int gdb200_poisson_solve(const float *dx, const float *dy, const float *throughput,
                         const float *direct, int w, int h, float alpha,
                         const char *preset, float *out_final, gdb200_stats *stats);


/* ------------------------------------------------------ flattened scene */

/* The plugin shim flattens Mitsuba's Scene into these plain structs once per
 * render.  All reals are fp64 (the reference requires a
 * DOUBLE_PRECISION build, README.txt:115-118).  Matrices are row-major 4x4. */

enum { GDB200_SHAPE_RECTANGLE = 0,   /* src/shapes/rectangle.cpp: unit square [-1,1]^2 in z=0 under to_world */
       GDB200_SHAPE_SPHERE    = 1,   /* src/shapes/sphere.cpp: centre + radius (no rotation)               */
       GDB200_SHAPE_MESH      = 2 }; /* TriMesh: triangles [first_tri, first_tri+tri_count), flat or with vertex normals */

enum { GDB200_BSDF_DIFFUSE        = 0,   /* src/bsdfs/diffuse.cpp        */
       GDB200_BSDF_ROUGHCONDUCTOR = 1,   /* src/bsdfs/roughconductor.cpp (sampleVisible = true) */
       GDB200_BSDF_CONDUCTOR      = 2,   /* src/bsdfs/conductor.cpp      */
       GDB200_BSDF_DIELECTRIC     = 3,   /* src/bsdfs/dielectric.cpp     */
       GDB200_BSDF_PLASTIC        = 4,   /* src/bsdfs/plastic.cpp: delta reflection + diffuse lobe (the two-component case of gpt.cpp:194-226) */
       GDB200_BSDF_ROUGHDIELECTRIC = 5 }; /* src/bsdfs/roughdielectric.cpp (sampleVisible = true, isotropic alpha): glossy reflection +
                                           * transmission, the non-delta refraction half-vector shift (gpt.cpp:245-290); draws one extra
                                           * sampler value inside BSDF::sample (EUsesSampler) */

enum { GDB200_MICROFACET_BECKMANN = 0, GDB200_MICROFACET_GGX = 1 };   /* src/bsdfs/microfacet.h */

typedef struct gdb200_camera {          /* src/sensors/perspective.cpp:126-180,271-298 */
    double sample_to_camera[16];        /* m_sampleToCamera (projective)               */
    double camera_to_world[16];         /* m_worldTransform at shutter open            */
    double near_clip, far_clip;
    int    width, height;               /* film / crop size                            */
    double aperture_radius;             /* 0 = pinhole `perspective`; > 0 = `thinlens` (src/sensors/thinlens.cpp:289-318): one */
    double focus_distance;              /* aperture sample per camera sample, shared by the base and the four offset rays (gpt.cpp:1263-1265) */
} gdb200_camera;

typedef struct gdb200_shape {
    int    type;                        /* GDB200_SHAPE_*                              */
    int    material;                    /* index into materials (every shape has one; shape.cpp:48-72) */
    int    emitter;                     /* index into emitters or -1                   */
    int    flip_normals;                /* sphere only                                 */
    double to_world[16], to_object[16]; /* rectangle                                   */
    double center[3], radius;           /* sphere                                      */
    int    first_tri, tri_count;        /* mesh                                        */
    int    has_vertex_normals;          /* mesh: shading normals interpolated from gdb200_scene_desc.normals (skdtree.h:383-394) */
    int    reserved;
} gdb200_shape;

typedef struct gdb200_material {
    int    type;                        /* GDB200_BSDF_*                               */
    int    distribution;                /* GDB200_MICROFACET_* (roughconductor)        */
    double reflectance[3];              /* diffuse; plastic: diffuseReflectance        */
    double specular_reflectance[3];     /* conductors, dielectric, plastic             */
    double specular_transmittance[3];   /* dielectric                                  */
    double eta[3], k[3];                /* conductors: complex IOR per channel         */
    double alpha;                       /* roughconductor, roughdielectric             */
    double ior_ratio;                   /* dielectric, plastic, roughdielectric: intIOR / extIOR */
    int    twosided;                    /* wrapped in <bsdf type="twosided"> (same BRDF on both sides, src/bsdfs/twosided.cpp); reflection-only BSDFs */
    int    nonlinear;                   /* plastic: nonlinear colour shifts (plastic.cpp:176)                       */
} gdb200_material;

enum { GDB200_EMITTER_AREA   = 0,     /* src/emitters/area.cpp on a rectangle, a sphere or a triangle mesh      */
       GDB200_EMITTER_ENVMAP = 1,     /* src/emitters/envmap.cpp: the scene's environment emitter (at most one) */
       GDB200_EMITTER_POINT  = 2,     /* src/emitters/point.cpp: isotropic point light (the EDiscrete branch of gpt.cpp:668-672) */
       GDB200_EMITTER_SPOT   = 3 };   /* src/emitters/spot.cpp: point light with a linear cone falloff (no projection texture)  */

typedef struct gdb200_emitter {
    int    shape;                       /* area: the rectangle / sphere / mesh shape that emits; envmap, point: -1 */
    int    type;                        /* GDB200_EMITTER_*                            */
    double radiance[3];                 /* area: radiance; point, spot: intensity      */
    double sampling_weight;             /* emitter.cpp:103, default 1                  */
    double position[3];                 /* point, spot: world position                 */
    double to_local[9];                 /* spot: rows of the inverse world transform's 3x3 block (trafo.inverse() applied to a vector, spot.cpp:199); the cone axis is local +z */
    double cutoff_angle, beam_width;    /* spot: radians (spot.cpp:71-74, defaults 20 deg and 3/4 of it); cutoff_angle >= beam_width */
} gdb200_emitter;

/* Latitude-longitude environment map (envmap.cpp).  Texels are the top MIP level as the reference holds it
 * (Emitter::getBitmap: the pyramid stores HALF-precision texels, envmap.cpp:102-103, so these are float16-representable values); lookups are bilinear on that level, u repeats, v clamps (envmap.cpp:392-396,
 * mipmap.h:503-596).  The bounding sphere is what EnvironmentMap::createShape derives from the scene:
 * scene->getAABB().getBSphere() with its radius * 1.5 (envmap.cpp:325-329). */
typedef struct gdb200_envmap {
    int          width, height;
    const float *rgb;                   /* height * width * 3, row 0 = +y pole         */
    double       scale;                 /* 'scale' parameter                           */
    double       to_world[16], to_object[16];
    double       bsphere_center[3], bsphere_radius;
} gdb200_envmap;

typedef struct gdb200_scene_desc {
    gdb200_camera          camera;
    double                 rfilter_radius;   /* box filter radius incl. its +1e-5 (box.cpp:38) */
    int                    n_shapes, n_materials, n_emitters, n_vertices, n_triangles;
    const gdb200_shape    *shapes;
    const gdb200_material *materials;
    const gdb200_emitter  *emitters;
    const double          *vertices;         /* n_vertices * 3                          */
    const int             *triangles;        /* n_triangles * 3 vertex indices          */
    const gdb200_envmap   *envmap;           /* NULL, or the map of the emitter whose type is GDB200_EMITTER_ENVMAP */
    const double          *normals;          /* NULL, or n_vertices * 3 vertex normals (read for shapes with has_vertex_normals) */
    /* The film's reconstruction filter as ImageBlock::put uses it: ReconstructionFilter::m_values (rfilter.cpp:37-55,
     * MTS_FILTER_RESOLUTION = 31 taps over [0, radius] + a trailing 0), looked up with evalDiscretized (rfilter.h:76-77).
     * All zeros = the box filter: taps 1/(2*rfilter_radius). */
    double                 rfilter_table[32];
} gdb200_scene_desc;

/* ------------------------------------------------------- G-PT integrator */

/* Same names, defaults and validation as the reference's XML parameters
 * (gpt.cpp:1194-1210, integrator.cpp:190-225, docs.xml:537-559). */
typedef struct gdb200_gpt_params {
    int      max_depth;          /* maxDepth, -1 = infinite                         */
    int      rr_depth;           /* rrDepth (5)                                     */
    int      strict_normals;     /* strictNormals (false)                           */
    double   shift_threshold;    /* shiftThreshold (0.001)                          */
    int      spp;                /* sampler sampleCount                             */
    int      skip_preview;       /* 1: do not accumulate the "-final" preview (gpt.cpp:1319-1324); use when a reconstruction will overwrite it */
    uint64_t seed;               /* gdb200_counter sampler seed                     */
    int      y_begin, y_end;     /* rows of base pixels this call renders (tile sharding); 0,0 = all */
    /* Interleaved row bands (load-balanced tile sharding): when band_count > 1 this call renders the
     * rows y with (y / band_rows) % band_count == band_index, and y_begin/y_end are ignored. */
    int      band_rows, band_count, band_index;
    /* Sample streams per pixel (0 or 1 = one).  The gdb200_counter sampler gives every pixel ONE stream that
     * its spp samples consume sequentially (Sampler::generate(pixel), gpt.cpp:1250-1251), which caps the
     * parallelism at one path per pixel.  With C > 1 the spp samples of a pixel are split into C chunks
     * (chunk c gets spp/C samples, +1 for c < spp%C), chunk 0 on the pixel's stream and chunk c > 0 on an
     * independently re-keyed one: the same film as C reference passes with sampleCount spp/C summed. */
    int      streams_per_pixel;
    /* GDB200_GPT_* bits.  REF_UNINIT_MEASURE: gpt.cpp:957 default-constructs the DirectSamplingRecord of a reconnected
     * offset path that lands on an emitter and never sets .measure before Shape::pdfDirect (shape.cpp:116-126) reads it
     * (undefined behaviour).  By default the intended ESolidAngle is used; with this bit the tracer reproduces what a
     * g++ -O2 build of the reference does there (area emitters report density 0), which is how images are compared
     * bit for bit with the compiled reference.  FUSED_BOUNCE: the single-kernel bounce of round 1 (A/B measurements). */
    int      flags;
    int      max_slots;          /* resident path slots; 0 = default (streams beyond it are dealt out as slots drain) */
} gdb200_gpt_params;

#define GDB200_GPT_REF_UNINIT_MEASURE 1
#define GDB200_GPT_FUSED_BOUNCE       2

/* Host output buffers, each width*height*3 fp64, interleaved RGB; any may be NULL.
 * Developed like MultiFilm::developMulti (value * 1/weight, fmtconv.cpp:1036-1045). */
typedef struct gdb200_buffers {
    double *throughput, *dx, *dy, *direct, *preview_final;
} gdb200_buffers;

typedef struct gdb200_scene gdb200_scene;

int  gdb200_scene_create(const gdb200_scene_desc *desc, gdb200_scene **out_scene);
void gdb200_scene_destroy(gdb200_scene *scene);

/* Renders the five G-PT buffers (replaces GradientPathIntegrator::render's
 * scheduling of renderBlock, gpt.cpp:1397-1410 / 1220-1355).  Accumulators stay
 * resident on the device inside `scene`; `out` (optional) receives developed copies. */
int  gdb200_gpt_render(gdb200_scene *scene, const gdb200_gpt_params *params,
                       gdb200_buffers *out, gdb200_stats *stats);

/* Device pointers (fp32, w*h*3 interleaved) of the developed throughput, dx, dy
 * and direct buffers after gdb200_gpt_render: the solver inputs of gpt.cpp:1439-1442. */
int  gdb200_gpt_solver_inputs(gdb200_scene *scene, const float **d_dx, const float **d_dy,
                              const float **d_throughput, const float **d_direct);

/* Raw accumulators (value RGB + weight) for multi-GPU tile merging: 5 buffers
 * [final,throughput,dx,dy,direct] x height x width x 4 fp64, device pointer. */
int  gdb200_gpt_accumulators(gdb200_scene *scene, double **d_accum, size_t *bytes);
/* Re-develop after the accumulators were modified externally (tile merge). */
int  gdb200_gpt_develop(gdb200_scene *scene, gdb200_buffers *out);

/* Self-check: traces n_rays random rays through `scene` with and without the bounds-based candidate
 * selection of the intersection routine and counts answers that differ (must be 0). */
int  gdb200_debug_check_culling(gdb200_scene *scene, int n_rays, unsigned long long seed,
                                unsigned long long *out_mismatches, unsigned long long *out_hits);

/* sizeof() of the structs of this header as the library was compiled, in declaration order: stats, poisson_config,
 * camera, shape, material, emitter, envmap, scene_desc, gpt_params, buffers.  Bindings compare them with their own
 * layout at load time so that a stale library fails loudly instead of reading shifted fields.  Returns the count. */
int  gdb200_abi_sizes(int *out_sizes, int capacity);

/* The wavefront workspace (path-slot state and queues, ~1.6 KB per resident path) is scratch memory kept per device
 * and reused across scenes and renders; this frees it on every device (e.g. before handing the GPU to another library). */
void gdb200_release_workspace(void);

/* Asynchronous cancel (Integrator::cancel, integrator.h:77-84). */
void gdb200_cancel(gdb200_scene *scene);

#ifdef __cplusplus
}
#endif
#endif /* GDB200_H */
