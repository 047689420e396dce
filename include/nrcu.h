/*
 * nrcu.h — C ABI of the B200-native NRenderer path-tracing backend (libnrcuda.so).
 *
 * This is the drop-in boundary between host code that speaks NRenderer's plugin API
 * (a g++-compiled `RenderComponent` adapter, see nrenderer_b200/plugin/) and the
 * nvcc-compiled sm_100a kernels.  Plain pointers and sizes only; no C++/torch types.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference/code):
 *   nrcu_upload_scene   <- what `RenderComponent::render(SharedScene)` receives
 *                          (include/component/RenderComponent.hpp:12-18, include/scene/Scene.hpp:40-67)
 *                          + VertexTransformer::exec (components/acc_path_tracing/src/VertexTransformer.cpp:6-54)
 *                          + mesh flattening (components/simple_path_tracing/src/SimplePathTracer.cpp:57-78,
 *                            components/acc_path_tracing/include/BVH.hpp:34-60)
 *                          + BVHTree::build (components/acc_path_tracing/include/BVH.hpp:166-222)
 *   nrcu_render         <- RayCastRenderer::render        (components/ray_cast/src/RayCastRenderer.cpp:14-38)
 *                          SimplePathTracerRenderer::render (components/simple_path_tracing/src/SimplePathTracer.cpp:39-97)
 *                          AccPathTracerRenderer::render   (components/acc_path_tracing/src/AccPathTracer.cpp:41-80)
 *                          result layout = what Screen::set consumes (server/server/Screen.cpp:54-66)
 *   nrcu_render_accumulate / nrcu_resolve
 *                       <- the same render loop split at the "sum over samples" / "÷spp, sqrt gamma"
 *                          boundary (AccPathTracer.cpp:22-34) so that sample slices rendered on
 *                          several GPUs can be reduced in linear space before the gamma.
 *   nrcu_render_multi   <- the reference's row-striped std::thread pool (AccPathTracer.cpp:63-70: 16 threads,
 *                          `for i = off; i < h; i += step`) at box scale: one host thread per GPU, sample slices
 *                          instead of row stripes, partial frames combined in linear space before the gamma
 *   nrcu_trace_batch    <- closestHitObject (SimplePathTracer.cpp:104-129 brute force;
 *                          AccPathTracer.cpp:87-99 -> BVHTree::Intersect BVH.hpp:93-164)
 *
 * All functions return 0 on success, a non-zero nrcu_status otherwise; the message is
 * available from nrcu_last_error().  Nothing throws across this boundary.  There is no
 * CPU fallback: without a CUDA device nrcu_create fails with NRCU_ERR_NO_DEVICE.
 */
#ifndef NRCU_H
#define NRCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRCU_ABI_VERSION 2

typedef struct nrcu_ctx nrcu_ctx;

enum nrcu_status {
    NRCU_OK = 0,
    NRCU_ERR_NO_DEVICE = 1,   /* no CUDA device / driver */
    NRCU_ERR_CUDA = 2,        /* a CUDA runtime call or kernel failed */
    NRCU_ERR_INVALID = 3,     /* bad argument / malformed scene */
    NRCU_ERR_STATE = 4,       /* call order violated (e.g. render before upload) */
    NRCU_ERR_PEER = 5,        /* multi-device reduce failed: a peer's partial frame could be neither mapped nor copied */
    NRCU_ERR_OVERFLOW = 6     /* a fixed-capacity device structure overflowed - a traversal stack (a hit may have been missed)
                                 or the ray queue of the branching glass mode even at one sample per wave - so the frame
                                 would be wrong: it is reported, never returned as NRCU_OK */
};

/* Which reference component the context reproduces. */
enum nrcu_mode {
    NRCU_MODE_RAYCAST = 0,    /* components/ray_cast          ("RayCast")          */
    NRCU_MODE_SIMPLE  = 1,    /* components/simple_path_tracing ("SimplePathTracer") */
    NRCU_MODE_ACC     = 2     /* components/acc_path_tracing  ("AccPathTracer")     */
};

/* Node::Type, include/scene/Model.hpp:86-92 */
enum nrcu_node_type {
    NRCU_NODE_SPHERE = 0, NRCU_NODE_TRIANGLE = 1, NRCU_NODE_PLANE = 2, NRCU_NODE_MESH = 3
};

/* Ambient::Type, include/scene/Scene.hpp:31-34 */
enum nrcu_ambient_type { NRCU_AMBIENT_CONSTANT = 0, NRCU_AMBIENT_ENVIRONMENT_MAP = 1 };

/* Bits of nrcu_material.present: which properties the NRenderer::Material carried. */
enum nrcu_material_prop {
    NRCU_MP_DIFFUSE_COLOR  = 1u << 0,   /* "diffuseColor"  RGB   */
    NRCU_MP_SPECULAR_COLOR = 1u << 1,   /* "specularColor" RGB   */
    NRCU_MP_SPECULAR_EX    = 1u << 2,   /* "specularEx"    Float */
    NRCU_MP_ALBEDO         = 1u << 3,   /* "albedo"        RGB   */
    NRCU_MP_ETA_R          = 1u << 4,   /* "eta_r"         Vec3  */
    NRCU_MP_ETA_I          = 1u << 5,   /* "eta_i"         Vec3  */
    NRCU_MP_IOR            = 1u << 6,   /* "ior"           Float */
    NRCU_MP_ABSORBED       = 1u << 7,   /* "absorbed"      RGB   */
    NRCU_MP_ROUGHNESS      = 1u << 8,   /* "roughness"     Float */
    NRCU_MP_F0             = 1u << 9    /* "F0"            Float */
};

/*
 * POD mirror of NRenderer::Material (include/scene/Material.hpp:98-168) with the
 * property look-ups the reference shaders perform at construction already resolved
 * (ray_cast/src/shaders/{Lambertian,Phong}.cpp, acc_path_tracing/src/shaders/ *.cpp,
 * acc_path_tracing/include/shaders/{Conductor,Glass}.hpp).  A property that is absent
 * has its bit cleared in `present`; the library then applies the reference's default
 * where it has one ((1,1,1) colours, specularEx 1, roughness 0.2, F0 0.04) and ZERO where
 * the reference reads an uninitialised member (Conductor albedo/eta_r/eta_i, Glass
 * ior/absorbed).
 */
typedef struct nrcu_material {
    uint32_t type;
    uint32_t present;
    float diffuse_color[3];
    float specular_color[3];
    float specular_ex;
    float albedo[3];
    float eta_r[3];
    float eta_i[3];
    float ior;
    float absorbed[3];
    float roughness;
    float f0;
} nrcu_material;   /* 24 x 4 bytes */

/*
 * POD mirror of NRenderer::Scene (include/scene/Scene.hpp:40-67) in MODEL-LOCAL
 * coordinates, exactly as SceneBuilder hands it to a component: the library performs the
 * VertexTransformer step itself.  Material references are 0-based indices, -1 = invalid
 * Handle (geometry/vec.hpp:13-27).  All vectors are packed xyz fp32.
 */
typedef struct nrcu_scene {
    /* RenderOption (Scene.hpp:13-27) */
    uint32_t width, height, depth, samples_per_pixel;
    /* Camera (include/scene/Camera.hpp:13-48) */
    float cam_position[3], cam_up[3], cam_look_at[3];
    float cam_fov, cam_aperture, cam_focus_distance, cam_aspect;
    /* Ambient (Scene.hpp:29-38) */
    uint32_t ambient_type;
    float ambient_constant[3];
    int32_t ambient_environment_map;          /* texture index or -1 */
    /* models (Model.hpp:98-102): translation only; scale is stored by the reference but never applied */
    uint32_t n_models;
    const float* model_translation;           /* n_models x 3 */
    /* nodes (Model.hpp:84-96) */
    uint32_t n_nodes;
    const uint32_t* node_type;                /* nrcu_node_type */
    const uint32_t* node_entity;              /* index into the typed buffer */
    const uint32_t* node_model;               /* index into models */
    /* sphereBuffer (Model.hpp:23-28) */
    uint32_t n_spheres;
    const float* sphere_position;             /* n x 3 */
    const float* sphere_radius;               /* n */
    const int32_t* sphere_material;           /* n */
    /* triangleBuffer (Model.hpp:30-55) */
    uint32_t n_triangles;
    const float* triangle_vertices;           /* n x 9  (v1 v2 v3) */
    const float* triangle_normal;             /* n x 3  (as stored; not normalised) */
    const int32_t* triangle_material;         /* n */
    /* planeBuffer (Model.hpp:57-64): parallelogram position + a*u + b*v */
    uint32_t n_planes;
    const float* plane_normal;                /* n x 3 */
    const float* plane_position;              /* n x 3 */
    const float* plane_u;                     /* n x 3 */
    const float* plane_v;                     /* n x 3 */
    const int32_t* plane_material;            /* n */
    /* meshBuffer (Model.hpp:66-82): positions + positionIndices only (normals/uvs are never read) */
    uint32_t n_meshes;
    const uint32_t* mesh_vertex_offset;       /* n_meshes + 1, prefix offsets into mesh_positions (in vertices) */
    const uint32_t* mesh_index_offset;        /* n_meshes + 1, prefix offsets into mesh_indices (in indices)   */
    const float* mesh_positions;              /* total_vertices x 3 */
    const uint32_t* mesh_indices;             /* total_indices, mesh-local vertex indices, 3 per triangle */
    const int32_t* mesh_material;             /* n_meshes */
    /* materials */
    uint32_t n_materials;
    const nrcu_material* materials;
    /* pointLightBuffer / areaLightBuffer (include/scene/Light.hpp:36-49), world coordinates */
    uint32_t n_point_lights;
    const float* point_intensity;             /* n x 3 */
    const float* point_position;              /* n x 3 */
    uint32_t n_area_lights;
    const float* area_radiance;               /* n x 3 */
    const float* area_position;               /* n x 3 */
    const float* area_u;                      /* n x 3 */
    const float* area_v;                      /* n x 3 */
    /* textures (include/scene/Texture.hpp:12-39): RGBA fp32, row 0 first; only the ambient map is read */
    uint32_t n_textures;
    const uint32_t* texture_width;            /* n */
    const uint32_t* texture_height;           /* n */
    const uint64_t* texture_offset;           /* n, offset into texture_rgba in floats */
    const float* texture_rgba;
} nrcu_scene;

/* Glass (material type 2) handling in NRCU_MODE_ACC. */
enum nrcu_glass_mode {
    NRCU_GLASS_STOCHASTIC = 0, /* pick reflect/refract with probability F; same expectation as the
                                  reference's two-branch recursion (AccPathTracer.cpp:151-160) */
    NRCU_GLASS_BRANCH = 1      /* trace both branches like the reference (queue grows) */
};

/* nrcu_render_params.flags */
enum nrcu_render_flags {
    NRCU_FLAG_NEE = 1u << 0,   /* EXTENSION (not in the reference, whose area lights are only hit by chance): next-event
                                  estimation at Lambertian vertices - one shadow ray per diffuse bounce towards a uniformly
                                  sampled point of an area light; same expectation as the reference's estimator, far
                                  lower variance.  Ignored in RayCast mode. */
    NRCU_FLAG_ENV_IS = 1u << 1,/* EXTENSION to the environment-map extension: at Lambertian vertices one direction is
                                  drawn from the map's luminance x sin(theta) distribution (marginal/conditional CDF tables built at
                                  upload) and its shadow ray is combined with the hemisphere sample by the balance heuristic.  Same
                                  expectation as the plain miss lookup, far lower variance for maps with small bright regions.  Takes
                                  effect only when the scene's ambient is an environment map; it then replaces NRCU_FLAG_NEE's light
                                  sampling at those vertices (one shadow ray per vertex). */
    NRCU_FLAG_KERNEL_TIMES = 1u << 2 /* fill nrcu_stats.ms_trace / ms_shade / ms_stage2: two CUDA events around every kernel launch
                                  (~4 000 per cfg3 frame) and their read-back, about 1 % of a frame.  Without it a stats request costs
                                  one synchronisation: paths, rays, launches and ms_total only. */
};

/* How the (pixel, sample) paths are scheduled onto the GPU.  Both schedulers trace exactly the same paths (the RNG is
 * keyed by pixel, sample and bounce); only the fp32 order in which a pixel's samples are summed differs. */
enum nrcu_scheduler {
    NRCU_SCHED_AUTO  = 0,      /* NRCU_SCHED environment variable ("waves" / "regen"), else the library default */
    NRCU_SCHED_WAVES = 1,      /* per-bounce wavefront: k samples of every pixel per wave, queues compacted after every bounce,
                                  samples summed in sample order (the image does not depend on the wave size) */
    NRCU_SCHED_REGEN = 2       /* path regeneration: K slots per pixel, a slot whose path ends starts its next sample in place;
                                  no compaction, no per-bounce launches.  Samples are summed per slot, then over the K slots.
                                  Falls back to WAVES for NEE, the branching glass mode and depth 0 */
};

typedef struct nrcu_render_params {
    uint64_t seed;             /* counter-based RNG key; same seed => same image */
    uint32_t sample_begin;     /* global sample indices [sample_begin, sample_end) of samples_per_pixel */
    uint32_t sample_end;       /* 0,0 = all samples */
    uint32_t glass_mode;       /* nrcu_glass_mode */
    uint32_t samples_per_wave; /* WAVES: samples per wave, 0 = choose automatically; a non-zero value selects WAVES under AUTO.
                                  REGEN (explicitly selected): slots per pixel K, 0 = automatic */
    uint32_t flags;            /* nrcu_render_flags; 0 = the reference's estimator */
    uint32_t scheduler;        /* nrcu_scheduler */
} nrcu_render_params;

typedef struct nrcu_stats {
    uint64_t paths;            /* path samples rendered (w*h*samples in the slice) */
    uint64_t rays;             /* closest-hit queries + shadow rays answered (device counters): traced through the kernels, or - the
                                  camera rays of `dead_pixels` - answered by the film rectangles without a ray (they can hit nothing) */
    uint64_t kernel_launches;  /* launches of this library's kernels during the call */
    float ms_total;            /* CUDA-event time of the whole call on the context's stream */
    float ms_trace;            /* NRCU_FLAG_KERNEL_TIMES: time inside the closest-hit kernels: k_raygen (camera rays + fused stage 1), k_big, k_trace* */
    float ms_shade;            /* NRCU_FLAG_KERNEL_TIMES: time inside the shading kernels */
    float ms_setup;            /* scene preparation + BVH build at upload time */
    uint32_t bvh_nodes;        /* wide nodes */
    uint32_t n_primitives;     /* primitives after mesh flattening */
    uint32_t max_queue;        /* high-water mark of the ray queue (branching glass mode: the unclamped demand) */
    float ms_stage2;           /* NRCU_FLAG_KERNEL_TIMES: the part of ms_trace spent in the BVH traversal kernels (k_trace*) */
    uint32_t scheduler;        /* nrcu_scheduler that ran (WAVES or REGEN) */
    uint32_t iterations;       /* REGEN: stage-1/stage-2/shade rounds; WAVES: waves x bounces */
    uint32_t wave_retries;     /* branching glass mode: waves re-rendered with fewer samples because the queue overflowed */
    uint32_t dead_pixels;      /* pixels whose camera rays cannot meet a primitive's bounds or a light (pinhole camera, no environment
                                  map): no ray is generated for their samples; rays - dead_pixels * samples = rays traced */
} nrcu_stats;

/* --- lifetime ------------------------------------------------------------------------- */
int nrcu_abi_version(void);
int nrcu_device_count(void);
int nrcu_create(int device, nrcu_ctx** out);
int nrcu_destroy(nrcu_ctx* ctx);
const char* nrcu_last_error(const nrcu_ctx* ctx);    /* ctx may be NULL: last creation error */

/* --- scene ---------------------------------------------------------------------------- */
/* Copies everything it needs; the caller's arrays may be freed on return.  Performs the
 * world transform, mesh flattening, primitive-bounds and wide-BVH build on the device. */
int nrcu_upload_scene(nrcu_ctx* ctx, const nrcu_scene* scene, int mode);

/* Number of primitives after flattening, in the order closest-hit ids refer to:
 *   RAYCAST/SIMPLE: spheres, triangles (then flattened mesh triangles), planes   (buffer order)
 *   ACC:            scene.nodes order with each MESH node expanded in place        (BVH.hpp:34-60) */
int nrcu_primitive_count(const nrcu_ctx* ctx, uint32_t* out);

/* Read back the flattened world-space primitives (for parity tests of the scene-prep kernels).
 * kind[i]: nrcu_node_type of primitive i (mesh triangles report NRCU_NODE_MESH);
 * data[i*16..]: sphere: c.xyz r; triangle: v1 v2 v3 n (12); plane: n p u v (12);
 * material[i]: 0-based material index.  Any pointer may be NULL. */
int nrcu_download_primitives(const nrcu_ctx* ctx, uint32_t* kind, float* data16, int32_t* material);

/* --- rendering ------------------------------------------------------------------------ */
/* Whole frame: all samples, ÷spp, sqrt gamma, alpha 1, row 0 = top of the image; writes
 * width*height*4 floats to HOST memory `rgba_out` (what Screen::set takes).  stats may be NULL. */
int nrcu_render(nrcu_ctx* ctx, const nrcu_render_params* params, float* rgba_out, nrcu_stats* stats);

/* Sample slice: adds the LINEAR radiance sums of samples [sample_begin, sample_end) into
 * the DEVICE buffer `d_accum` (width*height*4 floats, rgb = sums, a = sample count), which
 * the caller owns (e.g. a torch tensor) and may reduce across GPUs. Asynchronous on the
 * context's stream unless stats != NULL. */
int nrcu_render_accumulate(nrcu_ctx* ctx, const nrcu_render_params* params, float* d_accum, nrcu_stats* stats);

/* Progressive form of nrcu_render (the GUI re-uploads the frame whenever Screen::isUpdated(),
 * app/src/ui/views/ScreenView.cpp:168-173, so a component may publish intermediate frames): after every
 * `samples_per_update` samples (0 = one wave) the frame resolved from the samples so far is copied to `rgba_out`
 * and `on_update(user, rgba_out, samples_done, samples_total)` is called on the calling thread; a non-zero return
 * stops the render early (the frame then holds samples_done samples).  The last update is the final frame. */
typedef int (*nrcu_update_fn)(void* user, const float* rgba, uint32_t samples_done, uint32_t samples_total);
int nrcu_render_progressive(nrcu_ctx* ctx, const nrcu_render_params* params, uint32_t samples_per_update,
                            float* rgba_out, nrcu_update_fn on_update, void* user, nrcu_stats* stats);

/* Whole frame on SEVERAL devices of one box (sample slices): ctxs[g] (one context per device, the
 * same scene uploaded to each in the same mode) renders samples [g*spp/n, (g+1)*spp/n) of every pixel on its own
 * host thread; ctxs[0] then sums the partial LINEAR frames straight out of its peers' HBM over NVLink (peer access;
 * staged copies when peer access is unavailable) inside the kernel that resolves (/ n, sqrt gamma), and the frame
 * is copied to HOST memory `rgba_out`.  With n_ctx == 1 this is nrcu_render.  stats (may be NULL): sums over the
 * devices; ms_* = maximum over the devices. */
int nrcu_render_multi(nrcu_ctx* const* ctxs, int n_ctx, const nrcu_render_params* params, float* rgba_out, nrcu_stats* stats);

/* Metropolis light transport (counterpart of the reference's MetropolisLightTransport component,
 * components/metropolis_light_transport/src/Metropolis.cpp:25-135): Kelemen-style Markov chains in primary sample space -
 * large steps with probability large_step_prob, the reference's exponential small-step perturbation, expected-value
 * accumulation, b from n_init independent samples - driving THIS backend's path sampler (the scene's own materials and
 * lights; the reference's sampler is bidirectional with hard-coded colours).  The frame has the expectation of nrcu_render's.
 * One chain per GPU thread; the scene must have been uploaded in NRCU_MODE_SIMPLE or NRCU_MODE_ACC, depth <= 32. */
enum nrcu_mlt_tone {
    NRCU_MLT_TONE_SQRT = 0,       /* sqrt, like the path tracers (AccPathTracer.cpp:14-16) */
    NRCU_MLT_TONE_REFERENCE = 1,  /* pow(1 - exp(-x), 1/2.2), the reference MLT's own (Metropolis.cpp:118-123) */
    NRCU_MLT_TONE_LINEAR = 2      /* none: linear radiance (what the tests compare with the path tracer's mean) */
};
typedef struct nrcu_mlt_params {
    uint64_t seed;
    uint32_t mutations_per_pixel; /* total mutations = this x width x height; 0 = the scene's samples_per_pixel */
    uint32_t chains;              /* 0 = automatic (at most 2^18, at least 64 mutations per chain) */
    uint32_t n_init;              /* samples that estimate b; 0 = 262144 (the reference uses 10000) */
    float large_step_prob;        /* 0 = the reference's 0.3 */
    uint32_t tone_map;            /* nrcu_mlt_tone */
    uint32_t reserved;
} nrcu_mlt_params;
/* stats: paths = mutations, rays as counted by the chains, max_queue = chains, iterations = mutations per chain,
 * wave_retries = accepted mutations / 1024. */
int nrcu_render_mlt(nrcu_ctx* ctx, const nrcu_mlt_params* params, float* rgba_out, nrcu_stats* stats);

/* d_rgba[p] = (sqrt(d_accum[p].rgb / d_accum[p].a), 1); both DEVICE pointers; may alias. */
int nrcu_resolve(nrcu_ctx* ctx, const float* d_accum, float* d_rgba);

/* Closest-hit parity probe: n rays (HOST arrays, 6 floats each: origin, direction) through the
 * same traversal kernel the renderer uses.  prim_id[i] = primitive id (see
 * nrcu_primitive_count) or -1; t[i] = hit distance or +inf. */
int nrcu_trace_batch(nrcu_ctx* ctx, const float* rays, uint32_t n, int32_t* prim_id, float* t);

/* Page-locked host memory for the frame a caller hands to nrcu_render* (the reference's components allocate their pixel
 * buffer with new[], ray_cast/src/RayCastRenderer.cpp:7-10; a pinned buffer makes the device -> host copy of the frame a
 * single DMA at PCIe speed instead of a staged copy through the driver's bounce buffer).  NULL when the allocation fails:
 * callers fall back to ordinary memory, which every entry point accepts as well. */
void* nrcu_host_alloc(size_t bytes);
void nrcu_host_free(void* p);

/* Stream plumbing for callers that own a CUDA stream (e.g. torch): cudaStream_t as void*. */
int nrcu_set_stream(nrcu_ctx* ctx, void* cuda_stream);
int nrcu_synchronize(nrcu_ctx* ctx);

/* Counter-based RNG known-answer probe: Philox4x32-10 block for (counter, key). */
void nrcu_philox4x32(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* NRCU_H */
