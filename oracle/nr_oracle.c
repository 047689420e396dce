/*
 * nr_oracle.c — CPU restatement of NRenderer's RayCast / SimplePathTracer / AccPathTracer hot path.
 *
 * TEST INFRASTRUCTURE (see nr_oracle.h).  Plain C, fp32 in the reference's operation order
 * (glm 0.9.9.9 as vendored under code/dependences/glm: dot = (x+y)+z func_geometric.inl,
 * normalize = v * (1/sqrt(dot)), inverse(mat3) by cofactors * 1/det func_matrix.inl,
 * mat3*vec3 type_mat3x3.inl:468-474).  Build with -ffp-contract=off: the reference build
 * (g++ -O2, x86-64, no -mfma) never fuses multiply-adds.
 *
 * Reference paths are relative to /root/reference/code/components.
 */
#include "nr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

typedef struct { float x, y, z; } v3;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vdiv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vdivs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 vadds(v3 a, float s) { return V(a.x + s, a.y + s, a.z + s); }
static inline v3 vsubs(v3 a, float s) { return V(a.x - s, a.y - s, a.z - s); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* glm compute_dot<vec3>: tmp = a*b; tmp.x + tmp.y + tmp.z */
static inline float vdot(v3 a, v3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
/* glm compute_cross */
static inline v3 vcross(v3 x, v3 y) { return V(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
/* glm compute_normalize: v * inversesqrt(dot(v,v)), inversesqrt = 1/sqrt */
static inline v3 vnormalize(v3 a) { float s = 1.0f / sqrtf(vdot(a, a)); return vscale(a, s); }
static inline float vlength(v3 a) { return sqrtf(vdot(a, a)); }
static inline float vget(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline v3 ld3(const float* p) { return V(p[0], p[1], p[2]); }
static inline void st3(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }

typedef struct { v3 o, d; } ray_t;
static inline v3 ray_at(ray_t r, float t) { return vadd(r.o, vscale(r.d, t)); } /* Ray.hpp:30-33 */

enum { K_SPHERE = 0, K_TRIANGLE = 1, K_PLANE = 2, K_MESH = 3 };

typedef struct {
    int kind;
    int material;
    /* sphere: c.xyz r | triangle: v1 v2 v3 n | plane: n p u v */
    float a[12];
    float bmin[3], bmax[3]; /* Bounds3 of the reference BVH leaf (ACC mode) */
} prim_t;

typedef struct {
    v3 position, lower_left, horizontal, vertical, u, v, w;
    float lens_radius;
} camera_t;

struct nro_scene {
    int mode;
    uint32_t width, height, depth, spp;
    uint32_t n_prims;
    prim_t* prims;
    uint32_t n_materials;
    nrcu_material* materials;
    uint32_t n_point; float* point_intensity; float* point_position;
    uint32_t n_area; float* area_radiance; float* area_position; float* area_u; float* area_v;
    camera_t cam;
    v3 ambient;
    int env_w, env_h; float* env_rgba; /* ambient environment map (our extension, A18) */
    float *env_sin, *env_row_cdf, *env_col_cdf; float env_total; /* importance-sampling tables (NRCU_FLAG_ENV_IS), built lazily */
    /* Microfacet Sampler(6) constants (acc_path_tracing/src/shaders/Microfacet.cpp:71-76) */
    float mf_u1, mf_u2;
};

static int g_threads = 0;
void nro_set_threads(int n) { g_threads = n; }

/* Minimal pthread parallel-for (dynamic chunks); the reference itself stripes rows over 16 std::threads
 * (AccPathTracer.cpp:63-70).  Results do not depend on the thread count: every item is independent. */
typedef void (*pf_body)(int64_t begin, int64_t end, void* ctx, int tid);
typedef struct { atomic_llong next; int64_t n, chunk; pf_body body; void* ctx; } pf_shared;
typedef struct { pf_shared* sh; int tid; } pf_arg;
static void* pf_worker(void* a_) {
    pf_arg* a = (pf_arg*)a_;
    for (;;) {
        int64_t b = atomic_fetch_add(&a->sh->next, a->sh->chunk);
        if (b >= a->sh->n) break;
        int64_t e = b + a->sh->chunk; if (e > a->sh->n) e = a->sh->n;
        a->sh->body(b, e, a->sh->ctx, a->tid);
    }
    return NULL;
}
#define PF_MAX_THREADS 256
static int pf_threads(void) {
    int t = g_threads;
    if (t <= 0) { long c = sysconf(_SC_NPROCESSORS_ONLN); t = c > 0 ? (int)c : 1; }
    if (t > PF_MAX_THREADS) t = PF_MAX_THREADS;
    return t;
}
static void parallel_for(int64_t n, int64_t chunk, pf_body body, void* ctx) {
    int nt = pf_threads();
    pf_shared sh; atomic_init(&sh.next, 0); sh.n = n; sh.chunk = chunk; sh.body = body; sh.ctx = ctx;
    if (nt <= 1 || n <= chunk) { pf_arg a = {&sh, 0}; pf_worker(&a); return; }
    pthread_t th[PF_MAX_THREADS]; pf_arg args[PF_MAX_THREADS];
    for (int i = 0; i < nt; i++) { args[i].sh = &sh; args[i].tid = i; pthread_create(&th[i], NULL, pf_worker, &args[i]); }
    for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).
 * -----------------------------------------------------------------------------------------*/
void nro_philox4x32(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4]) {
    uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3], k0 = k_in[0], k1 = k_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* 24-bit uniform in [0,1): what libstdc++'s generate_canonical<float,24> yields per draw. */
static inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

#define RNG_STREAM_CAMERA 0xFFFFFFFFu
static inline void rng_block(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stream, uint32_t block, uint32_t out[4]) {
    uint32_t c[4] = {pixel, sample, stream, block};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    nro_philox4x32(c, k, out);
}

/* ------------------------------------------------------------------------------------------
 * Scene preparation
 * -----------------------------------------------------------------------------------------*/
/* glm mat4*vec4 (type_mat4x4.inl:561-571): (m0*x + m1*y) + (m2*z + m3*w), w = 1, for the
 * matrices the reference uses: diag(s,s,s) + translation column t. */
static inline v3 xform(v3 p, float s, v3 t) {
    float x = (s * p.x + 0.0f * p.y) + (0.0f * p.z + t.x * 1.0f);
    float y = (0.0f * p.x + s * p.y) + (0.0f * p.z + t.y * 1.0f);
    float z = (0.0f * p.x + 0.0f * p.y) + (s * p.z + t.z * 1.0f);
    return V(x, y, z);
}

static inline float min3f(float a, float b, float c) { return fminf(a, fminf(b, c)); }
static inline float max3f(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

/* Bounds3 constructors, acc_path_tracing/include/Bounds3.hpp:35-103 */
static void prim_bounds(prim_t* p) {
    if (p->kind == K_SPHERE) {
        float r = p->a[3];
        for (int k = 0; k < 3; k++) { p->bmin[k] = p->a[k] - r; p->bmax[k] = p->a[k] + r; }
    } else if (p->kind == K_TRIANGLE || p->kind == K_MESH) {
        for (int k = 0; k < 3; k++) {
            p->bmin[k] = min3f(p->a[k], p->a[3 + k], p->a[6 + k]);
            p->bmax[k] = max3f(p->a[k], p->a[3 + k], p->a[6 + k]);
        }
    } else {
        float eps = 0.01f;
        v3 n = ld3(p->a), p1 = ld3(p->a + 3), u = ld3(p->a + 6), v = ld3(p->a + 9);
        v3 p2 = vadd(p1, u), p3 = vadd(p1, v), p4 = vadd(vadd(p1, u), v);
        v3 en = vscale(n, eps);
        p1 = vsub(p1, en); p2 = vsub(p2, en); p3 = vadd(p3, en); p4 = vadd(p4, en);
        for (int k = 0; k < 3; k++) {
            p->bmin[k] = fminf(vget(p1, k), fminf(vget(p2, k), fminf(vget(p3, k), vget(p4, k))));
            p->bmax[k] = fmaxf(vget(p1, k), fmaxf(vget(p2, k), fmaxf(vget(p3, k), vget(p4, k))));
        }
    }
}

/* Camera ctor: ray_cast/include/Camera.hpp:25-46 == acc_path_tracing/include/Camera.hpp:27-48 */
static void camera_init(camera_t* c, const nrcu_scene* s) {
    c->position = ld3(s->cam_position);
    c->lens_radius = s->cam_aperture / 2.f;
    float vfov = s->cam_fov;
    if (vfov > 160.f) vfov = 160.f; else if (vfov < 20.f) vfov = 20.f; /* clamp(x, max, min), geometry/vec.hpp:86-91 */
    float theta = vfov * 0.01745329251994329576923690768489f;      /* glm::radians */
    float half_h = tanf(theta / 2.f);
    float half_w = s->cam_aspect * half_h;
    v3 up = ld3(s->cam_up);
    c->w = vnormalize(vsub(c->position, ld3(s->cam_look_at)));
    c->u = vnormalize(vcross(up, c->w));
    c->v = vcross(c->w, c->u);
    float f = s->cam_focus_distance;
    /* position - halfWidth*focusDis*u - halfHeight*focusDis*v - focusDis*w */
    c->lower_left = vsub(vsub(vsub(c->position, vscale(c->u, half_w * f)), vscale(c->v, half_h * f)), vscale(c->w, f));
    c->horizontal = vscale(c->u, 2 * half_w * f);
    c->vertical = vscale(c->v, 2 * half_h * f);
}

static void mesh_triangle(prim_t* p, const float* pos, const uint32_t* idx, int material) {
    /* SimplePathTracer.cpp:65-74 / Bounds3.hpp:79-103: geometric normal from the winding */
    v3 v1 = ld3(pos + 3 * idx[0]), v2 = ld3(pos + 3 * idx[1]), v3_ = ld3(pos + 3 * idx[2]);
    p->kind = K_MESH; p->material = material;
    st3(p->a, v1); st3(p->a + 3, v2); st3(p->a + 6, v3_);
    st3(p->a + 9, vnormalize(vcross(vsub(v2, v1), vsub(v3_, v1))));
}

nro_scene* nro_prepare(const nrcu_scene* in, int mode) {
    nro_scene* s = (nro_scene*)calloc(1, sizeof(nro_scene));
    s->mode = mode;
    s->width = in->width; s->height = in->height; s->depth = in->depth; s->spp = in->samples_per_pixel;
    /* working copies the VertexTransformer mutates in place */
    float* sph = (float*)malloc(sizeof(float) * 3 * (in->n_spheres + 1));
    float* tri = (float*)malloc(sizeof(float) * 9 * (in->n_triangles + 1));
    float* pln = (float*)malloc(sizeof(float) * 3 * (in->n_planes + 1));
    uint32_t total_v = in->n_meshes ? in->mesh_vertex_offset[in->n_meshes] : 0;
    float* mpos = (float*)malloc(sizeof(float) * 3 * (total_v + 1));
    memcpy(sph, in->sphere_position, sizeof(float) * 3 * in->n_spheres);
    memcpy(tri, in->triangle_vertices, sizeof(float) * 9 * in->n_triangles);
    memcpy(pln, in->plane_position, sizeof(float) * 3 * in->n_planes);
    memcpy(mpos, in->mesh_positions, sizeof(float) * 3 * total_v);
    /* VertexTransformer::exec — ray_cast/src/VertexTransformer.cpp:6-27,
     * acc_path_tracing/src/VertexTransformer.cpp:6-54 (mesh branch :26-51, path tracers only) */
    for (uint32_t i = 0; i < in->n_nodes; i++) {
        v3 t = ld3(in->model_translation + 3 * in->node_model[i]);
        uint32_t e = in->node_entity[i];
        switch (in->node_type[i]) {
        case K_TRIANGLE: for (int k = 0; k < 3; k++) st3(tri + 9 * e + 3 * k, xform(ld3(tri + 9 * e + 3 * k), 1.0f, t)); break;
        case K_SPHERE: st3(sph + 3 * e, xform(ld3(sph + 3 * e), 1.0f, t)); break;
        case K_PLANE: st3(pln + 3 * e, xform(ld3(pln + 3 * e), 1.0f, t)); break;
        case K_MESH:
            if (mode != NRCU_MODE_RAYCAST) {
                for (uint32_t v = in->mesh_vertex_offset[e]; v < in->mesh_vertex_offset[e + 1]; v++)
                    st3(mpos + 3 * v, xform(ld3(mpos + 3 * v), 600.0f, V(40.f, -305.f, 920.f)));
            }
            break;
        }
    }
    uint32_t total_mesh_tris = 0;
    for (uint32_t i = 0; i < in->n_nodes; i++)
        if (in->node_type[i] == K_MESH && mode != NRCU_MODE_RAYCAST) {
            uint32_t e = in->node_entity[i];
            total_mesh_tris += (in->mesh_index_offset[e + 1] - in->mesh_index_offset[e]) / 3;
        }
    s->prims = (prim_t*)calloc(in->n_spheres + in->n_triangles + in->n_planes + total_mesh_tris + 1, sizeof(prim_t));
    uint32_t n = 0;
#define PUSH_SPHERE(e) do { prim_t* p = &s->prims[n++]; p->kind = K_SPHERE; p->material = in->sphere_material[e]; \
        memcpy(p->a, sph + 3 * (e), 12); p->a[3] = in->sphere_radius[e]; } while (0)
#define PUSH_TRIANGLE(e) do { prim_t* p = &s->prims[n++]; p->kind = K_TRIANGLE; p->material = in->triangle_material[e]; \
        memcpy(p->a, tri + 9 * (e), 36); memcpy(p->a + 9, in->triangle_normal + 3 * (e), 12); } while (0)
#define PUSH_PLANE(e) do { prim_t* p = &s->prims[n++]; p->kind = K_PLANE; p->material = in->plane_material[e]; \
        memcpy(p->a, in->plane_normal + 3 * (e), 12); memcpy(p->a + 3, pln + 3 * (e), 12); \
        memcpy(p->a + 6, in->plane_u + 3 * (e), 12); memcpy(p->a + 9, in->plane_v + 3 * (e), 12); } while (0)
#define PUSH_MESH(e) do { const uint32_t* ix = in->mesh_indices + in->mesh_index_offset[e]; \
        uint32_t nt = (in->mesh_index_offset[(e) + 1] - in->mesh_index_offset[e]) / 3; \
        for (uint32_t q = 0; q < nt; q++) mesh_triangle(&s->prims[n++], mpos + 3 * in->mesh_vertex_offset[e], ix + 3 * q, in->mesh_material[e]); } while (0)
    if (mode == NRCU_MODE_ACC) {
        /* BVHNode::buildBounds, acc_path_tracing/include/BVH.hpp:34-60: scene.nodes order, meshes expanded */
        for (uint32_t i = 0; i < in->n_nodes; i++) {
            uint32_t e = in->node_entity[i];
            switch (in->node_type[i]) {
            case K_SPHERE: PUSH_SPHERE(e); break;
            case K_TRIANGLE: PUSH_TRIANGLE(e); break;
            case K_PLANE: PUSH_PLANE(e); break;
            case K_MESH: PUSH_MESH(e); break;
            }
        }
    } else {
        /* closestHit loops the typed buffers: ray_cast/src/RayCastRenderer.cpp:66-91,
         * simple_path_tracing/src/SimplePathTracer.cpp:104-129 (mesh triangles appended :57-78) */
        for (uint32_t e = 0; e < in->n_spheres; e++) PUSH_SPHERE(e);
        for (uint32_t e = 0; e < in->n_triangles; e++) PUSH_TRIANGLE(e);
        if (mode == NRCU_MODE_SIMPLE)
            for (uint32_t i = 0; i < in->n_nodes; i++)
                if (in->node_type[i] == K_MESH) PUSH_MESH(in->node_entity[i]);
        for (uint32_t e = 0; e < in->n_planes; e++) PUSH_PLANE(e);
    }
    s->n_prims = n;
    for (uint32_t i = 0; i < n; i++) prim_bounds(&s->prims[i]);
    free(sph); free(tri); free(pln); free(mpos);

    s->n_materials = in->n_materials;
    s->materials = (nrcu_material*)malloc(sizeof(nrcu_material) * (in->n_materials + 1));
    memcpy(s->materials, in->materials, sizeof(nrcu_material) * in->n_materials);
    for (uint32_t i = 0; i < s->n_materials; i++) {
        nrcu_material* m = &s->materials[i];
        /* shader-constructor defaults: Lambertian.cpp:8-14, Phong.cpp:8-23, Microfacet.cpp:151-166;
         * Conductor.hpp:17-26 / Glass.hpp:16-22 leave absent members uninitialised -> zero here. */
        if (!(m->present & NRCU_MP_DIFFUSE_COLOR)) m->diffuse_color[0] = m->diffuse_color[1] = m->diffuse_color[2] = 1.f;
        if (!(m->present & NRCU_MP_SPECULAR_COLOR)) m->specular_color[0] = m->specular_color[1] = m->specular_color[2] = 1.f;
        if (!(m->present & NRCU_MP_SPECULAR_EX)) m->specular_ex = 1.f;
        if (m->type == 3) {
            if (!(m->present & NRCU_MP_ALBEDO)) m->albedo[0] = m->albedo[1] = m->albedo[2] = 1.f;
            if (!(m->present & NRCU_MP_ROUGHNESS)) m->roughness = 0.2f;
            if (!(m->present & NRCU_MP_F0)) m->f0 = 0.04;
        }
    }
#define DUP(dst, src, cnt) do { dst = (float*)malloc(sizeof(float) * 3 * ((cnt) + 1)); memcpy(dst, src, sizeof(float) * 3 * (cnt)); } while (0)
    s->n_point = in->n_point_lights; DUP(s->point_intensity, in->point_intensity, s->n_point); DUP(s->point_position, in->point_position, s->n_point);
    s->n_area = in->n_area_lights; DUP(s->area_radiance, in->area_radiance, s->n_area); DUP(s->area_position, in->area_position, s->n_area);
    DUP(s->area_u, in->area_u, s->n_area); DUP(s->area_v, in->area_v, s->n_area);
    camera_init(&s->cam, in);
    s->ambient = ld3(in->ambient_constant);
    if (in->ambient_type == NRCU_AMBIENT_ENVIRONMENT_MAP && in->ambient_environment_map >= 0 &&
        (uint32_t)in->ambient_environment_map < in->n_textures) {
        uint32_t ti = (uint32_t)in->ambient_environment_map;
        s->env_w = (int)in->texture_width[ti]; s->env_h = (int)in->texture_height[ti];
        size_t cnt = (size_t)s->env_w * s->env_h * 4;
        s->env_rgba = (float*)malloc(sizeof(float) * (cnt + 1));
        memcpy(s->env_rgba, in->texture_rgba + in->texture_offset[ti], sizeof(float) * cnt);
    }
    /* Sampler(6): minstd_rand seeded with 6, two uniform_real_distribution<float>(0,1) draws
     * (Microfacet.cpp:65-70).  minstd_rand: x <- 48271 x mod (2^31-1); generate_canonical<float,24>
     * = float(x - 1) / float(2147483646.0L). */
    {
        uint64_t x = 6; float range = (float)2147483646.0L;
        x = (48271u * x) % 2147483647u; s->mf_u1 = (float)(uint32_t)(x - 1) / range;
        x = (48271u * x) % 2147483647u; s->mf_u2 = (float)(uint32_t)(x - 1) / range;
    }
    return s;
}

void nro_free(nro_scene* s) {
    if (!s) return;
    free(s->prims); free(s->materials); free(s->point_intensity); free(s->point_position);
    free(s->area_radiance); free(s->area_position); free(s->area_u); free(s->area_v); free(s->env_rgba);
    free(s->env_sin); free(s->env_row_cdf); free(s->env_col_cdf);
    free(s);
}
uint32_t nro_primitive_count(const nro_scene* s) { return s->n_prims; }
void nro_get_primitives(const nro_scene* s, uint32_t* kind, float* data16, int32_t* material) {
    for (uint32_t i = 0; i < s->n_prims; i++) {
        if (kind) kind[i] = (uint32_t)s->prims[i].kind;
        if (material) material[i] = s->prims[i].material;
        if (data16) { memset(data16 + 16 * i, 0, 64); memcpy(data16 + 16 * i, s->prims[i].a, s->prims[i].kind == K_SPHERE ? 16 : 48); }
    }
}
void nro_get_bounds(const nro_scene* s, float* box6) {
    for (uint32_t i = 0; i < s->n_prims; i++) { memcpy(box6 + 6 * i, s->prims[i].bmin, 12); memcpy(box6 + 6 * i + 3, s->prims[i].bmax, 12); }
}
void nro_get_camera(const nro_scene* s, float* c, float* lens_radius) {
    st3(c, s->cam.position); st3(c + 3, s->cam.lower_left); st3(c + 6, s->cam.horizontal);
    st3(c + 9, s->cam.vertical); st3(c + 12, s->cam.u); st3(c + 15, s->cam.v);
    if (lens_radius) *lens_radius = s->cam.lens_radius;
}

/* ------------------------------------------------------------------------------------------
 * Intersections.  `rc` selects the RayCast variant (ray_cast/src/intersections/intersections.cpp:5-93:
 * normalised triangle/plane normals, exclusive t bounds t<=tMin) over the path-tracer variant
 * (acc_path_tracing/src/intersections/intersections.cpp:5-94 == simple_path_tracing: t<tMin).
 * -----------------------------------------------------------------------------------------*/
typedef struct { int hit; float t; v3 p, n; int material; } hit_t;

static hit_t x_triangle(ray_t ray, const float* a, int material, float tmin, float tmax, int rc) {
    hit_t h; h.hit = 0;
    v3 v1 = ld3(a), v2 = ld3(a + 3), v3_ = ld3(a + 6);
    v3 normal = rc ? vnormalize(ld3(a + 9)) : ld3(a + 9);
    v3 e1 = vsub(v2, v1), e2 = vsub(v3_, v1);
    v3 P = vcross(ray.d, e2);
    float det = vdot(e1, P);
    v3 T;
    if (det > 0) T = vsub(ray.o, v1); else { T = vsub(v1, ray.o); det = -det; }
    if (det < 0.000001f) return h;
    float u = vdot(T, P);
    if (u > det || u < 0.f) return h;
    v3 Q = vcross(T, e1);
    float v = vdot(ray.d, Q);
    if (v < 0.f || v + u > det) return h;
    float w = vdot(e2, Q);
    float inv_det = 1.f / det;
    w *= inv_det;
    if (rc) { if (w >= tmax || w <= tmin) return h; }
    else { if (w >= tmax || w < tmin) return h; }
    h.hit = 1; h.t = w; h.p = ray_at(ray, w); h.n = normal; h.material = material;
    return h;
}

static hit_t x_sphere(ray_t ray, const float* a, int material, float tmin, float tmax, int rc) {
    hit_t h; h.hit = 0;
    v3 position = ld3(a); float r = a[3];
    v3 oc = vsub(ray.o, position);
    float A = vdot(ray.d, ray.d);
    float b = vdot(oc, ray.d);
    float c = vdot(oc, oc) - r * r;
    float disc = b * b - A * c;
    float sq = sqrtf(disc);
    if (disc > 0) {
        float temp = (-b - sq) / A;
        int ok = rc ? (temp < tmax && temp > tmin) : (temp < tmax && temp >= tmin);
        if (!ok) { temp = (-b + sq) / A; ok = rc ? (temp < tmax && temp > tmin) : (temp < tmax && temp >= tmin); }
        if (ok) {
            h.hit = 1; h.t = temp; h.p = ray_at(ray, temp); h.n = vdivs(vsub(h.p, position), r); h.material = material;
        }
    }
    return h;
}

/* glm::inverse(mat3) restricted to what xPlane reads: the first two ROWS of inverse(mat3(u, v, cross(u,v))). */
static void plane_inverse_rows(v3 u, v3 v, float row0[3], float row1[3]) {
    v3 w = vcross(u, v);
    /* m[c][r]: m[0]=u, m[1]=v, m[2]=w */
    float m00 = u.x, m01 = u.y, m02 = u.z, m10 = v.x, m11 = v.y, m12 = v.z, m20 = w.x, m21 = w.y, m22 = w.z;
    float ood = 1.0f / (+m00 * (m11 * m22 - m21 * m12) - m10 * (m01 * m22 - m21 * m02) + m20 * (m01 * m12 - m11 * m02));
    row0[0] = +(m11 * m22 - m21 * m12) * ood; /* Inverse[0][0] */
    row0[1] = -(m10 * m22 - m20 * m12) * ood; /* Inverse[1][0] */
    row0[2] = +(m10 * m21 - m20 * m11) * ood; /* Inverse[2][0] */
    row1[0] = -(m01 * m22 - m21 * m02) * ood; /* Inverse[0][1] */
    row1[1] = +(m00 * m22 - m20 * m02) * ood; /* Inverse[1][1] */
    row1[2] = -(m00 * m21 - m20 * m01) * ood; /* Inverse[2][1] */
}

/* xPlane and xAreaLight share one body: `n` is normalize(p.normal) (RayCast), p.normal (PT) or cross(u,v) (light). */
static hit_t x_quad(ray_t ray, v3 n, v3 position, v3 u, v3 v, int material, float tmin, float tmax, int rc) {
    hit_t h; h.hit = 0;
    float nd = vdot(ray.d, n);
    if (nd < 0.0000001f && nd > -0.00000001f) return h;
    float dp = -vdot(position, n);
    float t = (-dp - vdot(n, ray.o)) / nd;
    if (rc) { if (t >= tmax || t <= tmin) return h; }
    else { if (t >= tmax || t < tmin) return h; }
    v3 hp = ray_at(ray, t);
    float r0[3], r1[3];
    plane_inverse_rows(u, v, r0, r1);
    v3 q = vsub(hp, position);
    float ru = r0[0] * q.x + r0[1] * q.y + r0[2] * q.z;
    float rv = r1[0] * q.x + r1[1] * q.y + r1[2] * q.z;
    if ((ru <= 1 && ru >= 0) && (rv <= 1 && rv >= 0)) { h.hit = 1; h.t = t; h.p = hp; h.n = n; h.material = material; }
    return h;
}

static hit_t x_prim(ray_t ray, const prim_t* p, float tmin, float tmax, int rc) {
    if (p->kind == K_SPHERE) return x_sphere(ray, p->a, p->material, tmin, tmax, rc);
    if (p->kind == K_PLANE) {
        v3 n = rc ? vnormalize(ld3(p->a)) : ld3(p->a);
        return x_quad(ray, n, ld3(p->a + 3), ld3(p->a + 6), ld3(p->a + 9), p->material, tmin, tmax, rc);
    }
    return x_triangle(ray, p->a, p->material, tmin, tmax, rc);
}

/* Bounds3::IntersectP, acc_path_tracing/include/Bounds3.hpp:141-168; invDir built as in
 * BVHTree::getIntersect, BVH.hpp:97 (double 1./d narrowed to float == 1.f/d, division is exact-rounded). */
static int bounds_intersectp(const float* bmin, const float* bmax, v3 o, v3 d) {
    if (o.x >= bmin[0] && o.x <= bmax[0] && o.y >= bmin[1] && o.y <= bmax[1] && o.z >= bmin[2] && o.z <= bmax[2]) return 1;
    float inv[3] = {(float)(1. / d.x), (float)(1. / d.y), (float)(1. / d.z)};
    float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    float tmn[3], tmx[3];
    for (int k = 0; k < 3; k++) {
        tmn[k] = (bmin[k] - oo[k]) * inv[k];
        tmx[k] = (bmax[k] - oo[k]) * inv[k];
        if (dd[k] < 0) { float tmp = tmn[k]; tmn[k] = tmx[k]; tmx[k] = tmp; }
    }
    /* std::max(a, b) = (a < b) ? b : a ; std::min(a, b) = (b < a) ? b : a */
#define SMAX(a, b) (((a) < (b)) ? (b) : (a))
#define SMIN(a, b) (((b) < (a)) ? (b) : (a))
    float in1 = SMAX(tmn[1], tmn[2]); float t_near = SMAX(tmn[0], in1);
    float in2 = SMIN(tmx[1], tmx[2]); float t_far = SMIN(tmx[0], in2);
    return (t_far >= 0 && t_near < t_far) ? 1 : 0;
}
int nro_bounds_intersectp(const float box6[6], const float origin[3], const float dir[3]) {
    return bounds_intersectp(box6, box6 + 3, ld3(origin), ld3(dir));
}

/* closestHit: RayCastRenderer.cpp:66-91 / SimplePathTracer.cpp:104-129 (shrinking tMax, strict <,
 * first in order wins) and AccPathTracer.cpp:87-99 -> BVH.hpp:93-164 (every leaf whose box passes
 * IntersectP is tested with tMax = inf; the minimum t wins).  second_t reports the runner-up for tie detection. */
static hit_t closest_hit(const nro_scene* s, ray_t ray, int* prim_id, float* second_t) {
    hit_t best; best.hit = 0; best.t = INFINITY;
    int rc = s->mode == NRCU_MODE_RAYCAST;
    float tmin = rc ? (float)0.01 : (float)0.000001;
    float second = INFINITY;
    int id = -1;
    for (uint32_t i = 0; i < s->n_prims; i++) {
        const prim_t* p = &s->prims[i];
        if (s->mode == NRCU_MODE_ACC) {
            if (!bounds_intersectp(p->bmin, p->bmax, ray.o, ray.d)) continue;
            hit_t h = x_prim(ray, p, tmin, INFINITY, 0);
            if (!h.hit) continue;
            if (h.t < best.t) { second = best.t; best = h; id = (int)i; }
            else if (h.t < second) second = h.t;
        } else {
            if (second_t) { /* tie probe only: does this primitive also hit at exactly best.t? */
                hit_t h2 = x_prim(ray, p, tmin, INFINITY, rc);
                if (h2.hit && best.hit && h2.t == best.t) second = best.t;
            }
            hit_t h = x_prim(ray, p, tmin, best.t, rc);
            if (h.hit && h.t < best.t) { best = h; id = (int)i; }
        }
    }
    if (prim_id) *prim_id = id;
    if (second_t) *second_t = second;
    return best;
}

typedef struct { const nro_scene* s; const float* rays; int32_t* prim_id; float* t; uint8_t* tie; } tb_ctx;
static void tb_body(int64_t b, int64_t e, void* c_, int tid) {
    tb_ctx* c = (tb_ctx*)c_; (void)tid;
    for (int64_t i = b; i < e; i++) {
        ray_t r; r.o = ld3(c->rays + 6 * i); r.d = ld3(c->rays + 6 * i + 3);
        int id; float second;
        hit_t h = closest_hit(c->s, r, &id, &second);
        if (c->prim_id) c->prim_id[i] = id;
        if (c->t) c->t[i] = h.hit ? h.t : INFINITY;
        if (c->tie) c->tie[i] = (h.hit && second == h.t) ? 1 : 0;
    }
}
void nro_trace_batch(const nro_scene* s, const float* rays, uint32_t n, int32_t* prim_id, float* t, uint8_t* tie) {
    tb_ctx c = {s, rays, prim_id, t, tie};
    parallel_for(n, 256, tb_body, &c);
}

/* ------------------------------------------------------------------------------------------
 * RayCast (A5, A6)
 * -----------------------------------------------------------------------------------------*/
static inline float clamp01(float x) { if (x > 1.f) return 1.f; if (x < 0.f) return 0.f; return x; } /* geometry/vec.hpp:86-91 */

static v3 raycast_shade(const nro_scene* s, int material, v3 in, v3 out, v3 normal) {
    const nrcu_material* m = &s->materials[material];
    v3 diffuse_color = ld3(m->diffuse_color);
    if (m->type == 1) {
        /* Phong::shade, ray_cast/src/shaders/Phong.cpp:5-7, 25-31 */
        v3 r = vsub(out, vscale(normal, 2 * vdot(out, normal)));
        v3 diffuse = vscale(diffuse_color, vdot(out, normal));
        v3 specular = vscale(ld3(m->specular_color), fabsf(powf(vdot(in, r), m->specular_ex)));
        return vadd(diffuse, specular);
    }
    /* Lambertian::shade, ray_cast/src/shaders/Lambertian.cpp:12-14 */
    return vscale(diffuse_color, vdot(out, normal));
}

/* RayCastRenderer::trace, ray_cast/src/RayCastRenderer.cpp:40-64 */
static v3 raycast_trace(const nro_scene* s, ray_t r) {
    if (s->n_point < 1) return V(0, 0, 0);
    v3 lpos = ld3(s->point_position), lint = ld3(s->point_intensity);
    hit_t h = closest_hit(s, r, NULL, NULL);
    if (!h.hit) return V(0, 0, 0);
    v3 out = vnormalize(vsub(lpos, h.p));
    if (vdot(out, h.n) < 0) return V(0, 0, 0);
    float distance = vlength(vsub(lpos, h.p));
    ray_t sr; sr.o = h.p; sr.d = out;
    hit_t sh = closest_hit(s, sr, NULL, NULL);
    v3 c = raycast_shade(s, h.material, vneg(r.d), out, h.n);
    if (!sh.hit || sh.t > distance) return vmul(c, lint);
    return V(0, 0, 0);
}

typedef struct { const nro_scene* s; float* rgba; } rc_ctx;
static void rc_body(int64_t b, int64_t e, void* c_, int tid) {
    rc_ctx* c = (rc_ctx*)c_; (void)tid;
    const nro_scene* s = c->s; float* rgba = c->rgba;
    int w = (int)s->width, hgt = (int)s->height;
    for (int i = (int)b; i < (int)e; i++) {
        for (int j = 0; j < w; j++) {
            /* Camera::shoot, ray_cast/include/Camera.hpp:49-57; pixel corner, no jitter (RayCastRenderer.cpp:27-35) */
            float sx = (float)j / (float)w, ty = (float)i / (float)hgt;
            ray_t r; r.o = s->cam.position;
            r.d = vnormalize(vsub(vadd(vadd(s->cam.lower_left, vscale(s->cam.horizontal, sx)), vscale(s->cam.vertical, ty)), s->cam.position));
            v3 c = raycast_trace(s, r);
            c = V(clamp01(c.x), clamp01(c.y), clamp01(c.z));
            c = V(sqrtf(c.x), sqrtf(c.y), sqrtf(c.z));
            float* px = rgba + 4 * ((size_t)(hgt - i - 1) * w + j);
            px[0] = c.x; px[1] = c.y; px[2] = c.z; px[3] = 1.f;
        }
    }
}
void nro_render_raycast(const nro_scene* s, float* rgba) {
    rc_ctx c = {s, rgba};
    parallel_for(s->height, 4, rc_body, &c);
}

/* ------------------------------------------------------------------------------------------
 * Path tracers (A7-A9, A13-A16)
 * -----------------------------------------------------------------------------------------*/
#define PT_PI 3.1415926535898f /* acc_path_tracing/include/shaders/Shader.hpp:17 */

/* closestHitLight, AccPathTracer.cpp:101-112 == SimplePathTracer.cpp:131-142 */
static float closest_light_which(const nro_scene* s, ray_t r, v3* radiance, int* which) {
    float closest = INFINITY;
    *radiance = V(0, 0, 0);
    if (which) *which = -1;
    for (uint32_t i = 0; i < s->n_area; i++) {
        v3 u = ld3(s->area_u + 3 * i), v = ld3(s->area_v + 3 * i);
        hit_t h = x_quad(r, vcross(u, v), ld3(s->area_position + 3 * i), u, v, -1, (float)0.000001, closest, 0);
        if (h.hit && closest > h.t) { closest = h.t; *radiance = ld3(s->area_radiance + 3 * i); if (which) *which = (int)i; }
    }
    return closest;
}
static float closest_light(const nro_scene* s, ray_t r, v3* radiance) { return closest_light_which(s, r, radiance, NULL); }

/* ---- next-event estimation: an EXTENSION of this backend (NRCU_FLAG_NEE), not in the reference, whose area lights are
 * only hit by chance (SimplePathTracer.cpp:144-177 has no light sampling).  Restated here independently of the CUDA
 * sources so that the kernels are checked against something other than themselves: at a Lambertian vertex one point of
 * one area light is sampled uniformly (light by e1, position by (frac(e1 * n), e2)); the shadow ray carries
 * f * Le / (p_light + p_hemisphere) (balance heuristic, solid-angle pdfs), and a hemisphere sample that then finds a
 * light is weighted by p_hemisphere / (p_hemisphere + p_light).  Visibility uses the same closest-hit / closest-light
 * queries as any other ray. */
#define PDF_HEMISPHERE (1.0f / (2.0f * PT_PI))
static int nee_sample(const nro_scene* s, v3 albedo, v3 hit_point, v3 normal, v3 thr, float e1, float e2, ray_t* shadow, v3* contrib, int* light) {
    if (s->n_area == 0) return 0;
    float fl = e1 * (float)s->n_area;
    int li = (int)fl; if (li > (int)s->n_area - 1) li = (int)s->n_area - 1;
    float a = fl - (float)li;
    v3 u = ld3(s->area_u + 3 * li), v = ld3(s->area_v + 3 * li), nl = vcross(u, v);
    v3 y = vadd(vadd(ld3(s->area_position + 3 * li), vscale(u, a)), vscale(v, e2));
    v3 wv = vsub(y, hit_point);
    float r2 = vdot(wv, wv);
    if (!(r2 > 0.f)) return 0;
    v3 w = vscale(wv, 1.0f / sqrtf(r2));
    float cos_s = vdot(normal, w);
    float cl = fabsf(vdot(nl, w));
    if (!(cos_s > 0.f) || !(cl > 0.f)) return 0;
    float p_l = r2 / (cl * (float)s->n_area);
    v3 f = vscale(vdivs(albedo, PT_PI), cos_s);
    *contrib = vscale(vmul(vmul(thr, f), ld3(s->area_radiance + 3 * li)), 1.0f / (p_l + PDF_HEMISPHERE));
    if (contrib->x == 0.f && contrib->y == 0.f && contrib->z == 0.f) return 0;
    shadow->o = hit_point; shadow->d = w; *light = li;
    return 1;
}
static float mis_light_weight(const nro_scene* s, ray_t ray, int which, float tl) {
    v3 nl = vcross(ld3(s->area_u + 3 * which), ld3(s->area_v + 3 * which));
    float cl = fabsf(vdot(nl, ray.d));
    if (!(cl > 0.f)) return 1.f;
    float p_l = tl * tl * vdot(ray.d, ray.d) / (cl * (float)s->n_area);
    return PDF_HEMISPHERE / (PDF_HEMISPHERE + p_l);
}

/* Camera::shoot, acc_path_tracing/include/Camera.hpp:51-63, with UniformInSquare jitter
 * (AccPathTracer.cpp:23-29; U(-1,1)^2) and UniformInCircle lens sample (UniformInCircle.hpp:20-26,
 * accept iff x*2 + y*2 <= 1 (sic)).  The lens draw is skipped when lensRadius == 0: its value is
 * multiplied by zero in the reference, only the stream position would differ. */
static ray_t camera_ray(const nro_scene* s, uint64_t seed, uint32_t pixel, uint32_t sample) {
    uint32_t rn[4];
    rng_block(seed, pixel, sample, RNG_STREAM_CAMERA, 0, rn);
    float rx = 2.f * u01(rn[0]) - 1.f, ry = 2.f * u01(rn[1]) - 1.f;
    int w = (int)s->width, h = (int)s->height;
    int row = (int)(pixel / (uint32_t)w), j = (int)(pixel % (uint32_t)w), i = h - 1 - row;
    float x = ((float)j + rx) / (float)w;
    float y = ((float)i + ry) / (float)h;
    v3 offset = V(0, 0, 0);
    if (s->cam.lens_radius != 0.f) {
        float lx = 0, ly = 0;
        for (uint32_t blk = 1;; blk++) {
            rng_block(seed, pixel, sample, RNG_STREAM_CAMERA, blk, rn);
            lx = 2.f * u01(rn[0]) - 1.f; ly = 2.f * u01(rn[1]) - 1.f;
            if (!((lx * 2 + ly * 2) > 1)) break;
            lx = 2.f * u01(rn[2]) - 1.f; ly = 2.f * u01(rn[3]) - 1.f;
            if (!((lx * 2 + ly * 2) > 1)) break;
        }
        offset = vadd(vscale(s->cam.u, lx * s->cam.lens_radius), vscale(s->cam.v, ly * s->cam.lens_radius));
    }
    ray_t r;
    r.o = vadd(s->cam.position, offset);
    r.d = vnormalize(vsub(vsub(vadd(vadd(s->cam.lower_left, vscale(s->cam.horizontal, x)), vscale(s->cam.vertical, y)), s->cam.position), offset));
    return r;
}
void nro_camera_ray(const nro_scene* s, uint64_t seed, uint32_t pixel, uint32_t sample, float* out6) {
    ray_t r = camera_ray(s, seed, pixel, sample); st3(out6, r.o); st3(out6 + 3, r.d);
}

typedef struct { ray_t ray; v3 attenuation; float pdf; int kind; /* 0 continue w/ factor, 1 glass, 2 dead */
                 ray_t reflex, refraction; v3 reflex_rate, refraction_rate; } scatter_t;

/* sin and cos of an angle in [0, 2*pi].  The reference calls libm's cosf/sinf (Hemisphere.hpp:28-29), whose last bit differs
 * between C libraries (and between any of them and CUDA's), which is enough to flip a tMin = 1e-6 self-intersection decision
 * and send a path down another branch.  This restatement - the device code and its host emulation evaluate the very same
 * sequence - is a fixed fp32 evaluation: three-constant Cody-Waite reduction by pi/2 and the Cephes sinf / cosf polynomials,
 * every product-sum an explicit fused multiply-add (identical on any IEEE machine).  Checked against double sin / cos over
 * ALL 1 086 918 636 floats of [0, 2 pi]: maximum absolute error 9.3e-8, maximum relative error 1.09 ulp, 98.8 % of the results
 * are the correctly rounded value - libm class, invisible to every statistical comparison with the real reference. */
static void sincos_det(float a, float* s, float* c) {
    float kf = rintf(a * 0.63661977236758134308f);              /* 2/pi */
    float r = fmaf(-kf, 1.5703125f, a);                          /* pi/2 = 1.5703125 + 4.8375...e-4 + 7.5497...e-8 */
    r = fmaf(-kf, 4.837512969970703125e-4f, r);
    r = fmaf(-kf, 7.54978995489188216e-8f, r);
    float z = r * r;
    float sp = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    float sr = fmaf(sp * z, r, r);
    float cp = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    float cr = fmaf(cp, z * z, fmaf(-0.5f, z, 1.0f));
    int k = ((int)kf) & 3;
    *s = (k == 0) ? sr : (k == 1) ? cr : (k == 2) ? -sr : -cr;
    *c = (k == 0) ? cr : (k == 1) ? -sr : (k == 2) ? -cr : sr;
}

/* Lambertian::shade, acc_path_tracing/src/shaders/Lambertian.cpp:16-34; HemiSphere::sample3d
 * (samplers/Hemisphere.hpp:24-32); Onb (include/Onb.hpp:17-27).  Returns the throughput factor
 * attenuation * dot(N, dir) / pdf of AccPathTracer.cpp:142 in *factor. */
static ray_t shade_lambertian(v3 albedo, v3 hit_point, v3 normal, float e1, float e2, v3* factor) {
    const float C_PI = 3.14159265358979323846264338327950288f;
    float r = sqrtf(1 - e1 * e1);
    float sn, cs;
    sincos_det(2 * C_PI * e2, &sn, &cs);
    float x = cs * r;
    float y = sn * r;
    float z = e1;
    v3 w = normal;
    v3 a = ((double)fabsf(w.x) > 0.9) ? V(0, 1, 0) : V(1, 0, 0);
    v3 v = vnormalize(vcross(w, a));
    v3 u = vcross(w, v);
    v3 local = vadd(vadd(vscale(u, x), vscale(v, y)), vscale(w, z));
    ray_t out; out.o = hit_point; out.d = vnormalize(local);
    float pdf = 1 / (2 * PT_PI);
    v3 attenuation = vdivs(albedo, PT_PI);
    float n_dot_in = vdot(normal, out.d);
    *factor = vdivs(vscale(attenuation, n_dot_in), pdf);
    return out;
}

/* Conductor::shade, acc_path_tracing/src/shaders/Conductor.cpp:6-42 */
static ray_t shade_conductor(const nrcu_material* m, ray_t ray, v3 hit_point, v3 normal, v3* factor) {
    v3 Vv = vneg(ray.d);
    v3 N = vnormalize(normal);
    v3 L = vnormalize(vadd(vneg(Vv), vscale(N, 2.f * vdot(Vv, N))));
    float cos_l = fabsf(vdot(L, N));
    float cos2 = cos_l * cos_l, sin2 = 1 - cos2, sin4 = sin2 * sin2;
    v3 eta_r = ld3(m->eta_r), eta_i = ld3(m->eta_i), albedo = ld3(m->albedo);
    v3 temp1 = vsubs(vsub(vmul(eta_r, eta_r), vmul(eta_i, eta_i)), sin2);
    v3 a2pb2 = vadd(vmul(temp1, temp1), vmul(vmul(vmul(vscale(eta_i, 4.0f), eta_i), eta_r), eta_r));
    a2pb2 = V(sqrtf(fmaxf(0.0f, a2pb2.x)), sqrtf(fmaxf(0.0f, a2pb2.y)), sqrtf(fmaxf(0.0f, a2pb2.z)));
    v3 a = vscale(vadd(a2pb2, temp1), 0.5f);
    a = V(sqrtf(fmaxf(0.f, a.x)), sqrtf(fmaxf(0.f, a.y)), sqrtf(fmaxf(0.f, a.z)));
    v3 term1 = vadds(a2pb2, cos2), term2 = vscale(a, 2.f * cos_l);
    v3 term3 = vadds(vscale(a2pb2, cos2), sin4), term4 = vscale(term2, sin2);
    v3 rs = vdiv(vsub(term1, term2), vadd(term1, term2));
    v3 rp = vdiv(vmul(rs, vsub(term3, term4)), vadd(term3, term4));
    v3 F = vscale(vadd(rs, rp), 0.5f);
    *factor = vmul(vscale(F, fabsf(vdot(L, N))), albedo);
    ray_t out; out.o = hit_point; out.d = L;
    return out;
}

/* Glass::shade, acc_path_tracing/src/shaders/Glass.cpp:15-57 */
static void shade_glass(const nrcu_material* m, ray_t ray, v3 hit_point, v3 normal, scatter_t* sc) {
    v3 absorbed = ld3(m->absorbed);
    float ior = m->ior;
    v3 N = vnormalize(normal), Vv = vnormalize(ray.d);
    float ior_inverse = ior;
    if (vdot(Vv, N) > 0.f) { N = vneg(N); ior_inverse = 1 / ior; }
    v3 reflex = vnormalize(vadd(Vv, vscale(vscale(N, 2.f), vdot(vneg(Vv), N))));
    float n12 = (ior_inverse - 1.f) / (ior_inverse + 1.f);
    n12 = n12 * n12;
    float vdn = fabsf(vdot(Vv, N));
    float p5 = (float)pow((double)(1 - vdn), 5.0);
    float Fs = n12 + (1.f - n12) * p5;
    v3 F = V(Fs, Fs, Fs);
    v3 reflex_rate = vmul(F, absorbed);
    v3 refraction_rate = vmul(V(1.f - Fs, 1.f - Fs, 1.f - Fs), absorbed);
    v3 x = vnormalize(vadd(reflex, Vv));
    v3 y = vnormalize(vneg(N));
    float x_ = (float)(sqrt(pow((double)(1 - fabsf(vdot(Vv, N))), 2.0)) / (double)ior_inverse);
    float y_ = (float)sqrt(1 - pow((double)x_, 2.0));
    v3 refraction = vnormalize(vadd(vscale(x, x_), vscale(y, y_)));
    if (x_ > 1.f) { reflex = absorbed; refraction_rate = V(0, 0, 0); refraction = V(0, 0, 0); }
    sc->reflex.o = hit_point; sc->reflex.d = reflex; sc->reflex_rate = reflex_rate;
    sc->refraction.o = hit_point; sc->refraction.d = refraction; sc->refraction_rate = refraction_rate;
}

/* SmithG1, Microfacet.cpp:16-32 */
static float smith_g1(v3 v, v3 h, v3 n, float roughness) {
    double cos_v_n = vdot(v, n);
    if (cos_v_n * vdot(v, h) <= 0.0f) return 0.f;
    if (fabs(cos_v_n - 1.0) < DBL_EPSILON) return 1.0f;
    float c2 = (float)pow(cos_v_n, 2.0), t2 = (1.0f - c2) / c2, a2 = roughness * roughness;
    return 2.0f / (1.0f + sqrtf(1.0f + a2 * t2));
}

/* Microfacet::shade + Sample/ToWorld/CoordinateSystem, Microfacet.cpp:71-118, 172-222 */
static int shade_microfacet(const nro_scene* s, const nrcu_material* m, ray_t ray, v3 hit_point, v3 normal, ray_t* out, v3* factor) {
    const float metalness = 0.2f;
    float roughness = m->roughness, F0 = m->f0;
    v3 albedo = ld3(m->albedo);
    v3 N = vnormalize(normal);
    /* Sample(N, roughness, &H, &D) */
    float phi = 2.0f * PT_PI * s->mf_u2;
    float cos_phi = cosf(phi), sin_phi = sinf(phi);
    float alpha_2 = roughness * roughness;
    float tan_theta_2 = alpha_2 * s->mf_u1 / (1.0f - s->mf_u1);
    float cos_theta = (float)(1.0 / sqrtf(1.0f + tan_theta_2));
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    v3 dir = V(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta);
    v3 B, C;
    if (fabsf(N.x) > fabsf(N.y)) {
        double len_inv = 1.0f / sqrtf(N.x * N.x + N.z * N.z);
        C = V((float)(N.z * len_inv), 0.f, (float)(-N.x * len_inv));
    } else {
        double len_inv = 1.0f / sqrtf(N.y * N.y + N.z * N.z);
        C = V(0.f, (float)(N.z * len_inv), (float)(-N.y * len_inv));
    }
    B = vcross(C, N);
    v3 H = vnormalize(vadd(vadd(vscale(B, dir.x), vscale(C, dir.y)), vscale(N, dir.z)));
    float D = 1.0f / (float)((double)(PT_PI * alpha_2) * pow((double)cos_theta, 3.0) * pow((double)(1.0f + tan_theta_2 / alpha_2), 2.0));
    H = vnormalize(H);
    v3 Vv = vneg(ray.d);
    float pdf = D * fabsf(1.0f / (4.0f * vdot(ray.d, H)));
    v3 L = vnormalize(vnormalize(vsub(ray.d, vscale(H, 2.0f * vdot(ray.d, H)))));
    float cos_theta_i = vdot(L, N);
    if (pdf == 0.f || vdot(ray.d, normal) >= 0.f || cos_theta_i <= 0.f) return 0;
    v3 specularF0 = vadd(vscale(V(F0, F0, F0), 1.f - metalness), vscale(albedo, metalness));
    float p5 = (float)pow((double)(1.0f - fabsf(vdot(L, H))), 5.0);
    v3 F = vadd(specularF0, vscale(vsub(V(1.0f, 1.0f, 1.0f), specularF0), p5));
    float cos_theta_o = fabsf(vdot(N, Vv));
    float G = smith_g1(L, H, N, roughness) * smith_g1(Vv, H, N, roughness);
    v3 att = vdivs(vscale(vscale(F, G), D), fabsf(4.0f * cos_theta_o));
    att = vdivs(att, pdf);
    att = vmul(att, albedo);
    out->o = hit_point; out->d = L;
    *factor = att;
    return 1;
}

/* Our extension (A18): latitude-longitude lookup of the ambient environment map on a miss. */
static v3 env_lookup(const nro_scene* s, v3 d) {
    if (!(vdot(d, d) > 0.f) || !(vdot(d, d) < INFINITY)) return V(0, 0, 0); /* zero / NaN direction: black */
    v3 n = vnormalize(d);
    float u = 0.5f + atan2f(n.x, n.z) * (0.5f / PT_PI);
    float cy = n.y; if (cy > 1.f) cy = 1.f; if (cy < -1.f) cy = -1.f;
    float v = acosf(cy) * (1.0f / PT_PI);
    int x = (int)(u * (float)s->env_w), y = (int)(v * (float)s->env_h);
    if (x < 0) x = 0; if (x > s->env_w - 1) x = s->env_w - 1;
    if (y < 0) y = 0; if (y > s->env_h - 1) y = s->env_h - 1;
    const float* px = s->env_rgba + 4 * ((size_t)y * s->env_w + x);
    return V(px[0], px[1], px[2]);
}

/* ---- environment-map importance sampling: an EXTENSION of the environment-map extension (NRCU_FLAG_ENV_IS), restated here
 * independently of the CUDA sources.  The map is piecewise constant, so a texel is drawn with probability proportional to
 * (r+g+b) * sin(theta at the row centre) - marginal CDF over the rows, conditional CDF inside the row - and the direction
 * uniformly inside the texel; solid-angle pdf = sin_row * lum * w * h / (total * 2 pi^2 * sin(theta of the direction)).
 * All sums are sequential fp32, sin via the shared sincos_det, so the tables equal the device's bit for bit. */
static void sincos_det(float a, float* s, float* c);
static void env_build_tables(nro_scene* s) {
    const float C_PI = 3.14159265358979323846264338327950288f;
    int w = s->env_w, h = s->env_h;
    s->env_sin = (float*)malloc(sizeof(float) * h); s->env_row_cdf = (float*)malloc(sizeof(float) * h);
    s->env_col_cdf = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        float sn, cs; sincos_det(C_PI * ((float)y + 0.5f) / (float)h, &sn, &cs);
        s->env_sin[y] = sn;
        float* cdf = s->env_col_cdf + (size_t)y * w;
        float sum = 0.f;
        for (int x = 0; x < w; x++) { const float* px = s->env_rgba + 4 * ((size_t)y * w + x); sum += (px[0] + px[1]) + px[2]; cdf[x] = sum; }
        for (int x = 0; x < w; x++) cdf[x] = sum > 0.f ? cdf[x] / sum : (float)(x + 1) / (float)w;
        cdf[w - 1] = 1.f;
        s->env_row_cdf[y] = sn * sum;
    }
    float total = 0.f;
    for (int y = 0; y < h; y++) { total += s->env_row_cdf[y]; s->env_row_cdf[y] = total; }
    for (int y = 0; y < h; y++) s->env_row_cdf[y] = total > 0.f ? s->env_row_cdf[y] / total : (float)(y + 1) / (float)h;
    s->env_row_cdf[h - 1] = 1.f;
    s->env_total = total;
}
static int cdf_pick(const float* cdf, int n, float e, float* frac) {
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (e < cdf[mid]) hi = mid; else lo = mid + 1; }
    float c0 = lo ? cdf[lo - 1] : 0.f, c1 = cdf[lo];
    *frac = c1 > c0 ? (e - c0) / (c1 - c0) : 0.5f;
    if (!(*frac >= 0.f)) *frac = 0.f;
    if (*frac > 0.99999994f) *frac = 0.99999994f;
    return lo;
}
static float env_pdf(const nro_scene* s, int x, int y, v3 n) {
    float st = sqrtf(fmaxf(0.f, 1.f - n.y * n.y));
    if (!(st > 0.f) || !(s->env_total > 0.f)) return 0.f;
    const float* px = s->env_rgba + 4 * ((size_t)y * s->env_w + x);
    float lum = (px[0] + px[1]) + px[2];
    return s->env_sin[y] * lum * (float)s->env_w * (float)s->env_h / (s->env_total * (2.f * PT_PI * PT_PI) * st);
}
static int env_sample(const nro_scene* s, v3 albedo, v3 hit_point, v3 normal, v3 thr, float e1, float e2, ray_t* shadow, v3* contrib) {
    const float C_PI = 3.14159265358979323846264338327950288f;
    int w = s->env_w, h = s->env_h;
    float jy, jx;
    int y = cdf_pick(s->env_row_cdf, h, e1, &jy);
    int x = cdf_pick(s->env_col_cdf + (size_t)y * w, w, e2, &jx);
    float u = ((float)x + jx) / (float)w, v = ((float)y + jy) / (float)h;
    float sp, cp, st, ct;
    sincos_det(2.f * C_PI * u, &sp, &cp);
    sincos_det(C_PI * v, &st, &ct);
    v3 d = V(st * -sp, ct, st * -cp);
    float cos_s = vdot(normal, d);
    if (!(cos_s > 0.f)) return 0;
    float p_env = env_pdf(s, x, y, d);
    if (!(p_env > 0.f)) return 0;
    const float* px = s->env_rgba + 4 * ((size_t)y * w + x);
    v3 f = vscale(vdivs(albedo, PT_PI), cos_s);
    *contrib = vscale(vmul(vmul(thr, f), V(px[0], px[1], px[2])), 1.0f / (p_env + PDF_HEMISPHERE));
    if (contrib->x == 0.f && contrib->y == 0.f && contrib->z == 0.f) return 0;
    shadow->o = hit_point; shadow->d = d;
    return 1;
}
static float mis_env_weight(const nro_scene* s, v3 d) {
    if (!(vdot(d, d) > 0.f) || !(vdot(d, d) < INFINITY)) return 1.f;
    v3 n = vnormalize(d);
    float u = 0.5f + atan2f(n.x, n.z) * (0.5f / PT_PI);
    float cy = n.y; if (cy > 1.f) cy = 1.f; if (cy < -1.f) cy = -1.f;
    float v = acosf(cy) * (1.0f / PT_PI);
    int x = (int)(u * (float)s->env_w), y = (int)(v * (float)s->env_h);
    if (x < 0) x = 0; if (x > s->env_w - 1) x = s->env_w - 1;
    if (y < 0) y = 0; if (y > s->env_h - 1) y = s->env_h - 1;
    return PDF_HEMISPHERE / (PDF_HEMISPHERE + env_pdf(s, x, y, n));
}

/* trace(): AccPathTracer.cpp:121-181 / SimplePathTracer.cpp:144-177, restated as a forward
 * throughput loop (the recursion is a product of per-bounce factors) with a small explicit stack
 * for the glass two-branch case.  branch_bits identifies the branch for the RNG. */
typedef struct { ray_t ray; v3 thr; uint32_t depth; uint32_t branch; } work_t;

static v3 pt_trace(const nro_scene* s, uint64_t seed, uint32_t pixel, uint32_t sample, ray_t ray0, int glass_mode, int nee, uint64_t* rays) {
    v3 L = V(0, 0, 0);
    work_t stack[256];
    int sp = 0;
    if (nee == 1 && s->n_area == 0) nee = 0;
    stack[sp].ray = ray0; stack[sp].thr = V(1, 1, 1); stack[sp].depth = 0; stack[sp].branch = 0; sp++;
    while (sp > 0) {
        work_t wk = stack[--sp];
        ray_t ray = wk.ray; v3 thr = wk.thr; uint32_t branch = wk.branch;
        int mis = 0;   /* the previous vertex also sampled the lights directly: a light found now gets the balance-heuristic weight */
        for (uint32_t d = wk.depth;; d++) {
            if (d == s->depth) { L = vadd(L, vmul(thr, s->ambient)); break; }
            (*rays)++;
            hit_t h = closest_hit(s, ray, NULL, NULL);
            v3 radiance;
            int which = -1;
            float tl = closest_light_which(s, ray, &radiance, &which);
            const int weigh_light = mis;
            mis = 0;
            if (h.hit && h.t < tl) {
                const nrcu_material* m = &s->materials[h.material];
                uint32_t type = s->mode == NRCU_MODE_ACC ? m->type : 0;
                uint32_t rn[4];
                rng_block(seed, pixel, sample, d, branch, rn);
                if (type == 2) {
                    scatter_t sc; shade_glass(m, ray, h.p, h.n, &sc);
                    if (glass_mode == NRCU_GLASS_BRANCH) {
                        /* reflex_emit*reflex_rate + (reflex_rate == 0 ? 0 : refraction_emit*refraction_rate), AccPathTracer.cpp:151-160 */
                        int refl_zero = sc.reflex_rate.x == 0.f && sc.reflex_rate.y == 0.f && sc.reflex_rate.z == 0.f;
                        if (!refl_zero && sp < 255) {
                            stack[sp].ray = sc.refraction; stack[sp].thr = vmul(thr, sc.refraction_rate);
                            stack[sp].depth = d + 1; stack[sp].branch = branch | (1u << (d & 31)); sp++;
                        }
                        thr = vmul(thr, sc.reflex_rate); ray = sc.reflex;
                    } else {
                        /* one-sample estimator of the same sum: reflect w.p. q, weight rate/q */
                        float q = sc.reflex_rate.x + sc.reflex_rate.y + sc.reflex_rate.z;
                        float q2 = sc.refraction_rate.x + sc.refraction_rate.y + sc.refraction_rate.z;
                        int refl_zero = sc.reflex_rate.x == 0.f && sc.reflex_rate.y == 0.f && sc.reflex_rate.z == 0.f;
                        if (refl_zero || !(q + q2 > 0.f)) { thr = V(0, 0, 0); break; }
                        float pr = q / (q + q2);
                        if (u01(rn[2]) < pr) { thr = vmul(thr, vdivs(sc.reflex_rate, pr)); ray = sc.reflex; }
                        else { thr = vmul(thr, vdivs(sc.refraction_rate, 1.f - pr)); ray = sc.refraction; }
                    }
                } else if (type == 1) {
                    v3 f; ray = shade_conductor(m, ray, h.p, h.n, &f); thr = vmul(thr, f);
                } else if (type == 3) {
                    v3 f; ray_t nr;
                    if (!shade_microfacet(s, m, ray, h.p, h.n, &nr, &f)) break; /* zero attenuation: contributes nothing */
                    ray = nr; thr = vmul(thr, f);
                } else {
                    /* type 0; any other type falls off the end of the reference's trace() (UB) - treated as Lambertian */
                    v3 albedo = ld3(m->diffuse_color);
                    if (nee == 2 && d + 1 < s->depth) {   /* the environment map sampled directly */
                        ray_t shadow; v3 contrib;
                        if (env_sample(s, albedo, h.p, h.n, thr, u01(rn[2]), u01(rn[3]), &shadow, &contrib)) {
                            (*rays)++;
                            hit_t sh = closest_hit(s, shadow, NULL, NULL);
                            v3 srad; int swhich;
                            closest_light_which(s, shadow, &srad, &swhich);
                            if (!sh.hit && swhich < 0) L = vadd(L, contrib);   /* seen iff the ray leaves the scene */
                        }
                        mis = 1;
                    } else if (nee && d + 1 < s->depth) {   /* only where the continuation is really traced (AccPathTracer.cpp:122) */
                        ray_t shadow; v3 contrib; int li;
                        if (nee_sample(s, albedo, h.p, h.n, thr, u01(rn[2]), u01(rn[3]), &shadow, &contrib, &li)) {
                            (*rays)++;
                            hit_t sh = closest_hit(s, shadow, NULL, NULL);
                            v3 srad; int swhich;
                            float stl = closest_light_which(s, shadow, &srad, &swhich);
                            if (swhich == li && !(sh.hit && sh.t < stl)) L = vadd(L, contrib);
                        }
                        mis = 1;
                    }
                    v3 f; ray = shade_lambertian(albedo, h.p, h.n, u01(rn[0]), u01(rn[1]), &f); thr = vmul(thr, f);
                }
            } else if (tl != INFINITY) {
                v3 add = vmul(thr, radiance);
                if (nee == 1 && weigh_light) add = vscale(add, mis_light_weight(s, ray, which, tl));
                L = vadd(L, add); break;
            } else {
                if (s->env_rgba && s->mode == NRCU_MODE_ACC) {
                    v3 add = vmul(thr, env_lookup(s, ray.d));
                    if (nee == 2 && weigh_light) add = vscale(add, mis_env_weight(s, ray.d));
                    L = vadd(L, add);
                }
                break;
            }
        }
    }
    return L;
}

typedef struct { const nro_scene* s; uint64_t seed; uint32_t s0, s1; int glass_mode; int nee; const uint32_t* pixels; float* accum4;
                 uint64_t rays[PF_MAX_THREADS]; } pt_ctx;
static void pt_body(int64_t b, int64_t e, void* c_, int tid) {
    pt_ctx* c = (pt_ctx*)c_;
    const nro_scene* s = c->s;
    uint64_t rays = 0;
    for (int64_t q = b; q < e; q++) {
        uint32_t p = c->pixels ? c->pixels[q] : (uint32_t)q;
        v3 sum = V(0, 0, 0);
        for (uint32_t k = c->s0; k < c->s1; k++) {
            ray_t r = camera_ray(s, c->seed, p, k);
            sum = vadd(sum, pt_trace(s, c->seed, p, k, r, c->glass_mode, c->nee, &rays));
        }
        float* a = c->accum4 + 4 * q;
        a[0] += sum.x; a[1] += sum.y; a[2] += sum.z; a[3] += (float)(c->s1 - c->s0);
    }
    c->rays[tid] += rays;
}
void nro_render_pt_pixels_flags(const nro_scene* s, uint64_t seed, uint32_t s0, uint32_t s1, int glass_mode, uint32_t flags,
                                const uint32_t* pixels, uint32_t n_pixels, float* accum4, uint64_t* rays_out);
void nro_render_pt_pixels(const nro_scene* s, uint64_t seed, uint32_t s0, uint32_t s1, int glass_mode,
                          const uint32_t* pixels, uint32_t n_pixels, float* accum4, uint64_t* rays_out) {
    nro_render_pt_pixels_flags(s, seed, s0, s1, glass_mode, 0u, pixels, n_pixels, accum4, rays_out);
}
/* flags: nrcu_render_flags (NRCU_FLAG_NEE = the next-event-estimation extension; 0 = the reference's estimator) */
void nro_render_pt_pixels_flags(const nro_scene* s, uint64_t seed, uint32_t s0, uint32_t s1, int glass_mode, uint32_t flags,
                                const uint32_t* pixels, uint32_t n_pixels, float* accum4, uint64_t* rays_out) {
    if (s1 == 0 && s0 == 0) s1 = s->spp;
    pt_ctx* c = (pt_ctx*)calloc(1, sizeof(pt_ctx));
    c->s = s; c->seed = seed; c->s0 = s0; c->s1 = s1; c->glass_mode = glass_mode; c->pixels = pixels; c->accum4 = accum4;
    c->nee = (flags & NRCU_FLAG_NEE) != 0;
    if ((flags & NRCU_FLAG_ENV_IS) && s->env_rgba && s->mode == NRCU_MODE_ACC) {
        if (!s->env_sin) env_build_tables((nro_scene*)s);
        if (s->env_total > 0.f) c->nee = 2;
    }
    if (!pixels) n_pixels = s->width * s->height;
    parallel_for(n_pixels, 64, pt_body, c);
    uint64_t total = 0;
    for (int i = 0; i < PF_MAX_THREADS; i++) total += c->rays[i];
    if (rays_out) *rays_out = total;
    free(c);
}

void nro_render_pt(const nro_scene* s, uint64_t seed, uint32_t s0, uint32_t s1, int glass_mode, float* accum, uint64_t* rays) {
    nro_render_pt_pixels(s, seed, s0, s1, glass_mode, NULL, s->width * s->height, accum, rays);
}

void nro_resolve(const float* accum, uint64_t n_pixels, float* rgba) {
    for (uint64_t i = 0; i < n_pixels; i++) {
        float c = accum[4 * i + 3];
        rgba[4 * i + 0] = sqrtf(accum[4 * i + 0] / c);
        rgba[4 * i + 1] = sqrtf(accum[4 * i + 1] / c);
        rgba[4 * i + 2] = sqrtf(accum[4 * i + 2] / c);
        rgba[4 * i + 3] = 1.f;
    }
}
