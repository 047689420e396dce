/*
 * nr_oracle.h — CPU restatement of the NRenderer hot path (RayCast, SimplePathTracer,
 * AccPathTracer).  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  It is the checker, never the
 * product: nrenderer_b200/ and libnrcuda.so must not link, import or call it.
 *
 * Parity pin: the restatement is validated against the real reference compiled into
 * oracle/_ref/ (oracle/build_ref.py) — bit-exactly for the deterministic RayCast frame and
 * statistically (linear-space mean / RMSE inside a Monte-Carlo bound) for the path tracers — and
 * against fixtures of those runs committed under tests/golden/.  The reference itself holds no
 * golden vectors or tests for this path.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/code/components unless noted).  Arithmetic is IEEE fp32 in the
 * reference's operation order; build with -ffp-contract=off (see oracle/Makefile).
 */
#ifndef NR_ORACLE_H
#define NR_ORACLE_H

#include <stdint.h>
#include "nrcu.h" /* only for the nrcu_scene / nrcu_material POD definitions and enums */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nro_scene nro_scene;

/* Scene preparation: VertexTransformer::exec + mesh flattening + Bounds3 per primitive + Camera ctor. */
nro_scene* nro_prepare(const nrcu_scene* scene, int mode);
void nro_free(nro_scene* s);
uint32_t nro_primitive_count(const nro_scene* s);
/* Same layout as nrcu_download_primitives. */
void nro_get_primitives(const nro_scene* s, uint32_t* kind, float* data16, int32_t* material);
/* Reference leaf boxes (Bounds3 ctors, acc_path_tracing/include/Bounds3.hpp:35-103): n x 6 (min, max). */
void nro_get_bounds(const nro_scene* s, float* box6);
/* Camera basis: position, lowerLeft, horizontal, vertical, u, v (18 floats) + lens radius. */
void nro_get_camera(const nro_scene* s, float* cam18, float* lens_radius);

/* Brute-force closest hit in primitive order.  tie[i] (may be NULL) = 1 when another primitive
 * hits at exactly the same t (the reference BVH's winner is then order dependent). */
void nro_trace_batch(const nro_scene* s, const float* rays, uint32_t n, int32_t* prim_id, float* t, uint8_t* tie);

/* Bounds3::IntersectP (acc_path_tracing/include/Bounds3.hpp:141-168). */
int nro_bounds_intersectp(const float box6[6], const float origin[3], const float dir[3]);

/* RayCastRenderer::render (ray_cast/src/RayCastRenderer.cpp:14-38): rgba = w*h*4, row 0 = top. */
void nro_render_raycast(const nro_scene* s, float* rgba);

/* Path tracers with the counter-based RNG of DESIGN.md.  Adds linear radiance sums of samples
 * [s0, s1) into accum (w*h*4: rgb sums, a = sample count; row 0 = top).  rays (may be NULL)
 * receives the number of closest-hit queries.  glass_mode: nrcu_glass_mode. */
void nro_render_pt(const nro_scene* s, uint64_t seed, uint32_t s0, uint32_t s1, int glass_mode,
                   float* accum, uint64_t* rays);
/* Same for a list of pixels only (pixel index = row_from_top*w + col), for sparse checks at full size. */
void nro_render_pt_pixels(const nro_scene* s, uint64_t seed, uint32_t s0, uint32_t s1, int glass_mode,
                          const uint32_t* pixels, uint32_t n_pixels, float* accum4, uint64_t* rays);
/* The same with nrcu_render_flags (NRCU_FLAG_NEE: next-event estimation, an extension of this backend restated here
 * independently of the CUDA sources); pixels == NULL renders the whole frame. */
void nro_render_pt_pixels_flags(const nro_scene* s, uint64_t seed, uint32_t s0, uint32_t s1, int glass_mode, uint32_t flags,
                                const uint32_t* pixels, uint32_t n_pixels, float* accum4, uint64_t* rays_out);
/* rgba = (sqrt(sum/count), 1)  (AccPathTracer.cpp:14-16, 32-34) */
void nro_resolve(const float* accum, uint64_t n_pixels, float* rgba);

/* One camera ray as the path tracers generate it for (pixel, sample): out6 = origin, direction. */
void nro_camera_ray(const nro_scene* s, uint64_t seed, uint32_t pixel, uint32_t sample, float* out6);

void nro_philox4x32(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);
void nro_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
