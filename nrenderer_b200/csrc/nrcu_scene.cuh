// nrcu_scene.cuh — device-side scene layout (HBM-resident, read through 128-bit __ldg loads).
//
// Primitive i (ids in the order nrcu_primitive_count documents) owns
//   prim_geom[3i..3i+2]  what the intersection test reads (48 B, three LDG.128):
//       sphere   : (c.x c.y c.z r) - -
//       triangle : (v1.x v1.y v1.z e1.x) (e1.y e1.z e2.x e2.y) (e2.z - - -)      e1 = v2-v1, e2 = v3-v1
//       plane    : (n.x n.y n.z p.x) (p.y p.z r0.x r0.y) (r0.z r1.x r1.y r1.z)   r0,r1 = first two rows of
//                  inverse(mat3(u, v, cross(u,v))) — the reference inverts this matrix per test
//                  (intersections.cpp:64-66); the rows are a pure function of (u,v), so they are
//                  precomputed with the same operation order and give bit-identical results.
//   prim_shade[i]        (normal.xyz, kind | material << 2 as bits) read once per hit by the shading kernel
//   prim_box[2i..2i+1]   (min.xyz -)(max.xyz -): the reference's Bounds3 of the leaf (Bounds3.hpp:35-103),
//                        used for the AccPathTracer leaf gate and as BVH build input
//   prim_meta[i]         kind | material << 2
#pragma once
#include "nrcu_math.cuh"

namespace nrcu {

struct alignas(16) f4 { float x, y, z, w; };
NR_HD f4 mk4(float x, float y, float z, float w) { f4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct alignas(16) i4 { int x, y, z, w; };

NR_HD f4 ldg4(const f4* p) {
#if defined(__CUDA_ARCH__)
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return mk4(v.x, v.y, v.z, v.w);
#else
    return *p;
#endif
}
NR_HD i4 ldg4i(const f4* p) {
#if defined(__CUDA_ARCH__)
    int4 v = __ldg(reinterpret_cast<const int4*>(p));
    i4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    i4 r; const int* q = reinterpret_cast<const int*>(p); r.x = q[0]; r.y = q[1]; r.z = q[2]; r.w = q[3]; return r;
#endif
}
NR_HD uint32_t ldg_u32(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
NR_HD int f2i(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i; __builtin_memcpy(&i, &f, 4); return i;
#endif
}
NR_HD float i2f(int i) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f; __builtin_memcpy(&f, &i, 4); return f;
#endif
}

enum { KIND_SPHERE = 0, KIND_TRIANGLE = 1, KIND_PLANE = 2, KIND_MESH = 3 };
enum { MODE_RAYCAST = 0, MODE_SIMPLE = 1, MODE_ACC = 2 };

// nrcu_material with the shader-constructor defaults applied (same 24-word layout)
struct DMaterial {
    uint32_t type, present;
    float diffuse_color[3], specular_color[3], specular_ex, albedo[3], eta_r[3], eta_i[3], ior, absorbed[3], roughness, f0;
};

struct DCamera {   // Camera ctor results, ray_cast/include/Camera.hpp:25-46
    vec3 position, lower_left, horizontal, vertical, u, v;
    float lens_radius;
};

#define NRCU_BVH_NODE_F4 7   // float4s per BVH4 node: cx hx cy hy cz hz (centre / half extent of the padded boxes, 4 children each) + 4 child refs
#ifndef NRCU_LEAF_MAX
#define NRCU_LEAF_MAX 4
#endif
// ^     // primitives per leaf the builder aims for
// child ref encoding: >= 0 inner node index; < 0 leaf: ~ref = (first << 4) | (count - 1), count <= 16;
// empty slots carry an inverted box (lo = +inf, hi = -inf) and are never entered.
#define NRCU_REF_EMPTY 0x7fffffff
#define NRCU_LIGHT_F4 6
#define NRCU_MAX_BIG 32       // capacity of the wide-primitive list
#ifndef NRCU_BIG_AREA_FRACTION
#define NRCU_BIG_AREA_FRACTION 0.01f
#endif
//   // a primitive is "wide" when its box has >= this fraction of the scene box's surface area

struct DScene {
    int mode;
    uint32_t width, height, depth;
    uint32_t n_prims;
    const f4* prim_geom;
    const f4* prim_shade;
    const f4* prim_box;
    const uint32_t* prim_meta;
    // wide BVH
    const f4* nodes;
    const uint32_t* leaf_prims;   // (prim id << 2) | kind, grouped by leaf
    const f4* leaf_geom;          // prim_geom gathered into leaf order (3 per leaf slot): no dependent load at the leaves
    const f4* leaf_box;           // prim_box gathered into leaf order (2 per leaf slot), read by the AccPathTracer leaf gate
    int root_ref;                 // child-ref encoding; NRCU_REF_EMPTY for an empty scene
    // "wide" primitives (boxes covering a sizeable part of the scene: walls, floors) are kept out of the BVH and
    // tested by every ray in a warp-uniform loop (k_big) before the traversal starts; see nrcu_bvh.cuh
    const f4* big_geom;           // 3 per wide primitive
    const f4* big_box;            // 2 per wide primitive (leaf gate)
    const f4* big_bound;          // 2 per wide primitive: padded true bounds as (centre, half extent) (conservative pre-test)
    const uint32_t* big_meta;     // (prim id << 2) | kind; planes first, then triangles, then spheres
    uint32_t n_big;
    const f4* big_rect;           // per wide primitive: film rectangle (x0, x1, y0, y1) of its bounds seen from a pinhole camera; null: none
    vec3 bvh_lo, bvh_hi;          // padded bounds of everything inside the BVH (lo > hi when it is empty)
    // LIVE PIXELS (GPU only, pinhole camera, no environment map): the pixels whose camera rays can meet anything at all - a wide
    // primitive's bounds, the BVH's bounds or a light - listed in pixel order; the others render black whatever is sampled, so
    // no camera ray is generated for them.  null: every pixel is live (n_live = width * height).
    const uint32_t* live_px;
    const unsigned char* live_flag;   // per pixel, for the accumulation (dead pixels have no radiance slots written)
    uint32_t n_live;
    // with an environment map (AccPathTracer mode) the dead pixels are not black: their samples look the map up along the
    // camera ray.  That is done at once, one thread per (dead pixel, sample), without a queue entry (k_env_dead).
    const uint32_t* dead_px;          // the dead pixels in pixel order (null unless dead_env)
    uint32_t dead_env;
    // shading
    const DMaterial* materials;
    const f4* mat_head;           // per material: (diffuseColor / pi  - Lambertian.cpp:30, the same fp32 division -, type bits): one gather for the common case
    uint32_t n_materials;
    const f4* area_lights;        // NRCU_LIGHT_F4 float4 per light: quad record (as a plane with n = cross(u,v)), radiance, u, v
    uint32_t n_area_lights;
    uint32_t n_point_lights;
    vec3 point_position, point_intensity;   // pointLightBuffer[0] (RayCastRenderer.cpp:41-42)
    DCamera cam;
    vec3 ambient;
    const f4* env_rgba; int env_w, env_h;   // ambient environment map (extension) or null
    // importance-sampling tables of the map (NRCU_FLAG_ENV_IS): [0,h) sin(theta) at the row centres, [h,2h) marginal CDF over
    // the rows (weight = sin x row luminance), [2h, 2h + w*h) conditional CDF inside each row; env_total = sum of the row weights
    const float* env_tab; float env_total;
    // Microfacet: the half-vector in the local frame is a constant because the reference reseeds
    // minstd_rand with 6 on every call (Microfacet.cpp:65-76); computed on the host with libm.
    float mf_u1, mf_u2, mf_cos_phi, mf_sin_phi;
    int nee;                      // direct sampling at Lambertian vertices (extensions): 0 off = the reference's estimator, 1 area lights (NRCU_FLAG_NEE), 2 environment map (NRCU_FLAG_ENV_IS)
    // Loud failure instead of a silently wrong frame: [0] traversal-stack entries that did not fit (a hit may have been
    // missed), [1] rays that did not fit the branching-glass queue.  Read back by the host at every synchronisation
    // point; non-zero => NRCU_ERR_OVERFLOW.
    uint32_t* overflow;
    int stack_limit;              // entries per traversal stack (NRCU_LOCAL_STACK; lowered only by the overflow test)
};

}  // namespace nrcu
