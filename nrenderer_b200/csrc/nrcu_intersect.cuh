// nrcu_intersect.cuh — ray/primitive tests on the packed device records + wide-BVH traversal.
//
// The primitive tests reproduce, operation for operation, the reference's
//   Intersection::xTriangle / xSphere / xPlane / xAreaLight
//     RayCast variant : code/components/ray_cast/src/intersections/intersections.cpp:5-93
//     path tracers    : code/components/acc_path_tracing/src/intersections/intersections.cpp:5-94
// (template parameter RC selects the RayCast variant: exclusive lower bound t <= tMin), so hit
// decisions and t values are bit-identical to the CPU components.  Only t is returned: the hit point
// (ray.at(t)) and the normal are recomputed from the same ray by the shading kernel.
//
// closest_hit_bvh() is the replacement for closestHitObject (SimplePathTracer.cpp:104-129,
// AccPathTracer.cpp:87-99 -> BVHTree::getIntersect BVH.hpp:93-164): inner nodes are culled with a
// conservative slab test, the exact reference tests run only at the leaves, and equal-t ties go to
// the lowest primitive id.  That is what the in-order loops of RayCast and SimplePathTracer produce.  For
// AccPathTracer it matches the reference EXCEPT in two corner cases the parity tests exclude: exact-t ties (the
// reference's tree returns `hit1->t < hit2->t ? hit1 : hit2`, BVH.hpp:153, i.e. the RIGHT subtree wins a tie, which
// depends on its unstable std::sort) and rays that pass a leaf's box but fail Bounds3::IntersectP of one of its
// ancestors' boxes through rounding (the reference gates every inner node, this code and the oracle gate the leaf only).
#pragma once
#include "nrcu_scene.cuh"

namespace nrcu {

// ---- exact primitive tests -------------------------------------------------------------------
template <bool RC>
NR_HD bool t_in_range(float t, float tmin, float tmax) {
    if (RC) return !(t >= tmax || t <= tmin);
    return !(t >= tmax || t < tmin);
}

template <bool RC>
NR_HD bool x_triangle(const Ray& ray, f4 g0, f4 g1, f4 g2, float tmin, float tmax, float& t_out) {
    vec3 v1 = mk3(g0.x, g0.y, g0.z), e1 = mk3(g0.w, g1.x, g1.y), e2 = mk3(g1.z, g1.w, g2.x);
    vec3 P = cross(ray.d, e2);
    float det = dot(e1, P);
    vec3 T;
    if (det > 0) T = ray.o - v1; else { T = v1 - ray.o; det = -det; }
    if (det < 0.000001f) return false;
    float u = dot(T, P);
    if (u > det || u < 0.f) return false;
    vec3 Q = cross(T, e1);
    float v = dot(ray.d, Q);
    if (v < 0.f || v + u > det) return false;
    float w = dot(e2, Q);
    float inv_det = 1.f / det;
    w *= inv_det;
    if (!t_in_range<RC>(w, tmin, tmax)) return false;
    t_out = w;
    return true;
}

template <bool RC>
NR_HD bool x_sphere(const Ray& ray, f4 g0, float tmin, float tmax, float& t_out) {
    vec3 position = mk3(g0.x, g0.y, g0.z);
    float r = g0.w;
    vec3 oc = ray.o - position;
    float a = dot(ray.d, ray.d);
    float b = dot(oc, ray.d);
    float c = dot(oc, oc) - r * r;
    float disc = b * b - a * c;
    if (!(disc > 0)) return false;
    float sq = sqrtf(disc);
    float temp = (-b - sq) / a;
    if (temp < tmax && (RC ? temp > tmin : temp >= tmin)) { t_out = temp; return true; }
    temp = (-b + sq) / a;
    if (temp < tmax && (RC ? temp > tmin : temp >= tmin)) { t_out = temp; return true; }
    return false;
}

// xPlane / xAreaLight: n is the normal the variant uses (normalised for RayCast, as stored for the
// path tracers, cross(u,v) for lights), folded into the record at upload time.
// `cull_t` (NRCU_INF = off): the caller cannot use a hit with t > cull_t.  Two rejections are taken before
// the division; both are exact, i.e. they only drop candidates the reference drops too (or that cannot win):
//   * num and nd of opposite sign  =>  t = num/nd <= -0 < tMin (tMin > 0 in every variant)
//   * |num| > cull_t*|nd|*(1+1e-5)  =>  the rounded quotient is strictly greater than cull_t
template <bool RC>
NR_HD bool x_quad(const Ray& ray, f4 g0, f4 g1, f4 g2, float tmin, float tmax, float cull_t, float& t_out) {
    vec3 n = mk3(g0.x, g0.y, g0.z), p = mk3(g0.w, g1.x, g1.y);
    float nd = dot(ray.d, n);
    if (nd < 0.0000001f && nd > -0.00000001f) return false;
    float dp = -dot(p, n);
    float num = -dp - dot(n, ray.o);
    // |nd| >= 1e-8 here, so |num| < 1e-30 gives |t| < 1e-22 < tMin: rejected exactly like the reference,
    // without the division (a zero / denormal numerator - a ray leaving the very plane it is tested
    // against - would otherwise take the slow path of the IEEE division on the device).
    if (fabsf(num) < 1e-30f) return false;
    if ((num < 0.f) != (nd < 0.f)) return false;
    if (fabsf(num) > cull_t * fabsf(nd) * 1.00001f) return false;
    float t = num / nd;
    if (!t_in_range<RC>(t, tmin, tmax)) return false;
    vec3 q = ray_at(ray, t) - p;
    // glm mat3*vec3 (type_mat3x3.inl:468-474): m[0][r]*v.x + m[1][r]*v.y + m[2][r]*v.z
    float ru = g1.z * q.x + g1.w * q.y + g2.x * q.z;
    float rv = g2.y * q.x + g2.z * q.y + g2.w * q.z;
    if ((ru <= 1 && ru >= 0) && (rv <= 1 && rv >= 0)) { t_out = t; return true; }
    return false;
}

// First two rows of glm::inverse(mat3(u, v, cross(u,v))) (glm/detail/func_matrix.inl compute_inverse<3,3>).
NR_HD void quad_inverse_rows(vec3 u, vec3 v, float r0[3], float r1[3]) {
    vec3 w = cross(u, v);
    float m00 = u.x, m01 = u.y, m02 = u.z, m10 = v.x, m11 = v.y, m12 = v.z, m20 = w.x, m21 = w.y, m22 = w.z;
    float ood = 1.0f / (+m00 * (m11 * m22 - m21 * m12) - m10 * (m01 * m22 - m21 * m02) + m20 * (m01 * m12 - m11 * m02));
    r0[0] = +(m11 * m22 - m21 * m12) * ood;
    r0[1] = -(m10 * m22 - m20 * m12) * ood;
    r0[2] = +(m10 * m21 - m20 * m11) * ood;
    r1[0] = -(m01 * m22 - m21 * m02) * ood;
    r1[1] = +(m00 * m22 - m20 * m02) * ood;
    r1[2] = -(m00 * m21 - m20 * m01) * ood;
}

template <bool RC>
NR_HD bool x_prim(const DScene& s, const Ray& ray, uint32_t id, uint32_t kind, float tmin, float tmax, float& t_out) {
    const f4* g = s.prim_geom + 3 * (size_t)id;
    f4 g0 = ldg4(g);
    if (kind == KIND_SPHERE) return x_sphere<RC>(ray, g0, tmin, tmax, t_out);
    f4 g1 = ldg4(g + 1), g2 = ldg4(g + 2);
    if (kind == KIND_PLANE) return x_quad<RC>(ray, g0, g1, g2, tmin, tmax, tmax, t_out);
    return x_triangle<RC>(ray, g0, g1, g2, tmin, tmax, t_out);
}

// Bounds3::IntersectP, acc_path_tracing/include/Bounds3.hpp:141-168, with invDir as built in
// BVHTree::getIntersect (BVH.hpp:97: double 1./d narrowed to float).  Rounding the double quotient
// to float gives the correctly rounded float quotient (double rounding is innocuous for division
// when the wide format has >= 2p+2 = 50 bits), so (ix, iy, iz) = IEEE 1.0f/d is bit-identical.
NR_HD bool bounds_intersectp_inv(f4 lo, f4 hi, const Ray& ray, float ix, float iy, float iz) {
    vec3 o = ray.o, d = ray.d;
    if (o.x >= lo.x && o.x <= hi.x && o.y >= lo.y && o.y <= hi.y && o.z >= lo.z && o.z <= hi.z) return true;
    float t1n = (lo.x - o.x) * ix, t2n = (lo.y - o.y) * iy, t3n = (lo.z - o.z) * iz;
    float t1x = (hi.x - o.x) * ix, t2x = (hi.y - o.y) * iy, t3x = (hi.z - o.z) * iz;
    if (d.x < 0) { float t = t1n; t1n = t1x; t1x = t; }
    if (d.y < 0) { float t = t2n; t2n = t2x; t2x = t; }
    if (d.z < 0) { float t = t3n; t3n = t3x; t3x = t; }
    // std::max(a,b) = (a < b) ? b : a;  std::min(a,b) = (b < a) ? b : a   (NaN behaviour preserved)
    float in1 = (t2n < t3n) ? t3n : t2n, t_near = (t1n < in1) ? in1 : t1n;
    float in2 = (t3x < t2x) ? t3x : t2x, t_far = (in2 < t1x) ? in2 : t1x;
    return t_far >= 0 && t_near < t_far;
}
NR_HD bool bounds_intersectp(f4 lo, f4 hi, const Ray& ray) {
    return bounds_intersectp_inv(lo, hi, ray, 1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
}

// ---- traversal -------------------------------------------------------------------------------
// Per-thread stack held in local memory: used by the host emulation and as the overflow area of
// the shared-memory stack in the device kernels.
#define NRCU_LOCAL_STACK 96
struct LocalStack {
    float t[NRCU_LOCAL_STACK]; int ref[NRCU_LOCAL_STACK]; int sp; int dropped;   // dropped: entries that did not fit (the answer may be wrong)
    NR_HD LocalStack() : sp(0), dropped(0) {}
    NR_HD void push(float tt, int r) { if (sp < NRCU_LOCAL_STACK) { t[sp] = tt; ref[sp] = r; sp++; } else dropped++; }
    NR_HD bool pop(float& tt, int& r) { if (sp == 0) return false; sp--; tt = t[sp]; r = ref[sp]; return true; }
};

struct RayPrep { vec3 inv, oinv; };
NR_HD RayPrep prep_ray(const Ray& ray) {
    // Reciprocal with the zero/denormal components replaced by a huge finite value: the inner-node
    // test only has to be conservative, the exact tests never see `inv`.
    RayPrep p;
    float dx = ray.d.x, dy = ray.d.y, dz = ray.d.z;
    p.inv.x = fabsf(dx) > 1e-18f ? 1.0f / dx : copysignf(1e18f, dx);
    p.inv.y = fabsf(dy) > 1e-18f ? 1.0f / dy : copysignf(1e18f, dy);
    p.inv.z = fabsf(dz) > 1e-18f ? 1.0f / dz : copysignf(1e18f, dz);
    p.oinv = ray.o * p.inv;
    return p;
}

#define NRCU_CSWAP(ta, ra, tb, rb) do { if (tb < ta) { float _t = ta; ta = tb; tb = _t; int _r = ra; ra = rb; rb = _r; } } while (0)

#define NRCU_REF_DONE ((int)0x80000000)   // "no more work": negative like a leaf so the node loop exits on it

// Pop the next node that can still contain a hit at t <= best_t (ties must stay reachable).
template <class Stack>
NR_HD int pop_next(Stack& stack, float best_t) {
    float pt; int c;
    for (;;) {
        if (!stack.pop(pt, c)) return NRCU_REF_DONE;
        if (pt <= best_t) return c;
    }
}

// One BVH4 node: conservative slab test of the four children (explicit fused multiply-adds; the
// boxes were padded at build time), nearest hit child returned, the other hit children pushed far-to-near.
template <class Stack>
NR_HD int node_step(const DScene& s, const RayPrep& rp, int cur, float best_t, Stack& stack) {
    const f4* n = s.nodes + (size_t)cur * NRCU_BVH_NODE_F4;
    f4 lox = ldg4(n), hix = ldg4(n + 1), loy = ldg4(n + 2), hiy = ldg4(n + 3), loz = ldg4(n + 4), hiz = ldg4(n + 5);
    i4 refs = ldg4i(n + 6);
    // lo* = centres, hi* = half extents: per axis m = c/d - o/d, near = m - h/|d|, far = m + h/|d| (three FFMAs, no min/max)
    const float aix = fabsf(rp.inv.x), aiy = fabsf(rp.inv.y), aiz = fabsf(rp.inv.z);
    float t0, t1, t2, t3;
#define NRCU_SLAB(k, out) do { \
    float mx = fmaf(lox.k, rp.inv.x, -rp.oinv.x), my = fmaf(loy.k, rp.inv.y, -rp.oinv.y), mz = fmaf(loz.k, rp.inv.z, -rp.oinv.z); \
    float tn = fmaxf(fmaxf(fmaf(-hix.k, aix, mx), fmaf(-hiy.k, aiy, my)), fmaxf(fmaf(-hiz.k, aiz, mz), 0.0f)); \
    float tf = fminf(fminf(fmaf(hix.k, aix, mx), fmaf(hiy.k, aiy, my)), fminf(fmaf(hiz.k, aiz, mz), best_t)); \
    out = (tn <= tf) ? tn : NRCU_INF; } while (0)
    NRCU_SLAB(x, t0); NRCU_SLAB(y, t1); NRCU_SLAB(z, t2); NRCU_SLAB(w, t3);
#undef NRCU_SLAB
    int r0 = refs.x, r1 = refs.y, r2 = refs.z, r3 = refs.w;
    // 5-comparator sorting network: nearest child first
    NRCU_CSWAP(t0, r0, t1, r1); NRCU_CSWAP(t2, r2, t3, r3);
    NRCU_CSWAP(t0, r0, t2, r2); NRCU_CSWAP(t1, r1, t3, r3);
    NRCU_CSWAP(t1, r1, t2, r2);
    if (t3 < NRCU_INF) stack.push(t3, r3);
    if (t2 < NRCU_INF) stack.push(t2, r2);
    if (t1 < NRCU_INF) stack.push(t1, r1);
    if (t0 < NRCU_INF) return r0;
    return pop_next(stack, best_t);
}

// One primitive record (g0..g2 = its intersection record, box2 = its reference leaf box, pk = id << 2 | kind):
// the exact reference test.  GATE = AccPathTracer leaf gate (a primitive counts only if its reference leaf box
// passes Bounds3::IntersectP, which silently drops zero-thickness boxes); it is evaluated only for candidates
// that would become the closest hit, with the exact reciprocals `ginv` of the ray direction.  Equal-t ties go
// to the lowest primitive id whatever the visiting order.
template <bool GATE>
NR_HD void prim_test(const Ray& ray, vec3 ginv, f4 g0, f4 g1, f4 g2, const f4* box2, uint32_t pk, float& best_t, int& best_id) {
    const float tmin = (float)0.000001;
    const uint32_t id = pk >> 2, kind = pk & 3u;
    float t;
    bool hit;
    if (kind == KIND_PLANE) hit = x_quad<false>(ray, g0, g1, g2, tmin, NRCU_INF, best_t, t);
    else if (kind == KIND_SPHERE) hit = x_sphere<false>(ray, g0, tmin, NRCU_INF, t);
    else hit = x_triangle<false>(ray, g0, g1, g2, tmin, NRCU_INF, t);
    if (hit && (t < best_t || (t == best_t && (int)id < best_id))) {
        if (GATE) {
            f4 lo = box2[0], hi = box2[1];
            if (!bounds_intersectp_inv(lo, hi, ray, ginv.x, ginv.y, ginv.z)) return;
        }
        best_t = t; best_id = (int)id;
    }
}
// Exact IEEE reciprocals of the direction for the leaf gate (BVH.hpp:97).
NR_HD vec3 gate_inverse(const Ray& ray) { return mk3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z); }
// The same from an already prepared ray: prep_ray's reciprocal IS the exact one unless it was clamped (|d| <= 1e-18),
// which almost never happens - the three divisions are only redone in that case.
NR_HD vec3 gate_inverse(const Ray& ray, const RayPrep& rp) {
    if (fabsf(ray.d.x) > 1e-18f && fabsf(ray.d.y) > 1e-18f && fabsf(ray.d.z) > 1e-18f) return rp.inv;
    return gate_inverse(ray);
}

// The same for leaf slot `slot` of the leaf-ordered arrays.
template <bool GATE>
NR_HD void prim_step(const DScene& s, const Ray& ray, vec3 ginv, uint32_t slot, uint32_t pk, float& best_t, int& best_id) {
    const f4* g = s.leaf_geom + 3 * (size_t)slot;
    f4 g0 = ldg4(g), g1 = ldg4(g + 1), g2 = ldg4(g + 2);
    prim_test<GATE>(ray, ginv, g0, g1, g2, s.leaf_box + 2 * (size_t)slot, pk, best_t, best_id);
}

template <bool GATE>
NR_HD void leaf_step(const DScene& s, const Ray& ray, vec3 ginv, int leaf_ref, float& best_t, int& best_id) {
    uint32_t code = (uint32_t)(~leaf_ref);
    uint32_t first = code >> 4, count = (code & 15u) + 1u;
    for (uint32_t j = 0; j < count; j++) prim_step<GATE>(s, ray, ginv, first + j, ldg_u32(s.leaf_prims + first + j), best_t, best_id);
}

NR_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
// Conservative slab test against a box stored as (centre c, half extent h): per axis m = c/d - o/d, near = m - h/|d|,
// far = m + h/|d| - three FFMAs and no per-axis min/max (13 instead of 16 instructions per box).  `nox` is -o.x/d.x, or
// -inf to make the test fail for a lane that holds no ray.  tn <= tf means "may be hit at t >= 0".
NR_HD void slab_center_extent(f4 c, f4 h, const RayPrep& rp, vec3 ainv, float nox, float& tn, float& tf) {
    const float mx = fmaf(c.x, rp.inv.x, nox), my = fmaf(c.y, rp.inv.y, -rp.oinv.y), mz = fmaf(c.z, rp.inv.z, -rp.oinv.z);
    tn = fmaxf(fmaxf(fmaf(-h.x, ainv.x, mx), fmaf(-h.y, ainv.y, my)), fmaxf(fmaf(-h.z, ainv.z, mz), 0.0f));
    tf = fminf(fminf(fmaf(h.x, ainv.x, mx), fmaf(h.y, ainv.y, my)), fmaf(h.z, ainv.z, mz));
}
// The wide primitives (DScene::big_*), tested by every ray before the traversal (best_t = inf, best_id = -1 on entry).
//   pass 1  every lane walks the same list: conservative slab test of each primitive's padded bounds -> candidate
//           bit mask (warp-uniform loop, shared-memory broadcasts, no divergence);
//   pass 2  each lane runs the exact reference test on its own candidates only (typically 2-3 of them).
// The leaf gate is applied optimistically: the ungated closest hit is found first and only the winner's box is
// tested; if it passes, it is also the closest of the gate-passing primitives (same t order, same lowest-id tie
// rule).  Only when the winner fails its gate (zero-thickness box or a grazing hit) are the candidates walked
// again with the gate applied per candidate.
NR_HD uint32_t big_list_mask(const DScene& s, const f4* bound, const RayPrep& rp) {
    uint32_t mask = 0;
    const vec3 ainv = mk3(fabsf(rp.inv.x), fabsf(rp.inv.y), fabsf(rp.inv.z));
    for (uint32_t k = 0; k < s.n_big; k++) {
        float tn, tf;
        slab_center_extent(bound[2 * k], bound[2 * k + 1], rp, ainv, -rp.oinv.x, tn, tf);
        if (tn <= tf) mask |= 1u << k;
    }
    return mask;
}
// pass 2 on a candidate mask (any superset of the primitives the ray hits gives the same answer)
template <bool GATE>
NR_HD void big_list_resolve(uint32_t mask, const f4* geom, const f4* box, const uint32_t* meta,
                            const Ray& ray, vec3 ginv, float& best_t, int& best_id) {
    uint32_t kb = 0;
    for (uint32_t m = mask; m; m &= m - 1u) {
        uint32_t k = (uint32_t)ctz32(m);
        int before = best_id;
        prim_test<false>(ray, ginv, geom[3 * k], geom[3 * k + 1], geom[3 * k + 2], box, meta[k], best_t, best_id);
        if (best_id != before) kb = k;
    }
    if (GATE && best_id >= 0 && !bounds_intersectp_inv(box[2 * kb], box[2 * kb + 1], ray, ginv.x, ginv.y, ginv.z)) {
        best_t = NRCU_INF; best_id = -1;
        for (uint32_t m = mask; m; m &= m - 1u) {
            uint32_t k = (uint32_t)ctz32(m);
            prim_test<true>(ray, ginv, geom[3 * k], geom[3 * k + 1], geom[3 * k + 2], box + 2 * k, meta[k], best_t, best_id);
        }
    }
}
template <bool GATE>
NR_HD void big_list_step(const DScene& s, const f4* geom, const f4* box, const f4* bound, const uint32_t* meta,
                         const Ray& ray, const RayPrep& rp, vec3 ginv, float& best_t, int& best_id) {
    big_list_resolve<GATE>(big_list_mask(s, bound, rp), geom, box, meta, ray, ginv, best_t, best_id);
}
// Camera rays of a pinhole camera (extension of pass 1, GPU only): every ray leaves the same point, so "can this pixel's
// ray meet the primitive's padded bounds?" is a rectangle test in film coordinates - DScene::big_rect holds, per wide
// primitive, the film-space bounding rectangle of its bounds' eight corners seen from the camera (k_big_rects, fp64, with
// a margin; the whole film when a corner is not in front of the camera).  Four compares instead of a slab test.
NR_HD uint32_t big_list_mask_film(const DScene& s, const f4* rect, float x, float y) {
    uint32_t mask = 0;
    for (uint32_t k = 0; k < s.n_big; k++) {
        const f4 rc = rect[k];
        if (x >= rc.x && x <= rc.y && y >= rc.z && y <= rc.w) mask |= 1u << k;
    }
    return mask;
}
// Can anything inside the BVH still beat best_t?  Conservative slab test against the padded BVH bounds.
NR_HD bool bvh_reachable(const DScene& s, const RayPrep& rp, float best_t) {
    if (s.root_ref == NRCU_REF_EMPTY) return false;
    float ax = fmaf(s.bvh_lo.x, rp.inv.x, -rp.oinv.x), bx = fmaf(s.bvh_hi.x, rp.inv.x, -rp.oinv.x);
    float ay = fmaf(s.bvh_lo.y, rp.inv.y, -rp.oinv.y), by = fmaf(s.bvh_hi.y, rp.inv.y, -rp.oinv.y);
    float az = fmaf(s.bvh_lo.z, rp.inv.z, -rp.oinv.z), bz = fmaf(s.bvh_hi.z, rp.inv.z, -rp.oinv.z);
    float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), best_t));
    return tn <= tf;
}

// Closest hit: the wide primitives first, then the BVH4 ("while-while": run inner nodes until a leaf is
// reached, then the leaf).  This is the single-ray statement of what k_big + k_trace* do on the device.
template <bool GATE, class Stack>
NR_HD void closest_hit_bvh(const DScene& s, const Ray& ray, Stack& stack, float& best_t, int& best_id) {
    best_t = NRCU_INF; best_id = -1;
    RayPrep rp = prep_ray(ray);
    vec3 ginv = gate_inverse(ray);
    big_list_step<GATE>(s, s.big_geom, s.big_box, s.big_bound, s.big_meta, ray, rp, ginv, best_t, best_id);
    if (!bvh_reachable(s, rp, best_t)) return;
    int cur = s.root_ref;
    for (;;) {
        while (cur >= 0) cur = node_step(s, rp, cur, best_t, stack);
        if (cur == NRCU_REF_DONE) return;
        leaf_step<GATE>(s, ray, ginv, cur, best_t, best_id);
        cur = pop_next(stack, best_t);
    }
}

// Brute force in primitive order with shrinking tMax (RayCastRenderer.cpp:66-91).
template <bool RC>
NR_HD void closest_hit_linear(const DScene& s, const Ray& ray, float& best_t, int& best_id) {
    best_t = NRCU_INF; best_id = -1;
    const float tmin = RC ? (float)0.01 : (float)0.000001;
    for (uint32_t i = 0; i < s.n_prims; i++) {
        uint32_t kind = ldg_u32(s.prim_meta + i) & 3u;
        float t;
        if (x_prim<RC>(s, ray, i, kind, tmin, best_t, t) && t < best_t) { best_t = t; best_id = (int)i; }
    }
}

}  // namespace nrcu
