// nrcu_kernels.cuh — the __global__ kernels (sm_100a).  Thin wrappers around the
// __host__ __device__ bodies in nrcu_{intersect,shade,bvh}.cuh plus what only exists on the GPU:
// persistent-thread work fetching, shared-memory traversal stacks, warp-aggregated queue compaction.
#pragma once
#include <cuda_runtime.h>
#include "nrcu_bvh.cuh"
#include "nrcu_prep.cuh"
#include "nrcu_shade.cuh"

namespace nrcu {

// ---------------------------------------------------------------------------------------------
// scene preparation (bodies in nrcu_prep.cuh)
// ---------------------------------------------------------------------------------------------
__global__ void k_mesh_transform(float* pos, uint32_t first_vertex, uint32_t n_vertices) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_vertices) mesh_transform_vertex(pos, first_vertex + i);
}
__global__ void k_build_prims(PrimSources ps, uint32_t n, int raycast, f4* geom, f4* shade, f4* box, uint32_t* meta, float* export16) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) build_prim(ps, i, raycast, geom, shade, box, meta, export16);
}

// ---------------------------------------------------------------------------------------------
// BVH build: one kernel per step body
// ---------------------------------------------------------------------------------------------
#define NRCU_STEP_KERNEL(name, body) \
    __global__ void name(BvhBuild b, int first, int count) { \
        int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= count) return; body(b, first + i); }
NRCU_STEP_KERNEL(k_bvh_clear, node_clear)
NRCU_STEP_KERNEL(k_bvh_init_prim, bvh_init_prim)
NRCU_STEP_KERNEL(k_bvh_level_prepare, bvh_level_prepare)
NRCU_STEP_KERNEL(k_bvh_bin, bvh_bin)
NRCU_STEP_KERNEL(k_bvh_split, bvh_split)
NRCU_STEP_KERNEL(k_bvh_partition, bvh_partition)
NRCU_STEP_KERNEL(k_bvh_leaf_alloc, bvh_leaf_alloc)
NRCU_STEP_KERNEL(k_bvh_leaf_fill, bvh_leaf_fill)
NRCU_STEP_KERNEL(k_bvh_leaf_sort, bvh_leaf_sort)
NRCU_STEP_KERNEL(k_bvh_wide_index, bvh_wide_index)
NRCU_STEP_KERNEL(k_bvh_wide_emit, bvh_wide_emit)

// ---------------------------------------------------------------------------------------------
// RayCast (deterministic; one thread per pixel, brute force over the handful of primitives)
// ---------------------------------------------------------------------------------------------
__global__ void k_raycast(DScene s, f4* rgba, unsigned long long* ray_counter) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = s.width * s.height;
    uint32_t rays = 0;
    if (p < n) {
        vec3 c = raycast_pixel(s, p, &rays);
        rgba[p] = mk4(c.x, c.y, c.z, 1.f);
    }
    // one 64-bit atomic per warp
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_down_sync(0xffffffffu, rays, o);
    if ((threadIdx.x & 31) == 0 && rays) atomicAdd(ray_counter, (unsigned long long)rays);
}

// ---------------------------------------------------------------------------------------------
// Wavefront path tracer
// ---------------------------------------------------------------------------------------------
// A queue entry is three float4:  a = (o.xyz, d.x)  b = (d.y, d.z, thr.x, thr.y)  c = (thr.z, slot, branch, -)
// slot = sample_in_wave * n_pixels + pixel identifies the path; its radiance lands in L[slot].
struct PathQueue { f4* a; f4* b; f4* c; };

__global__ void k_raygen(DScene s, uint64_t seed, uint32_t sample0, uint32_t n_slots, PathQueue q, f4* L, uint32_t* n_queue) {
    uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    if (slot == 0) *n_queue = s.depth == 0 ? 0u : n_slots;   // every slot starts one path: the bounce-0 queue is dense
    uint32_t npix = s.width * s.height;
    uint32_t pixel = slot % npix, sample = sample0 + slot / npix;
    if (s.depth == 0) { L[slot] = mk4(s.ambient.x, s.ambient.y, s.ambient.z, 0.f); return; }   // trace(): currDepth == depth
    L[slot] = mk4(0.f, 0.f, 0.f, 0.f);
    Ray r = pt_camera_ray(s, seed, pixel, sample);
    q.a[slot] = mk4(r.o.x, r.o.y, r.o.z, r.d.x);
    q.b[slot] = mk4(r.d.y, r.d.z, 1.f, 1.f);
    q.c[slot] = mk4(1.f, i2f((int)slot), i2f(0), 0.f);
}

#define NRCU_TRACE_THREADS 128
#define NRCU_SMEM_STACK 20
// Traversal stack: the first NRCU_SMEM_STACK entries of every thread live in shared memory
// ([entry][thread] so that a warp's accesses are conflict free), deeper ones in local memory.
struct SmemStack {
    uint2* base; int sp;
    float ot[NRCU_LOCAL_STACK - NRCU_SMEM_STACK]; int oref[NRCU_LOCAL_STACK - NRCU_SMEM_STACK];
    __device__ __forceinline__ void push(float t, int r) {
        if (sp < NRCU_SMEM_STACK) base[sp * NRCU_TRACE_THREADS] = make_uint2(__float_as_uint(t), (unsigned)r);
        else if (sp < NRCU_LOCAL_STACK) { ot[sp - NRCU_SMEM_STACK] = t; oref[sp - NRCU_SMEM_STACK] = r; }
        else return;
        sp++;
    }
    __device__ __forceinline__ bool pop(float& t, int& r) {
        if (sp == 0) return false;
        sp--;
        if (sp < NRCU_SMEM_STACK) { uint2 v = base[sp * NRCU_TRACE_THREADS]; t = __uint_as_float(v.x); r = (int)v.y; }
        else { t = ot[sp - NRCU_SMEM_STACK]; r = oref[sp - NRCU_SMEM_STACK]; }
        return true;
    }
};

// Persistent-thread closest-hit kernel: the grid is sized to fill the machine once; each warp
// fetches 32 rays at a time from the queue with one atomic (lane 0) and a shuffle.
template <bool GATE>
__global__ void __launch_bounds__(NRCU_TRACE_THREADS) k_trace(DScene s, PathQueue q, const uint32_t* n_ptr, float2* hits,
                                                              uint32_t* fetch, unsigned long long* ray_counter) {
    __shared__ uint2 stack_mem[NRCU_SMEM_STACK * NRCU_TRACE_THREADS];
    const uint32_t n = *n_ptr;
    const uint32_t lane = threadIdx.x & 31u;
    SmemStack stack; stack.base = stack_mem + threadIdx.x; stack.sp = 0;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(fetch, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        if (lane == 0) atomicAdd(ray_counter, (unsigned long long)min(32u, n - base));
        uint32_t i = base + lane;
        if (i < n) {
            f4 a = q.a[i], b = q.b[i];
            Ray r; r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
            float t; int id;
            stack.sp = 0;
            closest_hit_bvh<GATE>(s, r, stack, t, id);
            hits[i] = make_float2(t, __int_as_float(id));
        }
    }
}

// v2 of the closest-hit kernel: persistent threads, "while-while" traversal and lane refill.
// Every lane keeps one ray's traversal state in registers.  When a lane finishes its ray it goes
// idle; as soon as at least `refill` lanes of the warp are idle (or all of them), the warp claims that
// many new rays from the queue with ONE atomic (the first idle lane) and hands them out by
// ballot/popc rank, so warps stay full although path-traced rays have very different lengths.  Leaf
// work is postponed until every active lane has reached a leaf (Aila & Laine's while-while), which
// keeps node steps and primitive tests from serialising against each other.
template <bool GATE>
__global__ void __launch_bounds__(NRCU_TRACE_THREADS) k_trace2(DScene s, PathQueue q, const uint32_t* n_ptr, float2* hits,
                                                               uint32_t* fetch, unsigned long long* ray_counter, uint32_t refill) {
    __shared__ uint2 stack_mem[NRCU_SMEM_STACK * NRCU_TRACE_THREADS];
    const uint32_t n = *n_ptr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;
    SmemStack stack; stack.base = stack_mem + threadIdx.x; stack.sp = 0;
    Ray r; RayPrep rp;
    r.o = mk3(0.f); r.d = mk3(0.f); rp.inv = mk3(0.f); rp.oinv = mk3(0.f);
    float best_t = NRCU_INF; int best_id = -1; int cur = NRCU_REF_DONE; uint32_t idx = 0;
    bool active = false, exhausted = false;
    for (;;) {
        uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle == 0xffffffffu && exhausted) break;
        uint32_t n_idle = __popc(idle);
        if (!exhausted && n_idle >= (idle == 0xffffffffu ? 1u : refill)) {
            uint32_t base = 0;
            const uint32_t leader = __ffs(idle) - 1u;
            if (lane == leader) {
                base = atomicAdd(fetch, n_idle);
                if (base < n) atomicAdd(ray_counter, (unsigned long long)min(n_idle, n - base));
            }
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + n_idle >= n) exhausted = true;
            if (!active) {
                uint32_t i = base + __popc(idle & lt);
                if (i < n) {
                    f4 a = q.a[i], b = q.b[i];
                    r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
                    rp = prep_ray(r);
                    best_t = NRCU_INF; best_id = -1; stack.sp = 0; idx = i;
                    cur = s.root_ref == NRCU_REF_EMPTY ? NRCU_REF_DONE : s.root_ref;
                    active = true;
                }
            }
        }
        if (active) {
            while (cur >= 0) cur = node_step(s, rp, cur, best_t, stack);
            if (cur != NRCU_REF_DONE) {
                leaf_step<GATE>(s, r, cur, best_t, best_id);
                cur = pop_next(stack, best_t);
            }
            if (cur == NRCU_REF_DONE) {
                hits[idx] = make_float2(best_t, __int_as_float(best_id));
                active = false;
            }
        }
    }
}

// Brute-force variant for the RayCast-mode parity probe (nrcu_trace_batch in NRCU_MODE_RAYCAST).
__global__ void k_trace_linear_rc(DScene s, PathQueue q, uint32_t n, float2* hits) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f4 a = q.a[i], b = q.b[i];
    Ray r; r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
    float t; int id;
    closest_hit_linear<true>(s, r, t, id);
    hits[i] = make_float2(t, __int_as_float(id));
}

// Shading + next-ray generation for bounce `d`; surviving paths are compacted into `qo` with one
// atomic per warp (ballot + popc prefix).
__global__ void __launch_bounds__(256) k_shade(DScene s, uint64_t seed, uint32_t d, int glass_branch, uint32_t sample0,
                                              PathQueue qi, const uint32_t* n_in_ptr, const float2* hits,
                                              PathQueue qo, uint32_t* n_out_ptr, uint32_t out_capacity, f4* L) {
    const uint32_t n = *n_in_ptr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t npix = s.width * s.height;
    for (uint32_t base = warp_global * 32u; base < n; base += warps_total * 32u) {
        uint32_t i = base + lane;
        int n_out = 0;
        PathStep ps;
        uint32_t slot = 0, branch = 0;
        if (i < n) {
            f4 a = qi.a[i], b = qi.b[i], c = qi.c[i];
            float2 h = hits[i];
            Ray r; r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
            vec3 thr = mk3(b.z, b.w, c.x);
            slot = (uint32_t)f2i(c.y); branch = (uint32_t)f2i(c.z);
            uint32_t pixel = slot % npix, sample = sample0 + slot / npix;
            ps = path_vertex(s, seed, pixel, sample, d, branch, r, thr, h.x, __float_as_int(h.y), glass_branch);
            if (ps.action == PATH_TERMINATE) {
                if (ps.radiance.x != 0.f) atomicAdd(&L[slot].x, ps.radiance.x);
                if (ps.radiance.y != 0.f) atomicAdd(&L[slot].y, ps.radiance.y);
                if (ps.radiance.z != 0.f) atomicAdd(&L[slot].z, ps.radiance.z);
            } else n_out = ps.action == PATH_SPLIT ? 2 : 1;
        }
        // warp-aggregated allocation in the output queue
        uint32_t m1 = __ballot_sync(0xffffffffu, n_out >= 1), m2 = __ballot_sync(0xffffffffu, n_out == 2);
        uint32_t total = __popc(m1) + __popc(m2);
        uint32_t start = 0;
        if (lane == 0 && total) start = atomicAdd(n_out_ptr, total);
        start = __shfl_sync(0xffffffffu, start, 0);
        uint32_t lt = (1u << lane) - 1u;
        if (n_out >= 1) {
            uint32_t pos = start + __popc(m1 & lt);
            if (pos < out_capacity) {
                qo.a[pos] = mk4(ps.next.o.x, ps.next.o.y, ps.next.o.z, ps.next.d.x);
                qo.b[pos] = mk4(ps.next.d.y, ps.next.d.z, ps.thr.x, ps.thr.y);
                qo.c[pos] = mk4(ps.thr.z, i2f((int)slot), i2f((int)branch), 0.f);
            }
        }
        if (n_out == 2) {
            uint32_t pos = start + __popc(m1) + __popc(m2 & lt);
            if (pos < out_capacity) {
                qo.a[pos] = mk4(ps.next2.o.x, ps.next2.o.y, ps.next2.o.z, ps.next2.d.x);
                qo.b[pos] = mk4(ps.next2.d.y, ps.next2.d.z, ps.thr2.x, ps.thr2.y);
                qo.c[pos] = mk4(ps.thr2.z, i2f((int)slot), i2f((int)(branch | (1u << (d & 31u)))), 0.f);
            }
        }
    }
}

// The shade kernel may have tried to allocate past the queue capacity (glass branch mode only).
__global__ void k_clamp_count(uint32_t* n_ptr, uint32_t capacity, uint32_t* high_water) {
    uint32_t n = *n_ptr;
    if (n > capacity) { n = capacity; *n_ptr = n; }
    if (n > *high_water) *high_water = n;
}

// End of wave: accum[p].rgb += L[s*npix + p] for the k samples of the wave in sample order (fp32,
// the reference's `color += trace(...)`, AccPathTracer.cpp:30), accum[p].a += k.
__global__ void k_accumulate(const f4* L, f4* accum, uint32_t npix, uint32_t k) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    f4 acc = accum[p];
    for (uint32_t s = 0; s < k; s++) {
        f4 v = L[(size_t)s * npix + p];
        acc.x += v.x; acc.y += v.y; acc.z += v.z;
    }
    acc.w += (float)k;
    accum[p] = acc;
}

// color /= samples; gamma = sqrt (AccPathTracer.cpp:14-16, 32-34); alpha 1.
__global__ void k_resolve(const f4* accum, f4* rgba, uint32_t npix) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    f4 a = accum[p];
    rgba[p] = mk4(sqrtf(a.x / a.w), sqrtf(a.y / a.w), sqrtf(a.z / a.w), 1.f);
}

__global__ void k_pack_rays(const float* rays6, uint32_t n, PathQueue q) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = rays6 + 6 * (size_t)i;
    q.a[i] = mk4(r[0], r[1], r[2], r[3]);
    q.b[i] = mk4(r[4], r[5], 1.f, 1.f);
}

}  // namespace nrcu
