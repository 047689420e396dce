// nrcu_kernels.cuh — the __global__ kernels (sm_100a).  Thin wrappers around the
// __host__ __device__ bodies in nrcu_{intersect,shade,bvh}.cuh plus what only exists on the GPU:
// persistent-thread work fetching, shared-memory traversal stacks, warp-aggregated queue compaction.
#pragma once
#include <cuda_runtime.h>
#include "nrcu_bvh.cuh"
#include "nrcu_prep.cuh"
#include "nrcu_shade.cuh"
#include "nrcu_mlt.cuh"

namespace nrcu {

// ---------------------------------------------------------------------------------------------
// scene preparation (bodies in nrcu_prep.cuh)
// ---------------------------------------------------------------------------------------------
__global__ void k_mesh_transform(float* pos, uint32_t first_vertex, uint32_t n_vertices) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_vertices) mesh_transform_vertex(pos, first_vertex + i);
}
__global__ void k_build_prims(PrimSources ps, uint32_t n, int raycast, f4* geom, f4* shade, f4* box, f4* bound, uint32_t* meta, float* export16) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) build_prim(ps, i, raycast, geom, shade, box, bound, meta, export16);
}

// environment-map importance-sampling tables (bodies in nrcu_shade.cuh): one thread per row, then one thread
__global__ void k_env_rows(const f4* rgba, int w, int h, float* tab) {
    int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y < h) env_table_row(rgba, w, h, y, tab);
}
__global__ void k_env_marginal(int h, float* tab, float* total_out) { *total_out = env_table_marginal(h, tab); }

// ---------------------------------------------------------------------------------------------
// BVH build: one kernel per step body
// ---------------------------------------------------------------------------------------------
#define NRCU_STEP_KERNEL(name, body) \
    __global__ void name(BvhBuild b, int first, int count) { \
        int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= count) return; body(b, first + i); }
NRCU_STEP_KERNEL(k_bvh_clear, node_clear)
NRCU_STEP_KERNEL(k_bvh_init_prim, bvh_init_prim)
NRCU_STEP_KERNEL(k_bvh_big_candidate, bvh_big_candidate)
NRCU_STEP_KERNEL(k_bvh_select_big, bvh_select_big)
NRCU_STEP_KERNEL(k_bvh_init_prim_rest, bvh_init_prim_rest)
NRCU_STEP_KERNEL(k_bvh_level_prepare, bvh_level_prepare)
NRCU_STEP_KERNEL(k_bvh_bin, bvh_bin)
NRCU_STEP_KERNEL(k_bvh_split, bvh_split)
NRCU_STEP_KERNEL(k_bvh_partition, bvh_partition)
NRCU_STEP_KERNEL(k_bvh_leaf_alloc, bvh_leaf_alloc)
NRCU_STEP_KERNEL(k_bvh_leaf_fill, bvh_leaf_fill)
NRCU_STEP_KERNEL(k_bvh_leaf_sort, bvh_leaf_sort)
NRCU_STEP_KERNEL(k_bvh_leaf_gather, bvh_leaf_gather)
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
// A queue entry is 40 bytes in three arrays:  a = float4 (o.xyz, d.x)   b = float2 (d.y, d.z)   c = float4 (thr.xyz, slot)
// (+ d = uint32 glass-branch bits, only allocated and touched in the branching glass mode).  The closest-hit kernels
// read a and b (24 B/ray), the shading kernel reads and writes all of it.
// slot = sample_in_wave * n_pixels + pixel identifies the path; its radiance lands in L[slot].
struct PathQueue { f4* a; float2* b; f4* c; uint32_t* d; };

// QUEUE REGIONS.  Every warp of the shading kernel allocates its output entries with one lane-0 atomicAdd on the queue's
// size counter; the L2 serialises atomics per address (1.5/ns, tools/micro/atomic_bench.cu) and everything else that is
// routed to that slice waits behind them (ncu source page of k_shade: the two hottest stall sites are the shuffle that
// waits for the atomic and the first use of the prefetched queue entry).  So a queue is split into K = 2^logk REGIONS, each
// with a counter on a line of its own; the shading warp that works on input block pb (32 entries) appends to region
// pb mod K.  Regions are interleaved in the array by blocks of 32 entries - entry p of region r lives at
// ((p / 32) K + r) 32 + p mod 32 - so the queue stays ONE flat array that consumers walk block by block with their
// software pipeline intact; a block's live lanes are those below its region's count.  The regions receive every K-th
// block's survivors and stay equally long up to a few blocks, so the only dead lanes are those of the last blocks.
// K = 1 is the plain compacted queue.
struct QRegions { const uint32_t* cnt; uint32_t logk, cs, max_blocks; };   // count of region r at cnt[r * cs]; the array holds max_blocks blocks
// lane r keeps region r's count; returns the number of 32-entry blocks the queue spans in the array
__device__ __forceinline__ uint32_t regions_begin(const QRegions& qr, uint32_t lane, uint32_t& my_cnt) {
    my_cnt = lane < (1u << qr.logk) ? qr.cnt[(size_t)lane * qr.cs] : 0u;
    const uint32_t b = (my_cnt + 31u) >> 5;
    uint32_t ext = b ? ((b - 1u) << qr.logk) + lane + 1u : 0u;   // region r's last block sits at array block (b - 1) K + r
    for (int o = 16; o > 0; o >>= 1) ext = max(ext, __shfl_xor_sync(0xffffffffu, ext, o));
    return min(ext, qr.max_blocks);   // a producer that ran out of room has counted what it dropped (DScene::overflow); never read past the array
}
// is lane `lane` of array block pb a live entry?
__device__ __forceinline__ bool region_live(const QRegions& qr, uint32_t my_cnt, uint32_t pb, uint32_t lane) {
    const uint32_t c = __shfl_sync(0xffffffffu, my_cnt, pb & ((1u << qr.logk) - 1u));
    return ((pb >> qr.logk) << 5) + lane < c;
}

// The wide-primitive list staged in shared memory (read as warp-wide broadcasts).
struct BigList {
    f4 g[NRCU_MAX_BIG * 3]; f4 b[NRCU_MAX_BIG * 2]; f4 bd[NRCU_MAX_BIG * 2]; uint32_t m[NRCU_MAX_BIG];
    __device__ __forceinline__ void load(const DScene& s) {
        for (uint32_t k = threadIdx.x; k < s.n_big * 3; k += blockDim.x) g[k] = s.big_geom[k];
        for (uint32_t k = threadIdx.x; k < s.n_big * 2; k += blockDim.x) { b[k] = s.big_box[k]; bd[k] = s.big_bound[k]; }
        for (uint32_t k = threadIdx.x; k < s.n_big; k += blockDim.x) m[k] = s.big_meta[k];
        __syncthreads();
    }
};
// Closest hit, stage 1, for a ray that is still in registers: every lane walks the same short list of wide
// primitives (no traversal divergence), stores the provisional hit of queue entry `pos` and reports whether
// anything inside the BVH could still be closer (conservative test against the BVH bounds).
template <bool GATE>
__device__ __forceinline__ bool stage1(const DScene& s, const BigList& bl, const Ray& r, uint32_t pos, float2* hits) {
    RayPrep rp = prep_ray(r);
    float best_t = NRCU_INF; int best_id = -1;
    big_list_step<GATE>(s, bl.g, bl.b, bl.bd, bl.m, r, rp, gate_inverse(r, rp), best_t, best_id);
    hits[pos] = make_float2(best_t, __int_as_float(best_id));
    return bvh_reachable(s, rp, best_t);
}
// The same for a camera ray of film position (x, y): the candidates come from the film rectangles when the scene has them.
template <bool GATE>
__device__ __forceinline__ bool stage1_camera(const DScene& s, const BigList& bl, const f4* rects, const Ray& r, float x, float y, uint32_t pos, float2* hits) {
    RayPrep rp = prep_ray(r);
    float best_t = NRCU_INF; int best_id = -1;
    const uint32_t mask = rects ? big_list_mask_film(s, rects, x, y) : big_list_mask(s, bl.bd, rp);
    big_list_resolve<GATE>(mask, bl.g, bl.b, bl.m, r, gate_inverse(r, rp), best_t, best_id);
    hits[pos] = make_float2(best_t, __int_as_float(best_id));
    return bvh_reachable(s, rp, best_t);
}
// One thread per wide primitive: DScene::big_rect (see big_list_mask_film).  A film point (x, y) sends its ray along
// A + x H + y V with A = lower_left - position; P - position = l (A + x H + y V) is solved for (l, l x, l y) by Cramer's rule.
__device__ f4 film_rect_of_box(const DScene& s, const double lo[3], const double hi[3]);
__global__ void k_big_rects(DScene s, f4* rect) {
    const uint32_t k = threadIdx.x;
    if (k >= s.n_big) return;
    const f4 bc = s.big_bound[2 * k], bh = s.big_bound[2 * k + 1];
    const double lo[3] = {(double)bc.x - (double)bh.x, (double)bc.y - (double)bh.y, (double)bc.z - (double)bh.z};
    const double hi[3] = {(double)bc.x + (double)bh.x, (double)bc.y + (double)bh.y, (double)bc.z + (double)bh.z};
    rect[k] = film_rect_of_box(s, lo, hi);
}
__device__ f4 film_rect_of_box(const DScene& s, const double lo[3], const double hi[3]) {
    const double px = s.cam.position.x, py = s.cam.position.y, pz = s.cam.position.z;
    const double A[3] = {s.cam.lower_left.x - px, s.cam.lower_left.y - py, s.cam.lower_left.z - pz};
    const double H[3] = {s.cam.horizontal.x, s.cam.horizontal.y, s.cam.horizontal.z}, V[3] = {s.cam.vertical.x, s.cam.vertical.y, s.cam.vertical.z};
    auto det3 = [](const double* a, const double* b, const double* c) {
        return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
    };
    const double det = det3(A, H, V);
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    bool whole = !(fabs(det) > 1e-300);
    double scale = 0.0;
    for (int c = 0; c < 8 && !whole; c++) {
        const double q[3] = {((c & 1) ? hi[0] : lo[0]) - px, ((c & 2) ? hi[1] : lo[1]) - py, ((c & 4) ? hi[2] : lo[2]) - pz};
        const double l = det3(q, H, V) / det, lx = det3(A, q, V) / det, ly = det3(A, H, q) / det;
        scale = fmax(scale, fmax(fabs(l), fmax(fabs(lx), fabs(ly))));
        if (!(l > 1e-6 * scale) || !(l == l)) { whole = true; break; }   // a corner beside or behind the camera: no finite rectangle
        const double x = lx / l, y = ly / l;
        x0 = fmin(x0, x); x1 = fmax(x1, x); y0 = fmin(y0, y); y1 = fmax(y1, y);
    }
    // margin: the fp32 ray direction of film point (x, y) differs from the exact one by a few 1e-7 of the film size
    const double mg = 1e-3 * (1.0 + fmax(fmax(fabs(x0), fabs(x1)), fmax(fabs(y0), fabs(y1))));
    if (whole || !(x0 <= x1) || !(y0 <= y1)) return mk4(-NRCU_INF, NRCU_INF, -NRCU_INF, NRCU_INF);
    return mk4((float)(x0 - mg), (float)(x1 + mg), (float)(y0 - mg), (float)(y1 + mg));
}
// Live pixels (DScene::live_px): a pixel is live iff the film footprint of its jittered samples - the pixel corner +- one
// pixel, AccPathTracer.cpp:23-29 - overlaps the film rectangle of a wide primitive, of the BVH's bounds or of an area light
// (closestHitLight is asked for every ray, AccPathTracer.cpp:128).
#define NRCU_LIVE_LIGHTS 32
#define NRCU_LIVE_RECTS (NRCU_MAX_BIG + 1 + NRCU_LIVE_LIGHTS)
// rect[n_big] = the BVH's bounds, rect[n_big + 1 + i] = light i (empty rectangles where there is nothing); one thread each
__global__ void k_scene_rects(DScene s, f4* rect) {
    const uint32_t t = threadIdx.x;
    if (t > min(s.n_area_lights, (uint32_t)NRCU_LIVE_LIGHTS)) return;
    f4 r = mk4(NRCU_INF, -NRCU_INF, NRCU_INF, -NRCU_INF);   // empty
    if (t == 0) {
        if (s.root_ref != NRCU_REF_EMPTY) {
            const double lo[3] = {s.bvh_lo.x, s.bvh_lo.y, s.bvh_lo.z}, hi[3] = {s.bvh_hi.x, s.bvh_hi.y, s.bvh_hi.z};
            r = film_rect_of_box(s, lo, hi);
        }
    } else {   // a light quad p + a u + b v: the box of its four corners (record layout: nrcu_host_prep.hpp)
        const f4* L = s.area_lights + NRCU_LIGHT_F4 * (size_t)(t - 1u);
        const f4 l0 = L[0], l1 = L[1], lu = L[4], lv = L[5];
        const double p[3] = {l0.w, l1.x, l1.y}, u[3] = {lu.x, lu.y, lu.z}, v[3] = {lv.x, lv.y, lv.z};
        double lo[3], hi[3];
        for (int k = 0; k < 3; k++) {
            const double c0 = p[k], c1 = p[k] + u[k], c2 = p[k] + v[k], c3 = p[k] + u[k] + v[k];
            const double pad = 1e-4 * (1.0 + fabs(c0) + fabs(u[k]) + fabs(v[k]));
            lo[k] = fmin(fmin(c0, c1), fmin(c2, c3)) - pad; hi[k] = fmax(fmax(c0, c1), fmax(c2, c3)) + pad;
        }
        r = film_rect_of_box(s, lo, hi);
    }
    rect[s.n_big + t] = r;
}
// one thread per pixel: flag + the number of live pixels of the block (for the ordered compaction)
__global__ void __launch_bounds__(256) k_live_flags(DScene s, const f4* rect, uint32_t n_rect, int all_live, unsigned char* flag, uint32_t* block_count) {
    __shared__ f4 rc[NRCU_LIVE_RECTS];
    for (uint32_t k = threadIdx.x; k < n_rect; k += blockDim.x) rc[k] = rect[k];
    __syncthreads();
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    bool live = false;
    if (p < s.width * s.height) {
        const int w = (int)s.width, h = (int)s.height;
        const int row = (int)(p / (uint32_t)w), j = (int)(p % (uint32_t)w), i = h - 1 - row;   // pt_camera_ray's pixel coordinates
        const float fx0 = ((float)j - 1.f) / (float)w, fx1 = ((float)j + 1.f) / (float)w, fy0 = ((float)i - 1.f) / (float)h, fy1 = ((float)i + 1.f) / (float)h;
        live = all_live != 0;
        for (uint32_t k = 0; k < n_rect && !live; k++) live = fx1 >= rc[k].x && fx0 <= rc[k].y && fy1 >= rc[k].z && fy0 <= rc[k].w;
        flag[p] = live ? 1 : 0;
    }
    const int c = __syncthreads_count(live ? 1 : 0);
    if (threadIdx.x == 0) block_count[blockIdx.x] = (uint32_t)c;
}
// exclusive scan of the block counts (one block; a thread owns a contiguous range) -> block offsets, total
__global__ void __launch_bounds__(1024) k_live_scan(uint32_t* block_count, uint32_t n_blocks, uint32_t* n_live) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (n_blocks + 1023u) / 1024u, lo = min(n_blocks, threadIdx.x * per), hi = min(n_blocks, lo + per);
    uint32_t c = 0;
    for (uint32_t b = lo; b < hi; b++) c += block_count[b];
    part[threadIdx.x] = c;
    __syncthreads();
    for (uint32_t o = 1; o < 1024u; o <<= 1) {   // Hillis-Steele inclusive scan
        const uint32_t v = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - c;
    for (uint32_t b = lo; b < hi; b++) { const uint32_t v = block_count[b]; block_count[b] = run; run += v; }
    if (threadIdx.x == 1023u) *n_live = part[1023];
}
// live pixels of a block, in pixel order, behind the block's offset (want = 0: the dead pixels; block b then holds
// min(256, npix - 256 b) - live count of them, so its offset is 256 b - the live offset)
__global__ void __launch_bounds__(256) k_live_scatter(const unsigned char* flag, uint32_t npix, const uint32_t* block_offset, uint32_t* list, int want) {
    __shared__ uint32_t warp_base[8];
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const bool live = p < npix && (flag[p] != 0) == (want != 0);
    const uint32_t m = __ballot_sync(0xffffffffu, live);
    if (lane == 0) warp_base[wib] = __popc(m);
    __syncthreads();
    uint32_t base = want ? block_offset[blockIdx.x] : blockIdx.x * blockDim.x - block_offset[blockIdx.x];
    for (uint32_t k = 0; k < wib; k++) base += warp_base[k];
    if (live) list[base + __popc(m & ((1u << lane) - 1u))] = p;
}
// Environment-map scenes: the camera rays of the dead pixels leave the scene for certain, so their radiance is the map along
// the ray (path_vertex's miss branch at bounce 0: throughput 1) - written straight into the radiance slot, one thread per
// (dead pixel, sample of the wave); no queue entry, no closest-hit query, no shading pass.
__global__ void __launch_bounds__(256) k_env_dead(DScene s, uint64_t seed, uint32_t sample0, uint32_t n_entries, f4* L) {
    const uint32_t npix = s.width * s.height, n_dead = npix - s.n_live;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_entries; i += gridDim.x * blockDim.x) {
        const uint32_t j = i % n_dead, sw = i / n_dead, pixel = s.dead_px[j];
        const Ray r = pt_camera_ray(s, seed, pixel, sample0 + sw);
        const vec3 e = mk3(1.f, 1.f, 1.f) * env_lookup(s, r.d);
        L[(size_t)sw * npix + pixel] = mk4(e.x, e.y, e.z, 0.f);
    }
}
// Append the flagged lanes' queue positions to the survivor list with one atomic per warp.
__device__ __forceinline__ void append_survivors(bool more, uint32_t pos, uint32_t* surv, uint32_t* n_surv) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t m = __ballot_sync(0xffffffffu, more);
    if (m == 0) return;
    uint32_t start = 0;
    if (lane == 0) start = atomicAdd(n_surv, (uint32_t)__popc(m));
    start = __shfl_sync(0xffffffffu, start, 0);
    if (more) surv[start + __popc(m & ((1u << lane) - 1u))] = pos;
}

// Camera rays of one wave (slot = sample_in_wave * n_pixels + pixel) (+ STAGE1: stage 1 of their closest hit).
template <bool GATE, bool STAGE1>
__global__ void __launch_bounds__(256) k_raygen(DScene s, uint64_t seed, uint32_t sample0, uint32_t n_slots, PathQueue q, f4* L, uint32_t* n_queue,
                                               float2* hits, uint32_t* surv, uint32_t* n_surv, unsigned long long* ray_counter) {
    __shared__ BigList bl;
    __shared__ f4 film_rect[NRCU_MAX_BIG];
    if (STAGE1) {
        if (s.big_rect) for (uint32_t k = threadIdx.x; k < s.n_big; k += blockDim.x) film_rect[k] = s.big_rect[k];
        bl.load(s);
    }
    const f4* rects = s.big_rect ? film_rect : nullptr;
    const uint32_t npix = s.width * s.height;
    const uint32_t stride = gridDim.x * blockDim.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *n_queue = s.depth == 0 ? 0u : n_slots;   // every slot starts one path: the bounce-0 queue is dense
        if (STAGE1 && s.depth != 0 && ray_counter) atomicAdd(ray_counter, (unsigned long long)n_slots);
    }
    // queue entry i = sample_in_wave * n_live + j holds the camera ray of live pixel j; its radiance slot is
    // sample_in_wave * n_pixels + pixel (dead pixels have no entries and no slots written: DScene::live_px)
    const uint32_t n_live = s.live_px ? s.n_live : npix;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_slots; base += stride) {   // warp-uniform trip count
        const uint32_t i = base + threadIdx.x;
        bool more = false;
        if (i < n_slots) {
            const uint32_t j = i % n_live, sw = i / n_live;
            const uint32_t pixel = s.live_px ? s.live_px[j] : j, sample = sample0 + sw;
            const uint32_t slot = sw * npix + pixel;
            if (s.depth == 0) L[slot] = mk4(s.ambient.x, s.ambient.y, s.ambient.z, 0.f);   // trace(): currDepth == depth
            else {
                L[slot] = mk4(0.f, 0.f, 0.f, 0.f);
                float fx, fy;
                Ray r = pt_camera_ray(s, seed, pixel, sample, &fx, &fy);
                q.a[i] = mk4(r.o.x, r.o.y, r.o.z, r.d.x);
                q.b[i] = make_float2(r.d.y, r.d.z);
                q.c[i] = mk4(1.f, 1.f, 1.f, i2f((int)slot));
                if (q.d) q.d[i] = 0u;
                if (STAGE1) more = stage1_camera<GATE>(s, bl, rects, r, fx, fy, i, hits);
            }
        }
        if (STAGE1) append_survivors(more, i, surv, n_surv);
    }
}

#define NRCU_TRACE_THREADS 128

// ---------------------------------------------------------------------------------------------
// Closest hit, stage 1: the wide primitives, every ray, warp-uniform
// ---------------------------------------------------------------------------------------------
// All lanes walk the same short list (<= NRCU_MAX_BIG records staged in shared memory, read as broadcasts),
// so there is no traversal divergence at all; only the early exits inside the exact tests diverge.
// Writes the provisional hit of every ray and appends the rays that can still hit something inside the
// BVH (conservative test against the BVH bounds with the provisional t) to the survivor list with one
// atomic per warp.  On the Cornell-box scenes most rays end here.
template <bool GATE>
__global__ void __launch_bounds__(256) k_big(DScene s, PathQueue q, QRegions rin, float2* hits,
                                            uint32_t* surv, uint32_t* n_surv, unsigned long long* ray_counter) {
    __shared__ BigList bl;
    bl.load(s);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t my_cnt;
    const uint32_t nblocks = regions_begin(rin, lane, my_cnt);
    // software pipeline as in k_shade: the next ray is requested before waiting for the survivor-list atomic
    f4 a = mk4(0, 0, 0, 0); float2 b = make_float2(0.f, 0.f);
    if (warp_global < nblocks) { const uint32_t i0 = warp_global * 32u + lane; a = q.a[i0]; b = q.b[i0]; }
    for (uint32_t pb = warp_global; pb < nblocks; pb += warps_total) {
        const uint32_t i = pb * 32u + lane;
        const bool live = region_live(rin, my_cnt, pb, lane);
        bool more = false;
        if (live) {
            Ray r; r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
            more = stage1<GATE>(s, bl, r, i, hits);
        }
        if (ray_counter) { const uint32_t ml = __ballot_sync(0xffffffffu, live); if (lane == 0) atomicAdd(ray_counter, (unsigned long long)__popc(ml)); }
        const uint32_t m = __ballot_sync(0xffffffffu, more);
        uint32_t start = 0;
        if (lane == 0 && m) start = atomicAdd(n_surv, (uint32_t)__popc(m));
        if (pb + warps_total < nblocks) { const uint32_t inext = i + warps_total * 32u; a = q.a[inext]; b = q.b[inext]; }
        if (m) {
            start = __shfl_sync(0xffffffffu, start, 0);
            if (more) surv[start + __popc(m & ((1u << lane) - 1u))] = i;
        }
    }
}

// k_big with WARP-BALANCED exact tests.  In k_big the exact tests of a warp take as many rounds as its busiest
// lane has candidates (about 5 on incoherent rays, at 2-8 active lanes), although the warp holds only ~2 candidates
// per ray: 18 of 32 lanes per instruction overall (ncu).  Here pass 1 also builds, per warp, the list of (ray, wide
// primitive) candidate pairs in shared memory - primitive-major, by ballot/popc inside the warp-uniform slab loop -
// and the exact tests run 32 pairs at a time: consecutive lanes test the SAME primitive against different rays
// (uniform kind, broadcast reads of the record).  A test reads its ray from shared memory, culls against the owner's
// best hit so far and publishes a hit with a 64-bit shared-memory atomicMin on (t bits, id, list index): t > 0, so
// unsigned order is (t, id) order, the tie rule of the per-ray loop.  The optimistic leaf gate and its
// per-candidate fallback stay with the owning lane.
#define NRCU_BIGB_WARPS 8
#define NRCU_BEST_NONE 0x7f800000ffffffffull
// SLOTS (path-regeneration scheduler): the queue is the slot array of a partition - n_fixed entries, never compacted; a
// slot whose lane has run out of samples holds a NaN origin and is skipped; rays are counted by the shading kernel;
// `n_ptr` then points at the previous iteration's "somebody is still alive" flags (null in the first iteration).
#define NRCU_REGEN_FLAGS 16        // alive flags per iteration, each on a 128-byte line of its own
#define NRCU_REGEN_FLAG_STRIDE 32
__device__ __forceinline__ bool regen_anyone_alive(const uint32_t* flags) {
    if (!flags) return true;
    uint32_t v = 0;
    if (threadIdx.x < NRCU_REGEN_FLAGS) v = flags[threadIdx.x * NRCU_REGEN_FLAG_STRIDE];
    return __syncthreads_or((int)v) != 0;
}
#ifndef NRCU_BIGB_MINB
#define NRCU_BIGB_MINB 4
#endif
template <bool GATE, bool SLOTS>
__global__ void __launch_bounds__(32 * NRCU_BIGB_WARPS, NRCU_BIGB_MINB) k_big_balanced(DScene s, PathQueue q, const uint32_t* n_ptr, float2* hits,
                                                                       uint32_t* surv, uint32_t* n_surv, unsigned long long* ray_counter, uint32_t n_fixed,
                                                                       uint32_t in_logk, uint32_t in_cs, uint32_t in_max_blocks) {
    __shared__ BigList bl;
    __shared__ unsigned short pairs[NRCU_BIGB_WARPS][32 * NRCU_MAX_BIG];
    __shared__ float rays[NRCU_BIGB_WARPS][6][32];
    __shared__ unsigned long long best[NRCU_BIGB_WARPS][32];
    if (SLOTS && !regen_anyone_alive(n_ptr)) return;
    bl.load(s);
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // the wavefront's queue comes in regions (QRegions above); the slot array of the regeneration scheduler is one dense range
    const QRegions rin = {n_ptr, SLOTS ? 0u : in_logk, in_cs, in_max_blocks};
    uint32_t my_cnt = 0;
    const uint32_t n = SLOTS ? n_fixed : regions_begin(rin, lane, my_cnt) * 32u;   // array extent in entries (whole blocks)
    unsigned short* my_pairs = pairs[wib];
    f4 a = mk4(0, 0, 0, 0); float2 b = make_float2(0.f, 0.f);
    { const uint32_t i0 = warp_global * 32u + lane; if (i0 < n) { a = q.a[i0]; b = q.b[i0]; } }
    for (uint32_t base = warp_global * 32u; base < n; base += warps_total * 32u) {
        const uint32_t i = base + lane;
        const bool live = SLOTS ? (i < n && a.x == a.x) : region_live(rin, my_cnt, base >> 5, lane);
        if (SLOTS && !__any_sync(0xffffffffu, live)) {   // a warp of finished slots (end of the frame): only keep the pipeline going
            const uint32_t inext = i + warps_total * 32u; if (inext < n) { a = q.a[inext]; b = q.b[inext]; }
            continue;
        }
        Ray r; r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
        const RayPrep rp = prep_ray(r);
        rays[wib][0][lane] = r.o.x; rays[wib][1][lane] = r.o.y; rays[wib][2][lane] = r.o.z;
        rays[wib][3][lane] = r.d.x; rays[wib][4][lane] = r.d.y; rays[wib][5][lane] = r.d.z;
        best[wib][lane] = NRCU_BEST_NONE;   // t = +inf, id = all ones: reads back as "nothing yet" without a special case
        // pass 1: slab test of every wide primitive (warp-uniform loop, broadcast reads) -> candidate pairs, primitive-major
        uint32_t total = 0;
        // Lanes past the end of the queue: an x offset of -inf puts near and far at -inf, so tf = -inf < 0 <= tn and the
        // lane never becomes a candidate (prep_ray clamps 1/d to +-1e18, the products stay finite: no NaN) - one
        // predicate less per primitive than testing i < n inside the loop.
        const float nox = live ? -rp.oinv.x : -NRCU_INF;
        const vec3 ainv = mk3(fabsf(rp.inv.x), fabsf(rp.inv.y), fabsf(rp.inv.z));
        for (uint32_t k = 0; k < s.n_big; k++) {
            float tn, tf;
            slab_center_extent(bl.bd[2 * k], bl.bd[2 * k + 1], rp, ainv, nox, tn, tf);
            const bool cand = tn <= tf;
            const uint32_t m = __ballot_sync(0xffffffffu, cand);
            if (cand) my_pairs[total + __popc(m & lt)] = (unsigned short)(lane | (k << 5));
            total += __popc(m);
        }
        { const uint32_t inext = i + warps_total * 32u; if (inext < n) { a = q.a[inext]; b = q.b[inext]; } }   // prefetch
        __syncwarp();
        // pass 2: 32 exact tests per round
        for (uint32_t jb = 0; jb < total; jb += 32u) {
            const uint32_t j = jb + lane;
            if (j < total) {
                const uint32_t p = my_pairs[j], ol = p & 31u, k = p >> 5;
                Ray pr; pr.o = mk3(rays[wib][0][ol], rays[wib][1][ol], rays[wib][2][ol]); pr.d = mk3(rays[wib][3][ol], rays[wib][4][ol], rays[wib][5][ol]);
                float bt = __uint_as_float((uint32_t)(best[wib][ol] >> 32));   // the owner's best so far (+inf: none): only culls
                int bi = 0x7fffffff;                                                  // accept ties: atomicMin settles them by id
                prim_test<false>(pr, mk3(0.f), bl.g[3 * k], bl.g[3 * k + 1], bl.g[3 * k + 2], bl.b, bl.m[k], bt, bi);
                if (bi != 0x7fffffff) atomicMin(&best[wib][ol], ((unsigned long long)__float_as_uint(bt) << 32) | (unsigned long long)(((uint32_t)bi << 5) | k));
            }
        }
        __syncwarp();
        bool more = false;
        if (live) {
            const unsigned long long key = best[wib][lane];
            float best_t = NRCU_INF; int best_id = -1;
            if (key != NRCU_BEST_NONE) {
                best_t = __uint_as_float((uint32_t)(key >> 32)); best_id = (int)((uint32_t)key >> 5);
                if (GATE) {
                    const uint32_t kb = (uint32_t)key & 31u;
                    const vec3 ginv = gate_inverse(r, rp);
                    if (!bounds_intersectp_inv(bl.b[2 * kb], bl.b[2 * kb + 1], r, ginv.x, ginv.y, ginv.z)) {   // rare: gate per candidate
                        best_t = NRCU_INF; best_id = -1;
                        // every wide primitive again, each behind its own gate.  The slab pre-test only ever removes
                        // primitives the exact test rejects, so walking the whole list gives the same answer - and not
                        // carrying a per-lane candidate mask through pass 1 for this rare path is worth 3.4 % of the frame
                        for (uint32_t k = 0; k < s.n_big; k++)
                            prim_test<true>(r, ginv, bl.g[3 * k], bl.g[3 * k + 1], bl.g[3 * k + 2], bl.b + 2 * k, bl.m[k], best_t, best_id);
                    }
                }
            }
            hits[i] = make_float2(best_t, __int_as_float(best_id));
            more = bvh_reachable(s, rp, best_t);
        }
        __syncwarp();   // the shared lists are rewritten by the next iteration
        if (!SLOTS && ray_counter) { const uint32_t ml = __ballot_sync(0xffffffffu, live); if (lane == 0) atomicAdd(ray_counter, (unsigned long long)__popc(ml)); }
        append_survivors(more, i, surv, n_surv);
    }
}

// k_big_balanced with TWO rays per lane (64 rays per warp iteration).  Same passes, same arithmetic, same answers; what
// changes is the shape of the work:
//   * pass 1 tests one wide primitive against two independent rays per lane: the record is read once (2 LDS.128 per 64
//     rays), and the two slab tests interleave in the pipeline (the kernel ran at 77 % issue-active whether 3 or 4 CTAs
//     were resident per SM, i.e. it waits on its own dependent chains, not for warps);
//   * the pair list of 64 rays (~135 pairs) fills the 32-wide rounds of pass 2 better than two lists of ~67 (2.75 rounds
//     per 32 rays -> ~2.35), and the loop overhead, the prefetch and the survivor append are paid once per 64 rays.
// Four warps per CTA: the pair list of a warp is 64 x NRCU_MAX_BIG entries.
#define NRCU_BIG64_WARPS 4
#ifndef NRCU_BIG64_MINB
#define NRCU_BIG64_MINB 6
#endif
template <bool GATE>
__global__ void __launch_bounds__(32 * NRCU_BIG64_WARPS, NRCU_BIG64_MINB) k_big_balanced64(DScene s, PathQueue q, QRegions rin, float2* hits,
                                                                                         uint32_t* surv, uint32_t* n_surv) {
    __shared__ BigList bl;
    __shared__ unsigned short pairs[NRCU_BIG64_WARPS][64 * NRCU_MAX_BIG];
    __shared__ float rays[NRCU_BIG64_WARPS][6][64];
    __shared__ unsigned long long best[NRCU_BIG64_WARPS][64];
    bl.load(s);
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t my_cnt;
    const uint32_t n = regions_begin(rin, lane, my_cnt) * 32u;   // array extent in entries (whole blocks of 32)
    unsigned short* my_pairs = pairs[wib];
    f4 a0 = mk4(0, 0, 0, 0), a1 = a0; float2 b0 = make_float2(0.f, 0.f), b1 = b0;
    {
        const uint32_t i0 = warp_global * 64u + lane;
        if (i0 < n) { a0 = q.a[i0]; b0 = q.b[i0]; }
        if (i0 + 32u < n) { a1 = q.a[i0 + 32u]; b1 = q.b[i0 + 32u]; }
    }
    for (uint32_t base = warp_global * 64u; base < n; base += warps_total * 64u) {
        const uint32_t i0 = base + lane, i1 = i0 + 32u;
        const bool live0 = region_live(rin, my_cnt, base >> 5, lane), live1 = region_live(rin, my_cnt, (base >> 5) + 1u, lane);   // a block past the extent has no live lane
        Ray r0, r1;
        r0.o = mk3(a0.x, a0.y, a0.z); r0.d = mk3(a0.w, b0.x, b0.y);
        r1.o = mk3(a1.x, a1.y, a1.z); r1.d = mk3(a1.w, b1.x, b1.y);
        const RayPrep rp0 = prep_ray(r0), rp1 = prep_ray(r1);
        rays[wib][0][lane] = r0.o.x; rays[wib][1][lane] = r0.o.y; rays[wib][2][lane] = r0.o.z;
        rays[wib][3][lane] = r0.d.x; rays[wib][4][lane] = r0.d.y; rays[wib][5][lane] = r0.d.z;
        rays[wib][0][lane + 32] = r1.o.x; rays[wib][1][lane + 32] = r1.o.y; rays[wib][2][lane + 32] = r1.o.z;
        rays[wib][3][lane + 32] = r1.d.x; rays[wib][4][lane + 32] = r1.d.y; rays[wib][5][lane + 32] = r1.d.z;
        best[wib][lane] = NRCU_BEST_NONE; best[wib][lane + 32] = NRCU_BEST_NONE;
        // pass 1 (see k_big_balanced): an x offset of -inf keeps a lane without a ray out of the candidate lists
        uint32_t total = 0;
        const float nox0 = live0 ? -rp0.oinv.x : -NRCU_INF, nox1 = live1 ? -rp1.oinv.x : -NRCU_INF;
        const vec3 ainv0 = mk3(fabsf(rp0.inv.x), fabsf(rp0.inv.y), fabsf(rp0.inv.z)), ainv1 = mk3(fabsf(rp1.inv.x), fabsf(rp1.inv.y), fabsf(rp1.inv.z));
        for (uint32_t k = 0; k < s.n_big; k++) {
            const f4 bc = bl.bd[2 * k], bh = bl.bd[2 * k + 1];
            float tn0, tf0, tn1, tf1;
            slab_center_extent(bc, bh, rp0, ainv0, nox0, tn0, tf0);
            slab_center_extent(bc, bh, rp1, ainv1, nox1, tn1, tf1);
            const bool c0 = tn0 <= tf0, c1 = tn1 <= tf1;
            const uint32_t m0 = __ballot_sync(0xffffffffu, c0), m1 = __ballot_sync(0xffffffffu, c1);
            const uint32_t n0 = __popc(m0);
            if (c0) my_pairs[total + __popc(m0 & lt)] = (unsigned short)(lane | (k << 6));
            if (c1) my_pairs[total + n0 + __popc(m1 & lt)] = (unsigned short)((lane + 32u) | (k << 6));
            total += n0 + __popc(m1);
        }
        {   // prefetch the next 64 rays
            const uint32_t j0 = i0 + warps_total * 64u;
            if (j0 < n) { a0 = q.a[j0]; b0 = q.b[j0]; }
            if (j0 + 32u < n) { a1 = q.a[j0 + 32u]; b1 = q.b[j0 + 32u]; }
        }
        __syncwarp();
        // pass 2: 32 exact tests per round
        for (uint32_t jb = 0; jb < total; jb += 32u) {
            const uint32_t j = jb + lane;
            if (j < total) {
                const uint32_t p = my_pairs[j], ol = p & 63u, k = p >> 6;
                Ray pr; pr.o = mk3(rays[wib][0][ol], rays[wib][1][ol], rays[wib][2][ol]); pr.d = mk3(rays[wib][3][ol], rays[wib][4][ol], rays[wib][5][ol]);
                float bt = __uint_as_float((uint32_t)(best[wib][ol] >> 32));
                int bi = 0x7fffffff;
                prim_test<false>(pr, mk3(0.f), bl.g[3 * k], bl.g[3 * k + 1], bl.g[3 * k + 2], bl.b, bl.m[k], bt, bi);
                if (bi != 0x7fffffff) atomicMin(&best[wib][ol], ((unsigned long long)__float_as_uint(bt) << 32) | (unsigned long long)(((uint32_t)bi << 5) | k));
            }
        }
        __syncwarp();
        bool more0 = false, more1 = false;
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {
            const bool live = hlf ? live1 : live0;
            if (live) {
                const Ray& r = hlf ? r1 : r0;
                const RayPrep& rp = hlf ? rp1 : rp0;
                const unsigned long long key = best[wib][lane + 32 * hlf];
                float best_t = NRCU_INF; int best_id = -1;
                if (key != NRCU_BEST_NONE) {
                    best_t = __uint_as_float((uint32_t)(key >> 32)); best_id = (int)((uint32_t)key >> 5);
                    if (GATE) {
                        const uint32_t kb = (uint32_t)key & 31u;
                        const vec3 ginv = gate_inverse(r, rp);
                        if (!bounds_intersectp_inv(bl.b[2 * kb], bl.b[2 * kb + 1], r, ginv.x, ginv.y, ginv.z)) {   // rare: gate per candidate
                            best_t = NRCU_INF; best_id = -1;
                            for (uint32_t k = 0; k < s.n_big; k++)
                                prim_test<true>(r, ginv, bl.g[3 * k], bl.g[3 * k + 1], bl.g[3 * k + 2], bl.b + 2 * k, bl.m[k], best_t, best_id);
                        }
                    }
                }
                hits[hlf ? i1 : i0] = make_float2(best_t, __int_as_float(best_id));
                const bool more = bvh_reachable(s, rp, best_t);
                if (hlf) more1 = more; else more0 = more;
            }
        }
        __syncwarp();   // the shared lists are rewritten by the next iteration
        {   // both halves' survivors with one atomic
            const uint32_t m0 = __ballot_sync(0xffffffffu, more0), m1 = __ballot_sync(0xffffffffu, more1);
            if (m0 | m1) {
                uint32_t start = 0;
                if (lane == 0) start = atomicAdd(n_surv, (uint32_t)(__popc(m0) + __popc(m1)));
                start = __shfl_sync(0xffffffffu, start, 0);
                if (more0) surv[start + __popc(m0 & lt)] = i0;
                if (more1) surv[start + __popc(m0) + __popc(m1 & lt)] = i1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Closest hit, stage 2: BVH4 traversal of the surviving rays
// ---------------------------------------------------------------------------------------------
// Traversal stack: 16 (t, ref) entries per thread in shared memory ([entry][thread], conflict free), indexed by
// a register; deeper entries overflow to local memory (only very deep trees get there).
#define NRCU_T3_STACK 16
struct Stack3 {
    uint2* smem;                                   // this thread's column of the shared stack
    float ot[NRCU_LOCAL_STACK - NRCU_T3_STACK];
    int oref[NRCU_LOCAL_STACK - NRCU_T3_STACK];
};
__device__ __forceinline__ void push3(const DScene& s, Stack3& st, int& sp, float t, int r) {
    if (sp < NRCU_T3_STACK) st.smem[sp * NRCU_TRACE_THREADS] = make_uint2(__float_as_uint(t), (unsigned)r);
    else if (sp < s.stack_limit) { st.ot[sp - NRCU_T3_STACK] = t; st.oref[sp - NRCU_T3_STACK] = r; }
    else { atomicAdd(s.overflow, 1u); return; }   // cannot happen with trees from nrcu_bvh.cuh (depth cap); counted, never silent
    sp++;
}
__device__ __forceinline__ int pop3(Stack3& st, int& sp, float best_t) {
    while (sp > 0) {
        sp--;
        float t; int r;
        if (sp < NRCU_T3_STACK) { uint2 v = st.smem[sp * NRCU_TRACE_THREADS]; t = __uint_as_float(v.x); r = (int)v.y; }
        else { t = st.ot[sp - NRCU_T3_STACK]; r = st.oref[sp - NRCU_T3_STACK]; }
        if (t <= best_t) return r;   // ties must stay reachable
    }
    return NRCU_REF_DONE;
}
// One BVH4 node (same arithmetic as node_step in nrcu_intersect.cuh): conservative slab test of the four
// children, nearest hit child returned, the others pushed far-to-near with predicated shared-memory stores.
__device__ __forceinline__ int node_step3(const DScene& s, const RayPrep& rp, int cur, float best_t, Stack3& st, int& sp) {
    const f4* nd = s.nodes + (size_t)cur * NRCU_BVH_NODE_F4;
    f4 lox = ldg4(nd), hix = ldg4(nd + 1), loy = ldg4(nd + 2), hiy = ldg4(nd + 3), loz = ldg4(nd + 4), hiz = ldg4(nd + 5);
    i4 refs = ldg4i(nd + 6);
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
    NRCU_CSWAP(t0, r0, t1, r1); NRCU_CSWAP(t2, r2, t3, r3);
    NRCU_CSWAP(t0, r0, t2, r2); NRCU_CSWAP(t1, r1, t3, r3);
    NRCU_CSWAP(t1, r1, t2, r2);
    if (sp + 3 <= NRCU_T3_STACK) {   // fast path: no overflow checks
        if (t3 < NRCU_INF) { st.smem[sp * NRCU_TRACE_THREADS] = make_uint2(__float_as_uint(t3), (unsigned)r3); sp++; }
        if (t2 < NRCU_INF) { st.smem[sp * NRCU_TRACE_THREADS] = make_uint2(__float_as_uint(t2), (unsigned)r2); sp++; }
        if (t1 < NRCU_INF) { st.smem[sp * NRCU_TRACE_THREADS] = make_uint2(__float_as_uint(t1), (unsigned)r1); sp++; }
    } else {
        if (t3 < NRCU_INF) push3(s, st, sp, t3, r3);
        if (t2 < NRCU_INF) push3(s, st, sp, t2, r2);
        if (t1 < NRCU_INF) push3(s, st, sp, t1, r1);
    }
    return (t0 < NRCU_INF) ? r0 : pop3(st, sp, best_t);
}

// Work fetching shared by the stage-2 kernels: the idle lanes of a warp claim `n_idle` queue entries with ONE
// atomic (the first idle lane) and take them by ballot/popc rank.  `surv` (stage-1 survivor list) maps the queue
// position to the ray index; the provisional hit of stage 1 seeds (best_t, best_id).
struct Lane {
    Ray r; RayPrep rp; vec3 ginv;
    float best_t; int best_id; int cur; int sp; uint32_t idx;
};
template <bool GATE>
__device__ __forceinline__ bool fetch_ray(const DScene& s, const PathQueue& q, const uint32_t* surv, const float2* hits, uint32_t j, Lane& L) {
    const uint32_t i = surv ? surv[j] : j;
    f4 a = q.a[i]; float2 b = q.b[i];
    L.r.o = mk3(a.x, a.y, a.z); L.r.d = mk3(a.w, b.x, b.y);
    L.rp = prep_ray(L.r);
    if (GATE) L.ginv = gate_inverse(L.r, L.rp);
    L.best_t = NRCU_INF; L.best_id = -1;
    if (surv) { float2 h = hits[i]; L.best_t = h.x; L.best_id = __float_as_int(h.y); }
    L.sp = 0; L.idx = i;
    L.cur = s.root_ref == NRCU_REF_EMPTY ? NRCU_REF_DONE : s.root_ref;
    return true;
}

// v2: persistent threads, "while-while" traversal and lane refill.  Every lane keeps one ray's traversal state
// in registers; when at least `refill` lanes of the warp are idle (or all of them) the warp claims that many new
// rays.  Leaf work is postponed until every active lane has reached a leaf (Aila & Laine's while-while).
#ifndef NRCU_TRACE_MINB
#define NRCU_TRACE_MINB 1
#endif
template <bool GATE, bool IFIF>
__global__ void __launch_bounds__(NRCU_TRACE_THREADS, NRCU_TRACE_MINB) k_trace2(DScene s, PathQueue q, const uint32_t* n_ptr, const uint32_t* surv, float2* hits,
                                                               uint32_t* fetch, unsigned long long* ray_counter, uint32_t refill) {
    __shared__ uint2 stack_mem[NRCU_T3_STACK * NRCU_TRACE_THREADS];
    const uint32_t n = *n_ptr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;
    Stack3 st; st.smem = stack_mem + threadIdx.x;
    Lane L; L.r.o = mk3(0.f); L.r.d = mk3(0.f); L.rp.inv = mk3(0.f); L.rp.oinv = mk3(0.f); L.ginv = mk3(0.f);
    L.best_t = NRCU_INF; L.best_id = -1; L.cur = NRCU_REF_DONE; L.sp = 0; L.idx = 0;
    bool active = false, exhausted = false;
    for (;;) {
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle == 0xffffffffu && exhausted) break;
        const uint32_t n_idle = __popc(idle);
        if (!exhausted && n_idle >= (idle == 0xffffffffu ? 1u : refill)) {
            uint32_t base = 0;
            const uint32_t leader = __ffs(idle) - 1u;
            if (lane == leader) {
                base = atomicAdd(fetch, n_idle);
                if (!surv && base < n) atomicAdd(ray_counter, (unsigned long long)min(n_idle, n - base));
            }
            base = __shfl_sync(0xffffffffu, base, leader);
            if (base + n_idle >= n) exhausted = true;
            if (!active) {
                const uint32_t j = base + __popc(idle & lt);
                if (j < n) active = fetch_ray<GATE>(s, q, surv, hits, j, L);
            }
        }
        if (active) {
            if (IFIF) { if (L.cur >= 0) L.cur = node_step3(s, L.rp, L.cur, L.best_t, st, L.sp); }   // "if-if": one step of either kind per iteration
            else while (L.cur >= 0) L.cur = node_step3(s, L.rp, L.cur, L.best_t, st, L.sp);
            if (L.cur < 0 && L.cur != NRCU_REF_DONE) {
                leaf_step<GATE>(s, L.r, L.ginv, L.cur, L.best_t, L.best_id);
                L.cur = pop3(st, L.sp, L.best_t);
            }
            if (L.cur == NRCU_REF_DONE) {
                hits[L.idx] = make_float2(L.best_t, __int_as_float(L.best_id));
                active = false;
            }
        }
    }
}

// v3: persistent threads with VOTE-SCHEDULED phases.  A lane is a small state machine with up to two kinds of
// pending work:
//     cur >= 0   an inner BVH4 node to test                      ("node" work)
//     pn  >  0   pn primitives of a parked leaf left to test      ("prim" work, one primitive per step)
// and each warp iteration executes ONE phase, chosen by ballot/popc majority between the lanes that can do a
// node step and the lanes that can do a primitive test.  A lane that reaches a leaf parks it and keeps walking
// inner nodes speculatively (Aila & Laine's postponed leaf), so most lanes qualify for either phase.
template <bool GATE>
__global__ void __launch_bounds__(NRCU_TRACE_THREADS) k_trace3(DScene s, PathQueue q, const uint32_t* n_ptr, const uint32_t* surv, float2* hits,
                                                               uint32_t* fetch, unsigned long long* ray_counter, uint32_t refill,
                                                               uint32_t w_node, uint32_t w_prim) {
    __shared__ uint2 stack_mem[NRCU_T3_STACK * NRCU_TRACE_THREADS];
    const uint32_t n = *n_ptr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;
    Stack3 st; st.smem = stack_mem + threadIdx.x;
    Lane L; L.r.o = mk3(0.f); L.r.d = mk3(0.f); L.rp.inv = mk3(0.f); L.rp.oinv = mk3(0.f); L.ginv = mk3(0.f);
    L.best_t = NRCU_INF; L.best_id = -1; L.cur = NRCU_REF_DONE; L.sp = 0; L.idx = 0;
    uint32_t pl = 0, pn = 0, pk = 0;
    bool active = false, exhausted = false;
    for (;;) {
        // ---- refill idle lanes from the queue ------------------------------------------------------
        const uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (idle) {
            if (exhausted) { if (idle == 0xffffffffu) break; }
            else {
                const uint32_t n_idle = __popc(idle);
                if (idle == 0xffffffffu || n_idle >= refill) {
                    uint32_t base = 0;
                    const uint32_t leader = __ffs(idle) - 1u;
                    if (lane == leader) {
                        base = atomicAdd(fetch, n_idle);
                        if (!surv && base < n) atomicAdd(ray_counter, (unsigned long long)min(n_idle, n - base));
                    }
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (base + n_idle >= n) exhausted = true;
                    if (!active) {
                        const uint32_t j = base + __popc(idle & lt);
                        if (j < n) { active = fetch_ray<GATE>(s, q, surv, hits, j, L); pn = 0; }
                    }
                }
            }
        }
        // ---- park a reached leaf as pending primitive work, or retire the ray -------------------------
        if (active && pn == 0 && L.cur < 0) {
            if (L.cur != NRCU_REF_DONE) {
                const uint32_t code = (uint32_t)(~L.cur);
                pl = code >> 4; pn = (code & 15u) + 1u;
                pk = ldg_u32(s.leaf_prims + pl);
                L.cur = pop3(st, L.sp, L.best_t);
            } else {
                hits[L.idx] = make_float2(L.best_t, __int_as_float(L.best_id));
                active = false;
            }
        }
        // ---- vote: which phase does this warp iteration run? ------------------------------------------
        const bool want_node = active && L.cur >= 0;
        const bool want_prim = active && pn > 0;
        const uint32_t m_node = __ballot_sync(0xffffffffu, want_node), m_prim = __ballot_sync(0xffffffffu, want_prim);
        if (__popc(m_node) * w_node >= __popc(m_prim) * w_prim) {
            if (want_node) L.cur = node_step3(s, L.rp, L.cur, L.best_t, st, L.sp);
        } else {
            if (want_prim) {
                prim_step<GATE>(s, L.r, L.ginv, pl, pk, L.best_t, L.best_id);
                pl++; pn--;
                if (pn) pk = ldg_u32(s.leaf_prims + pl);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Metropolis light transport (bodies in nrcu_mlt.cuh): one Markov chain per thread
// ---------------------------------------------------------------------------------------------
// numbers of mutation `mut` of chain `chain`: [0] large-step decision, [1] acceptance, [2 + i] state i
struct MltNumbers {
    uint64_t seed; uint32_t chain, mut, stream; u32x4 blk; uint32_t have;
    __device__ __forceinline__ MltNumbers(uint64_t sd, uint32_t c, uint32_t m, uint32_t st) : seed(sd), chain(c), mut(m), stream(st), have(0xffffffffu) {}
    __device__ __forceinline__ float get(uint32_t k) {
        if ((k >> 2) != have) { have = k >> 2; blk = rng_block(seed, chain, mut, stream, have); }
        const uint32_t c = k & 3u;
        return u01(c == 0 ? blk.x : (c == 1 ? blk.y : (c == 2 ? blk.z : blk.w)));
    }
};
// film coordinate (pixel units) -> the (up to) four pixels whose two-pixel-wide footprint contains it; row 0 = top
__device__ __forceinline__ void mlt_splat(f4* film, const DScene& s, float fx, float fy, vec3 c, float wgt) {
    if (!(wgt > 0.f) || !(wgt < NRCU_INF)) return;
    const int j0 = (int)floorf(fx), i0 = (int)floorf(fy);
    for (int dj = 0; dj < 2; dj++) for (int di = 0; di < 2; di++) {
        const int j = j0 + dj, i = i0 + di;
        if (j < 0 || j >= (int)s.width || i < 0 || i >= (int)s.height) continue;
        f4* px = film + (size_t)((int)s.height - 1 - i) * s.width + j;
        atomicAdd(&px->x, c.x * wgt); atomicAdd(&px->y, c.y * wgt); atomicAdd(&px->z, c.z * wgt);
    }
}
// b = mean scalar contribution of independent samples (Metropolis.cpp:83-92, N_Init).  The scalar contribution of every
// sample is kept: the chains start from samples drawn in proportion to it (below), their numbers being reproducible from
// (seed, sample index).
__device__ __forceinline__ void mlt_init_numbers(uint64_t seed, uint32_t sample, uint32_t ns, float* u) {
    MltNumbers rn(seed, sample, 0u, NRCU_STREAM_MLT - 1u);
    for (uint32_t k = 0; k < ns; k++) u[k] = rn.get(k);
}
template <bool GATE>
__global__ void __launch_bounds__(128) k_mlt_b(DScene s, uint64_t seed, uint32_t n_samples, float* scalar, unsigned long long* ray_counter) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t rays = 0;
    if (i < n_samples) {
        float u[NRCU_MLT_STATES(NRCU_MLT_MAX_DEPTH)];
        mlt_init_numbers(seed, i, NRCU_MLT_STATES(s.depth), u);
        float fx, fy;
        scalar[i] = mlt_scalar(mlt_eval<GATE>(s, u, fx, fy, rays));
    }
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_down_sync(0xffffffffu, rays, o);
    if ((threadIdx.x & 31) == 0 && rays) atomicAdd(ray_counter, (unsigned long long)rays);
}
// Inclusive prefix sums (double) of the n scalar contributions, one block: cdf[i] = sum_{j<=i} scalar[j]; cdf[n-1] = n b.
__global__ void __launch_bounds__(1024) k_mlt_cdf(const float* scalar, uint32_t n, double* cdf) {
    __shared__ double part[1024];
    const uint32_t per = (n + 1023u) / 1024u, lo = threadIdx.x * per, hi = min(n, lo + per);
    double sum = 0.0;
    for (uint32_t i = lo; i < hi; i++) sum += (double)scalar[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) { double run = 0.0; for (int t = 0; t < 1024; t++) { double v = part[t]; part[t] = run; run += v; } }
    __syncthreads();
    double run = part[threadIdx.x];
    for (uint32_t i = lo; i < hi; i++) { run += (double)scalar[i]; cdf[i] = run; }
}
// The chains (Metropolis.cpp:25-69).  counters: [0] rays, [1] accepted mutations.  A chain starts from one of the b-estimation
// samples, drawn with probability proportional to its scalar contribution: the chain is then in its stationary
// distribution from the first mutation on, which the expected-value weights below assume - the reference starts ONE long
// chain anywhere and lets 2 M mutations forget the start; a GPU runs 10^5 short chains, where that start-up bias would
// darken the frame.
template <bool GATE>
__global__ void __launch_bounds__(128) k_mlt_chains(DScene s, uint64_t seed, uint32_t n_chains, uint32_t mutations, const double* cdf, uint32_t n_init,
                                                  float p_large, f4* film, unsigned long long* counters) {
    const uint32_t chain = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t rays = 0, accepted = 0;
    const double total = cdf[n_init - 1];
    const float b = (float)(total / (double)n_init);
    if (chain < n_chains && b > 0.f) {
        const uint32_t ns = NRCU_MLT_STATES(s.depth);
        float cur[NRCU_MLT_STATES(NRCU_MLT_MAX_DEPTH)], prop[NRCU_MLT_STATES(NRCU_MLT_MAX_DEPTH)];
        float cfx, cfy;
        {
            MltNumbers rn(seed, chain, 0u, NRCU_STREAM_MLT);
            const double target = ((double)rn.get(0) + (double)rn.get(1) * (1.0 / 16777216.0)) * total;   // 48 random bits
            uint32_t lo = 0, hi = n_init - 1;
            while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (target < cdf[mid]) hi = mid; else lo = mid + 1; }
            mlt_init_numbers(seed, lo, ns, cur);
        }
        vec3 Lc = mlt_eval<GATE>(s, cur, cfx, cfy, rays);
        float Ic = mlt_scalar(Lc);
        const float s1f = 2.0f / (float)(s.width + s.height);
        for (uint32_t m = 1; m <= mutations; m++) {
            MltNumbers rn(seed, chain, m, NRCU_STREAM_MLT);
            const bool large = rn.get(0) <= p_large;
            const float r_acc = rn.get(1);
            if (large) for (uint32_t k = 0; k < ns; k++) prop[k] = rn.get(2u + k);                 // large_step: a fresh path
            else {
                prop[0] = mlt_perturb(cur[0], s1f, 0.1f, rn.get(2)); prop[1] = mlt_perturb(cur[1], s1f, 0.1f, rn.get(3));   // the pixel location
                for (uint32_t k = 2; k < ns; k++) prop[k] = mlt_perturb(cur[k], 1.0f / 1024.0f, 1.0f / 64.0f, rn.get(2u + k));
            }
            float pfx, pfy;
            const vec3 Lp = mlt_eval<GATE>(s, prop, pfx, pfy, rays);
            const float Ip = mlt_scalar(Lp);
            // the reference leaves `a` uninitialised when the current path carries nothing; a dark state is always left
            const float a = Ic > 0.f ? fminf(1.f, Ip / Ic) : 1.f;
            if (Ip > 0.f) mlt_splat(film, s, pfx, pfy, Lp, (a + (large ? 1.f : 0.f)) / (Ip / b + p_large));
            if (Ic > 0.f) mlt_splat(film, s, cfx, cfy, Lc, (1.f - a) / (Ic / b + p_large));
            if (r_acc <= a) {
                for (uint32_t k = 0; k < ns; k++) cur[k] = prop[k];
                Lc = Lp; Ic = Ip; cfx = pfx; cfy = pfy; accepted++;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { rays += __shfl_down_sync(0xffffffffu, rays, o); accepted += __shfl_down_sync(0xffffffffu, accepted, o); }
    if ((threadIdx.x & 31) == 0) { if (rays) atomicAdd(counters, (unsigned long long)rays); if (accepted) atomicAdd(counters + 1, (unsigned long long)accepted); }
}
// film sums -> frame: x scale (film area / (4 mutations): every film sample lies in four pixel footprints), then the tone map:
// 0 sqrt (the path tracers' gamma, AccPathTracer.cpp:14-16), 1 the reference MLT's pow(1 - exp(-x), 1/2.2) (Metropolis.cpp:118-123), 2 none
__global__ void k_mlt_resolve(const f4* film, f4* rgba, uint32_t npix, float scale, int tone) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    f4 v = film[p];
    float c[3] = {v.x * scale, v.y * scale, v.z * scale};
    for (int k = 0; k < 3; k++) {
        if (tone == 0) c[k] = sqrtf(c[k]);
        else if (tone == 1) c[k] = powf(1.f - expf(-c[k]), 1.f / 2.2f);
    }
    rgba[p] = mk4(c[0], c[1], c[2], 1.f);
}

// Brute-force variant for the RayCast-mode parity probe (nrcu_trace_batch in NRCU_MODE_RAYCAST).
__global__ void k_trace_linear_rc(DScene s, PathQueue q, uint32_t n, float2* hits) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f4 a = q.a[i]; float2 b = q.b[i];
    Ray r; r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
    float t; int id;
    closest_hit_linear<true>(s, r, t, id);
    hits[i] = make_float2(t, __int_as_float(id));
}

// Shading + next-ray generation for bounce `d`; surviving paths are compacted into `qo` with one atomic per
// warp (ballot + popc prefix).  The queue entry of the NEXT iteration is requested right after this iteration's
// output-slot atomic has been issued, so the atomic's round trip and the loads' latency overlap (ncu on the
// unpipelined loop: 38 % of the stall samples sat on the shuffle that waits for the atomic, 16 % on the first use of
// the loaded entry).  Fusing stage 1 of the closest hit into this kernel - after the shading (next ray still in
// registers) or before it (SHADE_STAGE1: rays that end at stage 1 shaded at once, the rest deferred) - was built and
// measured three times and lost every time (profiles/r1_history.md): 64 -> 73 registers or spills, and two
// divergent phases back to back in the same warps.
#ifndef NRCU_SHADE_MINB
#define NRCU_SHADE_MINB 4
#endif
// NRCU_OPT_BRANCH_TEMPLATE=1 makes BRANCH (the reference's two-branch glass recursion) a template parameter, so that the
// stochastic default carries none of the second-branch bookkeeping (split ballots, branch-bit loads and stores, shared-slot
// float atomics): 280 fewer SASS instructions - and 1.5 % SLOWER in the same-call A/B (3015-3027 vs 3067 Mpath-samples/s,
// profiles/r2_history.md), like every other "leaner" form of this latency-bound kernel.  Off: the mode is a kernel argument.
#ifndef NRCU_OPT_BRANCH_TEMPLATE
#define NRCU_OPT_BRANCH_TEMPLATE 0
#endif
template <bool NEE, bool BRANCH_T>
__global__ void __launch_bounds__(256, NRCU_SHADE_MINB) k_shade(DScene s, uint64_t seed, uint32_t d, int glass_branch_rt, uint32_t sample0,
                                              PathQueue qi, QRegions rin, const float2* hits,
                                              PathQueue qo, uint32_t* n_out_ptr, uint32_t out_logk, uint32_t out_capacity, f4* L,
                                              PathQueue qs, uint32_t* n_shadow_ptr) {
#if NRCU_OPT_BRANCH_TEMPLATE
    constexpr bool BRANCH = BRANCH_T;
#else
    const bool BRANCH = glass_branch_rt != 0;   // A/B only: the round-1 form with the mode as a kernel argument
#endif
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t npix = s.width * s.height;
    uint32_t my_cnt;
    const uint32_t nblocks = regions_begin(rin, lane, my_cnt);
    f4 a = mk4(0, 0, 0, 0), c = a; float2 b = make_float2(0.f, 0.f), h = b; uint32_t br = 0;
    // Whole blocks are read (lanes past their region's count hold entries nobody uses), and the block index is clamped
    // instead of tested: with a branch around them ptxas sank these loads below the shuffle that waits for the slot
    // atomic, so the atomic's round trip and the loads' latency added up in every iteration (ncu source page: the two
    // hottest stall sites of the kernel); without control flow they issue right behind the atomic.
    auto load_entry = [&](uint32_t blk) {
        const uint32_t j = min(blk, nblocks - 1u) * 32u + lane;
        a = qi.a[j]; b = qi.b[j]; c = qi.c[j]; h = hits[j];
        if (BRANCH) br = qi.d[j];
    };
    if (nblocks == 0) return;
    load_entry(warp_global);
    for (uint32_t pb = warp_global; pb < nblocks; pb += warps_total) {
        const bool live = region_live(rin, my_cnt, pb, lane);
        int n_out = 0;
        PathStep ps;
        uint32_t slot = 0, branch = 0;
        if (live) {
            Ray r; r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
            vec3 thr = mk3(c.x, c.y, c.z);
            slot = (uint32_t)f2i(c.w) & 0x7fffffffu; branch = br;
            const bool skip_light = ((uint32_t)f2i(c.w) >> 31) != 0u;   // the previous vertex sent a shadow ray (NEE)
            uint32_t pixel = slot % npix, sample = sample0 + slot / npix;
            ps = path_vertex<NEE>(s, seed, pixel, sample, d, branch, r, thr, h.x, __float_as_int(h.y), BRANCH ? 1 : 0, skip_light);
            if (ps.action == PATH_TERMINATE) {
                if (NEE) {            // shadow rays of earlier bounces add to the same slot (k_shadow_resolve)
                    if (ps.radiance.x != 0.f || ps.radiance.y != 0.f || ps.radiance.z != 0.f) {
                        if (BRANCH) { atomicAdd(&L[slot].x, ps.radiance.x); atomicAdd(&L[slot].y, ps.radiance.y); atomicAdd(&L[slot].z, ps.radiance.z); }
                        else { f4 v = L[slot]; L[slot] = mk4(v.x + ps.radiance.x, v.y + ps.radiance.y, v.z + ps.radiance.z, 0.f); }
                    }
                } else if (BRANCH) {   // several branches of one path share the slot
                    if (ps.radiance.x != 0.f) atomicAdd(&L[slot].x, ps.radiance.x);
                    if (ps.radiance.y != 0.f) atomicAdd(&L[slot].y, ps.radiance.y);
                    if (ps.radiance.z != 0.f) atomicAdd(&L[slot].z, ps.radiance.z);
                } else if (ps.radiance.x != 0.f || ps.radiance.y != 0.f || ps.radiance.z != 0.f) {
                    L[slot] = mk4(ps.radiance.x, ps.radiance.y, ps.radiance.z, 0.f);   // one path per slot, it ends once: plain store into the zeroed slot
                }
            } else n_out = ps.action == PATH_SPLIT ? 2 : 1;
        }
        // warp-aggregated allocation in the output queue
        const uint32_t m1 = __ballot_sync(0xffffffffu, n_out >= 1), m2 = BRANCH ? __ballot_sync(0xffffffffu, n_out == 2) : 0u;
        const uint32_t total = __popc(m1) + __popc(m2);
        uint32_t start = 0;
        const uint32_t r_out = pb & ((1u << out_logk) - 1u);   // this block's survivors go to region pb mod K of the output queue
        if (lane == 0 && total) start = atomicAdd(n_out_ptr + (size_t)r_out * rin.cs, total);
        load_entry(pb + warps_total);   // prefetch the next iteration's entry while the atomic is in flight
        const uint32_t lt = (1u << lane) - 1u;
        if (NEE) {   // shadow rays of this bounce, compacted into their own queue
            const bool sh = live && ps.nee;
            const uint32_t ms = __ballot_sync(0xffffffffu, sh);
            if (ms) {
                uint32_t s0 = 0;
                if (lane == 0) s0 = atomicAdd(n_shadow_ptr, (uint32_t)__popc(ms));
                s0 = __shfl_sync(0xffffffffu, s0, 0);
                const uint32_t sp = s0 + __popc(ms & lt);
                if (sh && sp < out_capacity) {
                    qs.a[sp] = mk4(ps.shadow.o.x, ps.shadow.o.y, ps.shadow.o.z, ps.shadow.d.x);
                    qs.b[sp] = make_float2(ps.shadow.d.y, ps.shadow.d.z);
                    qs.c[sp] = mk4(ps.nee_contrib.x, ps.nee_contrib.y, ps.nee_contrib.z, i2f((int)slot));
                    qs.d[sp] = (uint32_t)ps.nee_light;
                }
            }
        }
        if (total == 0) continue;   // warp-uniform
        start = __shfl_sync(0xffffffffu, start, 0);
        // entry p of region r sits at ((p / 32) K + r) 32 + p mod 32 (K = 1: at p)
        const uint32_t p1 = start + __popc(m1 & lt), p2 = start + __popc(m1) + __popc(m2 & lt);
        const uint32_t pos1 = ((((p1 >> 5) << out_logk) + r_out) << 5) | (p1 & 31u), pos2 = ((((p2 >> 5) << out_logk) + r_out) << 5) | (p2 & 31u);
        if (n_out >= 1 && pos1 < out_capacity) {
            qo.a[pos1] = mk4(ps.next.o.x, ps.next.o.y, ps.next.o.z, ps.next.d.x);
            qo.b[pos1] = make_float2(ps.next.d.y, ps.next.d.z);
            qo.c[pos1] = mk4(ps.thr.x, ps.thr.y, ps.thr.z, i2f((int)(slot | ((NEE && ps.next_skips_light) ? 0x80000000u : 0u))));
            if (BRANCH) qo.d[pos1] = branch;
        } else if (n_out >= 1 && !BRANCH) atomicAdd(s.overflow + 1, 1u);   // without branching the queue cannot outgrow its array: counted, never silent (branching: k_clamp_count)
        if (BRANCH && n_out == 2 && pos2 < out_capacity) {   // glass branch mode only
            qo.a[pos2] = mk4(ps.next2.o.x, ps.next2.o.y, ps.next2.o.z, ps.next2.d.x);
            qo.b[pos2] = make_float2(ps.next2.d.y, ps.next2.d.z);
            qo.c[pos2] = mk4(ps.thr2.x, ps.thr2.y, ps.thr2.z, i2f((int)slot));
            qo.d[pos2] = branch | (1u << (d & 31u));
        }
    }
}

// k_shade with a per-warp POOL of surface hits (the default estimator: no NEE, no branching glass).
// In k_shade a lane whose ray left the scene or reached the light has nothing to do during the ~700 instructions of the
// surface shading (hit record and material gathers, Philox, hemisphere sample, Onb, two normalisations, six IEEE
// divisions): 22-24 of 32 lanes per instruction at the bounces after the first (ncu), and the whole frame runs at the issue
// rate the stage-1 kernel has alone, so idle lanes are lost throughput.  Here a warp works in two phases:
//   A  (every entry of a block)  closest light, "is the object hit in front of it?"; a path that ends here stores its
//      radiance at once; a surface hit is appended - ballot/popc rank, 12 words - to the warp's ring in shared memory;
//   B  (whenever the ring holds 32 hits, and once at the end for the rest)  path_vertex_hit on 32 DENSE lanes, then the
//      usual warp-aggregated append to the next queue.
// Same arithmetic per path, same result per slot; only the order in which a warp meets its paths changes.
// Output region of a round: the warp's round counter, rotated by the warp index - every warp spreads its rounds evenly
// over the K regions, so the regions stay equally long up to one round per warp (the slack nrcu_api.cu allocates).
#define NRCU_POOL_RING 64
#ifndef NRCU_POOL_EARLY_ATOMIC
#define NRCU_POOL_EARLY_ATOMIC 1
#endif
// Where the next block's queue entry is requested: 0 before phase A, 1 before the shading round, 2 at the end of the
// iteration (default), 3 = 0 through cp.async into shared memory.  Same-call A/B on cfg3 (profiles/r2_history.md):
// 0 3144, 1 3181, 2 3305, 3 3297 Mpath-samples/s - the twelve registers of an entry in flight across the shading round cost
// 60-70 bytes of spills per thread, which matters more than the exposed load latency (other warps cover it).
#ifndef NRCU_POOL_PREFETCH_LATE
#define NRCU_POOL_PREFETCH_LATE 2
#endif
__global__ void __launch_bounds__(256, NRCU_SHADE_MINB) k_shade_pool(DScene s, uint64_t seed, uint32_t d, uint32_t sample0,
                                                                       PathQueue qi, QRegions rin, const float2* hits,
                                                                       PathQueue qo, uint32_t* n_out_ptr, uint32_t out_logk, uint32_t out_capacity, f4* L) {
    __shared__ float ring[8][12][NRCU_POOL_RING];   // [warp][word][entry]: o.xyz d.xyz thr.xyz slot t id
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t npix = s.width * s.height;
    uint32_t my_cnt;
    const uint32_t nblocks = regions_begin(rin, lane, my_cnt);
    if (nblocks == 0) return;
    float (*rg)[NRCU_POOL_RING] = ring[wib];
    uint32_t head = 0, fill = 0, round = warp_global;   // ring state and round counter: warp-uniform
    f4 a, c; float2 b, h;
#if NRCU_POOL_PREFETCH_LATE == 3
    // the next block travels global -> shared with cp.async (no registers held across the shading round)
    __shared__ f4 st_a[8][32], st_c[8][32];
    __shared__ float2 st_b[8][32], st_h[8][32];
    auto cp16 = [](void* dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory"); };
    auto cp8 = [](void* dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory"); };
    auto load_entry = [&](uint32_t blk) {
        const uint32_t j = min(blk, nblocks - 1u) * 32u + lane;
        cp16(&st_a[wib][lane], qi.a + j); cp8(&st_b[wib][lane], qi.b + j); cp16(&st_c[wib][lane], qi.c + j); cp8(&st_h[wib][lane], hits + j);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto take_entry = [&]() {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        a = st_a[wib][lane]; b = st_b[wib][lane]; c = st_c[wib][lane]; h = st_h[wib][lane];
    };
#else
    auto load_entry = [&](uint32_t blk) {   // clamped, no branch: see k_shade
        const uint32_t j = min(blk, nblocks - 1u) * 32u + lane;
        a = qi.a[j]; b = qi.b[j]; c = qi.c[j]; h = hits[j];
    };
    auto take_entry = [&]() {};
#endif
    // phase B on the `take` oldest ring entries
    const bool last_bounce = d + 1 == s.depth;
    auto shade_round = [&](uint32_t take) {
        int n_out = 0;
        PathStep ps;
        uint32_t slot = 0;
        const bool act = lane < take;
        Ray r; vec3 thr = mk3(0.f); HitSetup hs; hs.type = 0u;
        if (act) {
            const uint32_t e = (head + lane) & (NRCU_POOL_RING - 1u);
            r.o = mk3(rg[0][e], rg[1][e], rg[2][e]); r.d = mk3(rg[3][e], rg[4][e], rg[5][e]);
            thr = mk3(rg[6][e], rg[7][e], rg[8][e]);
            slot = (uint32_t)f2i(rg[9][e]);
            hs = hit_setup(s, r, rg[10][e], f2i(rg[11][e]));
        }
        // When every vertex of the round sits on a material that always continues the path (Lambertian, conductor), the
        // number of output entries is known before the shading: the slot atomic is issued here and its round trip hides
        // behind the ~600 instructions of the shading instead of being waited for right after them (ncu: 13 % of the
        // kernel's stall samples sat on that shuffle).
        const uint32_t r_out = round & ((1u << out_logk) - 1u);
        round++;
        const bool early = NRCU_POOL_EARLY_ATOMIC && __all_sync(0xffffffffu, !act || type_always_continues(hs.type));
        uint32_t start = 0, total = last_bounce ? 0u : take;
        if (early && lane == 0 && total) start = atomicAdd(n_out_ptr + (size_t)r_out * rin.cs, total);
        if (act) {
            path_step_init(ps, r, thr);
            path_vertex_shade<false>(ps, s, seed, slot % npix, sample0 + slot / npix, d, 0u, r, thr, hs, 0);
            if (ps.action == PATH_TERMINATE) {
                if (ps.radiance.x != 0.f || ps.radiance.y != 0.f || ps.radiance.z != 0.f) L[slot] = mk4(ps.radiance.x, ps.radiance.y, ps.radiance.z, 0.f);
            } else n_out = 1;
        }
        const uint32_t m1 = __ballot_sync(0xffffffffu, n_out != 0);
        if (!early) {
            total = __popc(m1);
            if (lane == 0 && total) start = atomicAdd(n_out_ptr + (size_t)r_out * rin.cs, total);
        }
        if (total == 0) return;   // warp-uniform
        start = __shfl_sync(0xffffffffu, start, 0);
        const uint32_t p1 = start + __popc(m1 & lt);
        const uint32_t pos1 = ((((p1 >> 5) << out_logk) + r_out) << 5) | (p1 & 31u);
        if (n_out && pos1 < out_capacity) {
            qo.a[pos1] = mk4(ps.next.o.x, ps.next.o.y, ps.next.o.z, ps.next.d.x);
            qo.b[pos1] = make_float2(ps.next.d.y, ps.next.d.z);
            qo.c[pos1] = mk4(ps.thr.x, ps.thr.y, ps.thr.z, i2f((int)slot));
        } else if (n_out) atomicAdd(s.overflow + 1, 1u);   // cannot happen (the regions' slack bounds the extent); counted, never silent
    };
    load_entry(warp_global);
    for (uint32_t pb = warp_global; pb < nblocks; pb += warps_total) {
        // ---- phase A: where does the path go? ----------------------------------------------------------------------
        const bool live = region_live(rin, my_cnt, pb, lane);
        bool surface = false;
        take_entry();
        Ray r; r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
        const vec3 thr = mk3(c.x, c.y, c.z);
        const float slot_f = c.w, t = h.x, id_f = h.y;
#if !NRCU_POOL_PREFETCH_LATE || NRCU_POOL_PREFETCH_LATE == 3
        load_entry(pb + warps_total);   // the next block: in flight during this block's phase A and the shading round
#endif
        if (live) {
            vec3 radiance;
            const float tl = closest_light(s, r, radiance);
            const int id = __float_as_int(id_f);
            if (id >= 0 && t < tl) surface = true;
            else {   // path_vertex's other branches: the light, or nothing (environment map / black)
                vec3 out = mk3(0.f);
                if (tl != NRCU_INF) out = thr * radiance;
                else if (s.env_rgba && s.mode == MODE_ACC) out = thr * env_lookup(s, r.d);
                if (out.x != 0.f || out.y != 0.f || out.z != 0.f) L[(uint32_t)f2i(slot_f) & 0x7fffffffu] = mk4(out.x, out.y, out.z, 0.f);
            }
        }
        const uint32_t ms = __ballot_sync(0xffffffffu, surface);
        if (surface) {
            const uint32_t e = (head + fill + __popc(ms & lt)) & (NRCU_POOL_RING - 1u);
            rg[0][e] = r.o.x; rg[1][e] = r.o.y; rg[2][e] = r.o.z; rg[3][e] = r.d.x; rg[4][e] = r.d.y; rg[5][e] = r.d.z;
            rg[6][e] = thr.x; rg[7][e] = thr.y; rg[8][e] = thr.z; rg[9][e] = i2f((int)((uint32_t)f2i(slot_f) & 0x7fffffffu)); rg[10][e] = t; rg[11][e] = id_f;
        }
        fill += __popc(ms);
        __syncwarp();
#if NRCU_POOL_PREFETCH_LATE == 1
        load_entry(pb + warps_total);   // requested before the shading round, not live during phase A
#endif
        // ---- phase B: a dense warp of surface hits --------------------------------------------------------------------
        if (fill >= 32u) {
            shade_round(32u);
            head = (head + 32u) & (NRCU_POOL_RING - 1u); fill -= 32u;
            __syncwarp();
        }
#if NRCU_POOL_PREFETCH_LATE == 2
        load_entry(pb + warps_total);
#endif
    }
    if (fill) shade_round(fill);
}

// ---------------------------------------------------------------------------------------------
// Path-regeneration scheduler (the default for the reference's estimator; NEE and the branching glass mode keep the
// per-bounce wavefront above)
// ---------------------------------------------------------------------------------------------
// A SLOT is bound to (pixel, lane): slot = lane * n_pixels + pixel, lane < K.  It carries ONE path at a time, in place
// (ray a/b, throughput + state c, hit), and renders the samples  sample0 + lane + j*K,  j = 0, 1, ...  of its pixel one
// after the other: when a path ends, the same thread adds its radiance to the slot's accumulator and starts the lane's
// next sample in the same slot.  Nothing is compacted, so there is no output-queue allocation (the atomicAdd the
// wavefront's k_shade spends 54 % of its stall samples on), every launch is full until the last samples of the frame
// drain, and one iteration = stage 1 + stage 2 + shade over all slots whatever bounce each path is at.
//   state (c.w bits) = j << 12 | bounce, or NRCU_SLOT_DEAD once the lane has no samples left (a.x = NaN then, which is
//   what stage 1 looks at).  lacc[slot] = (sum of the lane's radiances in sample order, rays traced for them).
// Per-pixel result = sum over lanes, in lane order, of the lane sums: a fixed fp32 summation tree for a given K.
#define NRCU_SLOT_DEAD 0xffffffffu
#define NRCU_SLOT_BOUNCE_BITS 12

// First iteration: the lanes' first samples (+ fused stage 1: these camera rays are still coherent).
template <bool GATE>
__global__ void __launch_bounds__(256) k_regen_init(DScene s, uint64_t seed, uint32_t sample0, uint32_t n_samples, uint32_t lane0, uint32_t n_slots,
                                                   PathQueue q, f4* lacc, float2* hits, uint32_t* surv, uint32_t* n_surv) {
    __shared__ BigList bl;
    bl.load(s);
    const uint32_t npix = s.width * s.height;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n_slots; base += stride) {   // warp-uniform trip count
        const uint32_t slot = base + threadIdx.x;
        bool more = false;
        if (slot < n_slots) {
            const uint32_t pixel = slot % npix, lane = lane0 + slot / npix;
            lacc[slot] = mk4(0.f, 0.f, 0.f, 0.f);
            if (lane < n_samples) {
                Ray r = pt_camera_ray(s, seed, pixel, sample0 + lane);
                q.a[slot] = mk4(r.o.x, r.o.y, r.o.z, r.d.x);
                q.b[slot] = make_float2(r.d.y, r.d.z);
                q.c[slot] = mk4(1.f, 1.f, 1.f, i2f(0));
                if (r.o.x == r.o.x) more = stage1<GATE>(s, bl, r, slot, hits);
                else hits[slot] = make_float2(NRCU_INF, __int_as_float(-1));   // a NaN origin hits nothing (and stage 1 of later iterations skips it)
            } else {
                q.a[slot] = mk4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
                q.c[slot] = mk4(0.f, 0.f, 0.f, __int_as_float((int)NRCU_SLOT_DEAD));
            }
        }
        append_survivors(more, slot, surv, n_surv);
    }
}

// One path vertex per live slot: shade, then either continue in place or finish the sample and start the lane's next one.
// Grid: x over pixels (256 per CTA), y over the lanes of this partition.  No atomics.
__global__ void __launch_bounds__(256, NRCU_SHADE_MINB) k_shade_regen(DScene s, uint64_t seed, uint32_t sample0, uint32_t n_samples, uint32_t K, uint32_t lane0,
                                                                      PathQueue q, float2* hits, f4* lacc, const uint32_t* flags_prev, uint32_t* flags_out) {
    if (!regen_anyone_alive(flags_prev)) return;
    const uint32_t npix = s.width * s.height;
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive_after = false;
    if (pixel < npix) {
        const size_t i = (size_t)blockIdx.y * npix + pixel;
        const f4 c = q.c[i];
        const uint32_t state = (uint32_t)f2i(c.w);
        if (state != NRCU_SLOT_DEAD) {
            const f4 a = q.a[i]; const float2 b = q.b[i]; float2 h = hits[i];
            if (!(a.x == a.x)) h = make_float2(NRCU_INF, __int_as_float(-1));   // stage 1 skipped it: its hit record is stale
            const uint32_t lane = lane0 + blockIdx.y, bounce = state & ((1u << NRCU_SLOT_BOUNCE_BITS) - 1u);
            uint32_t j = state >> NRCU_SLOT_BOUNCE_BITS;
            Ray r; r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
            const PathStep ps = path_vertex<false>(s, seed, pixel, sample0 + lane + j * K, bounce, 0u, r, mk3(c.x, c.y, c.z), h.x, __float_as_int(h.y), 0, false);
            if (ps.action == PATH_CONTINUE) {
                q.a[i] = mk4(ps.next.o.x, ps.next.o.y, ps.next.o.z, ps.next.d.x);
                q.b[i] = make_float2(ps.next.d.y, ps.next.d.z);
                q.c[i] = mk4(ps.thr.x, ps.thr.y, ps.thr.z, i2f((int)(state + 1u)));
                alive_after = true;
            } else {
                f4 v = lacc[i];
                lacc[i] = mk4(v.x + ps.radiance.x, v.y + ps.radiance.y, v.z + ps.radiance.z, v.w + (float)(bounce + 1u));   // w: rays of the finished path
                j++;
                const uint32_t next_lane_sample = lane + j * K;
                if (next_lane_sample < n_samples) {
                    Ray nr = pt_camera_ray(s, seed, pixel, sample0 + next_lane_sample);
                    q.a[i] = mk4(nr.o.x, nr.o.y, nr.o.z, nr.d.x);
                    q.b[i] = make_float2(nr.d.y, nr.d.z);
                    q.c[i] = mk4(1.f, 1.f, 1.f, i2f((int)(j << NRCU_SLOT_BOUNCE_BITS)));
                    alive_after = true;
                } else {
                    q.a[i] = mk4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
                    q.c[i] = mk4(0.f, 0.f, 0.f, __int_as_float((int)NRCU_SLOT_DEAD));
                }
            }
        }
    }
    if (__syncthreads_or((int)alive_after) && threadIdx.x == 0)
        flags_out[((blockIdx.x + blockIdx.y) % NRCU_REGEN_FLAGS) * NRCU_REGEN_FLAG_STRIDE] = 1u;
}

// Fused iteration of the regeneration scheduler (NRCU_REGEN_FUSED=1): shade the slot's vertex with the hit the previous
// iteration found, continue or regenerate in place, and run stage 1 of the closest hit on the NEW ray while it is still
// in registers - one kernel per iteration (+ the stage-2 traversal of the survivors).  The point is not the saved ray
// round trip but the mix: the shading phase of a warp is latency bound (dependent gathers), the stage-1 phase issue bound
// (~1 000 ALU instructions per 32 rays); in one kernel the slot data of the NEXT loop iteration is requested before stage 1
// starts, so its latency is covered by a thousand instructions of the same warp, and warps in different phases share an SM.
#ifndef NRCU_FUSED_MINB
#define NRCU_FUSED_MINB 3
#endif
template <bool GATE>
__global__ void __launch_bounds__(32 * NRCU_BIGB_WARPS, NRCU_FUSED_MINB) k_regen_fused(DScene s, uint64_t seed, uint32_t sample0, uint32_t n_samples, uint32_t K, uint32_t lane0, uint32_t n_slots,
                                                                                      PathQueue q, float2* hits, f4* lacc, uint32_t* surv, uint32_t* n_surv,
                                                                                      const uint32_t* flags_prev, uint32_t* flags_out) {
    __shared__ BigList bl;
    __shared__ unsigned short pairs[NRCU_BIGB_WARPS][32 * NRCU_MAX_BIG];
    __shared__ float rays[NRCU_BIGB_WARPS][6][32];
    __shared__ unsigned long long best[NRCU_BIGB_WARPS][32];
    if (!regen_anyone_alive(flags_prev)) return;
    bl.load(s);
    const uint32_t n = n_slots, npix = s.width * s.height;
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned short* my_pairs = pairs[wib];
    f4 a = mk4(0, 0, 0, 0), c = mk4(0, 0, 0, __int_as_float((int)NRCU_SLOT_DEAD)); float2 b = make_float2(0.f, 0.f), h = b;
    auto load_slot = [&](uint32_t j) {
        c = mk4(0, 0, 0, __int_as_float((int)NRCU_SLOT_DEAD));
        if (j >= n) return;
        c = q.c[j]; a = q.a[j]; b = q.b[j]; h = hits[j];
    };
    load_slot(warp_global * 32u + lane);
    bool any_alive = false;
    for (uint32_t base = warp_global * 32u; base < n; base += warps_total * 32u) {
        const uint32_t i = base + lane;
        const uint32_t state = (uint32_t)f2i(c.w);
        bool live = state != NRCU_SLOT_DEAD;
        if (!__any_sync(0xffffffffu, live)) { load_slot(i + warps_total * 32u); continue; }
        // ---- shading phase: one path vertex per live slot ----------------------------------------------------------
        Ray r; r.o = mk3(0.f); r.d = mk3(0.f);
        if (live) {
            if (!(a.x == a.x)) h = make_float2(NRCU_INF, __int_as_float(-1));
            const uint32_t pixel = i % npix, lane_id = lane0 + i / npix, bounce = state & ((1u << NRCU_SLOT_BOUNCE_BITS) - 1u);
            uint32_t j = state >> NRCU_SLOT_BOUNCE_BITS;
            Ray ray; ray.o = mk3(a.x, a.y, a.z); ray.d = mk3(a.w, b.x, b.y);
            const PathStep ps = path_vertex<false>(s, seed, pixel, sample0 + lane_id + j * K, bounce, 0u, ray, mk3(c.x, c.y, c.z), h.x, __float_as_int(h.y), 0, false);
            if (ps.action == PATH_CONTINUE) {
                r = ps.next;
                q.c[i] = mk4(ps.thr.x, ps.thr.y, ps.thr.z, i2f((int)(state + 1u)));
            } else {
                f4 v = lacc[i];
                lacc[i] = mk4(v.x + ps.radiance.x, v.y + ps.radiance.y, v.z + ps.radiance.z, v.w + (float)(bounce + 1u));
                j++;
                const uint32_t next_lane_sample = lane_id + j * K;
                if (next_lane_sample < n_samples) {
                    r = pt_camera_ray(s, seed, pixel, sample0 + next_lane_sample);
                    q.c[i] = mk4(1.f, 1.f, 1.f, i2f((int)(j << NRCU_SLOT_BOUNCE_BITS)));
                } else {
                    live = false;
                    q.a[i] = mk4(__int_as_float(0x7fc00000), 0.f, 0.f, 0.f);
                    q.c[i] = mk4(0.f, 0.f, 0.f, __int_as_float((int)NRCU_SLOT_DEAD));
                }
            }
            if (live) {
                q.a[i] = mk4(r.o.x, r.o.y, r.o.z, r.d.x);
                q.b[i] = make_float2(r.d.y, r.d.z);
                live = r.o.x == r.o.x;      // a NaN origin hits nothing: no stage 1, the next iteration shades it as a miss
                if (!live) hits[i] = make_float2(NRCU_INF, __int_as_float(-1));
                any_alive = true;
            }
        }
        // ---- the next loop iteration's slot data: in flight during stage 1 -------------------------------------------
        load_slot(i + warps_total * 32u);
        // ---- stage 1 of the closest hit on the new rays (k_big_balanced's body) ---------------------------------------
        const RayPrep rp = prep_ray(r);
        rays[wib][0][lane] = r.o.x; rays[wib][1][lane] = r.o.y; rays[wib][2][lane] = r.o.z;
        rays[wib][3][lane] = r.d.x; rays[wib][4][lane] = r.d.y; rays[wib][5][lane] = r.d.z;
        best[wib][lane] = NRCU_BEST_NONE;
        uint32_t total = 0;
        const float nox = live ? -rp.oinv.x : -NRCU_INF;
        const vec3 ainv = mk3(fabsf(rp.inv.x), fabsf(rp.inv.y), fabsf(rp.inv.z));
        for (uint32_t k = 0; k < s.n_big; k++) {
            float tn, tf;
            slab_center_extent(bl.bd[2 * k], bl.bd[2 * k + 1], rp, ainv, nox, tn, tf);
            const bool cand = tn <= tf;
            const uint32_t m = __ballot_sync(0xffffffffu, cand);
            if (cand) my_pairs[total + __popc(m & lt)] = (unsigned short)(lane | (k << 5));
            total += __popc(m);
        }
        __syncwarp();
        for (uint32_t jb = 0; jb < total; jb += 32u) {
            const uint32_t jj = jb + lane;
            if (jj < total) {
                const uint32_t p = my_pairs[jj], ol = p & 31u, k = p >> 5;
                Ray pr; pr.o = mk3(rays[wib][0][ol], rays[wib][1][ol], rays[wib][2][ol]); pr.d = mk3(rays[wib][3][ol], rays[wib][4][ol], rays[wib][5][ol]);
                float bt = __uint_as_float((uint32_t)(best[wib][ol] >> 32));
                int bi = 0x7fffffff;
                prim_test<false>(pr, mk3(0.f), bl.g[3 * k], bl.g[3 * k + 1], bl.g[3 * k + 2], bl.b, bl.m[k], bt, bi);
                if (bi != 0x7fffffff) atomicMin(&best[wib][ol], ((unsigned long long)__float_as_uint(bt) << 32) | (unsigned long long)(((uint32_t)bi << 5) | k));
            }
        }
        __syncwarp();
        bool more = false;
        if (live) {
            const unsigned long long key = best[wib][lane];
            float best_t = NRCU_INF; int best_id = -1;
            if (key != NRCU_BEST_NONE) {
                best_t = __uint_as_float((uint32_t)(key >> 32)); best_id = (int)((uint32_t)key >> 5);
                if (GATE) {
                    const uint32_t kb = (uint32_t)key & 31u;
                    const vec3 ginv = gate_inverse(r, rp);
                    if (!bounds_intersectp_inv(bl.b[2 * kb], bl.b[2 * kb + 1], r, ginv.x, ginv.y, ginv.z)) {
                        best_t = NRCU_INF; best_id = -1;
                        for (uint32_t k = 0; k < s.n_big; k++)
                            prim_test<true>(r, ginv, bl.g[3 * k], bl.g[3 * k + 1], bl.g[3 * k + 2], bl.b + 2 * k, bl.m[k], best_t, best_id);
                    }
                }
            }
            hits[i] = make_float2(best_t, __int_as_float(best_id));
            more = bvh_reachable(s, rp, best_t);
        }
        __syncwarp();
        append_survivors(more, i, surv, n_surv);
    }
    if (__any_sync(0xffffffffu, any_alive) && lane == 0) flags_out[(warp_global % NRCU_REGEN_FLAGS) * NRCU_REGEN_FLAG_STRIDE] = 1u;
}

// End of the frame (or slice): accum[p].rgb += the lanes' sums in lane order, accum[p].a += n_samples (passed for the
// first partition only); the rays the lanes counted go to the context's ray counter, one atomic per CTA.
__global__ void __launch_bounds__(256) k_accumulate_lanes(const f4* lacc, f4* accum, uint32_t npix, uint32_t lanes, uint32_t n_samples, unsigned long long* ray_counter) {
    __shared__ unsigned long long warp_rays[8];
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long rays = 0;
    if (p < npix) {
        f4 acc = accum[p];
        for (uint32_t k = 0; k < lanes; k++) {
            f4 v = lacc[(size_t)k * npix + p];
            acc.x += v.x; acc.y += v.y; acc.z += v.z;
            rays += (unsigned long long)v.w;
        }
        acc.w += (float)n_samples;
        accum[p] = acc;
    }
    for (int o = 16; o > 0; o >>= 1) rays += __shfl_down_sync(0xffffffffu, rays, o);
    if ((threadIdx.x & 31) == 0) warp_rays[threadIdx.x >> 5] = rays;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; w++) t += warp_rays[w];
        if (t) atomicAdd(ray_counter, t);
    }
}

// NEE: the shadow rays of one bounce after their closest-hit query - an unoccluded ray adds its contribution to
// the radiance slot of its path (one shadow ray per path and bounce: plain read-modify-write; atomics when the glass
// branches of a path share the slot).
__global__ void __launch_bounds__(256) k_shadow_resolve(DScene s, PathQueue qs, const uint32_t* n_ptr, const float2* hits, f4* L, int glass_branch) {
    const uint32_t n = *n_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        f4 a = qs.a[i], c = qs.c[i]; float2 b = qs.b[i], h = hits[i];
        Ray r; r.o = mk3(a.x, a.y, a.z); r.d = mk3(a.w, b.x, b.y);
        if (nee_visible(s, r, (int)qs.d[i], h.x, __float_as_int(h.y))) {
            const uint32_t slot = (uint32_t)f2i(c.w);
            if (glass_branch) { atomicAdd(&L[slot].x, c.x); atomicAdd(&L[slot].y, c.y); atomicAdd(&L[slot].z, c.z); }   // branches share the slot
            else { f4 v = L[slot]; L[slot] = mk4(v.x + c.x, v.y + c.y, v.z + c.z, 0.f); }
        }
    }
}

// The shade kernel may have tried to allocate past the queue capacity (glass branch mode only).
__global__ void k_clamp_count(uint32_t* n_ptr, uint32_t capacity, uint32_t* high_water, uint32_t* dropped) {
    uint32_t n = *n_ptr;
    if (n > *high_water) *high_water = n;   // the unclamped demand: a full queue and an overflow are told apart
    if (n > capacity) { atomicAdd(dropped, n - capacity); *n_ptr = capacity; }
}

// End of wave: accum[p].rgb += L[s*npix + p] for the k samples of the wave in sample order (fp32,
// the reference's `color += trace(...)`, AccPathTracer.cpp:30), accum[p].a += k.
// The rays of the wave are the sizes of its queues - every entry of every bounce's queue (and of every shadow queue) went
// through one closest-hit query - so thread 0 adds those counters up here instead of every warp of the closest-hit kernels
// sending an atomic per 32 rays.
__global__ void k_accumulate(const f4* L, f4* accum, uint32_t npix, uint32_t k, uint32_t n_samples,
                             const uint32_t* qn, const uint32_t* qr, uint32_t regions, const uint32_t* nshadow, uint32_t counter_stride, uint32_t depth, unsigned long long* ray_counter,
                             const unsigned char* live_flag, uint32_t n_dead) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0 && ray_counter) {
        // the camera rays of the dead pixels are closest-hit queries too - answered by the film rectangles (nothing can be hit)
        unsigned long long rays = depth ? (unsigned long long)n_dead * k : 0ull;
        for (uint32_t d = 0; d < depth; d++) {
            // bounce 0's queue is dense (qn[0]); with regions > 1 the queue entering bounce d >= 1 is counted per region in qr
            if (d == 0 || regions <= 1) rays += qn[(size_t)counter_stride * d];
            else for (uint32_t r = 0; r < regions; r++) rays += qr[(size_t)counter_stride * ((size_t)regions * d + r)];
            if (nshadow && d + 1 < depth) rays += nshadow[(size_t)counter_stride * d];
        }
        *ray_counter += rays;   // one wave accumulates at a time on this counter (its own block of counters)
    }
    if (p >= npix) return;
    f4 acc = accum[p];
    if (!live_flag || live_flag[p]) {   // a dead pixel's samples are black and its radiance slots were never written
        for (uint32_t s = 0; s < k; s++) {
            f4 v = L[(size_t)s * npix + p];
            acc.x += v.x; acc.y += v.y; acc.z += v.z;
        }
    }
    acc.w += (float)n_samples;
    accum[p] = acc;
}

// color /= samples; gamma = sqrt (AccPathTracer.cpp:14-16, 32-34); alpha 1.
__global__ void k_resolve(const f4* accum, f4* rgba, uint32_t npix) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    f4 a = accum[p];
    rgba[p] = mk4(sqrtf(a.x / a.w), sqrtf(a.y / a.w), sqrtf(a.z / a.w), 1.f);
}

// Multi-device resolve: sum the partial linear frames (pointers into this device's and its peers' memory - peer
// loads travel over NVLink) and apply the reference's epilogue in the same pass.
#define NRCU_MAX_DEVICES 16
struct PartialFrames { const f4* part[NRCU_MAX_DEVICES]; int n; };
__global__ void k_resolve_multi(PartialFrames pf, f4* accum_out, f4* rgba, uint32_t npix) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    f4 a = pf.part[0][p];
    for (int g = 1; g < pf.n; g++) { f4 v = pf.part[g][p]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    if (accum_out) accum_out[p] = a;
    rgba[p] = mk4(sqrtf(a.x / a.w), sqrtf(a.y / a.w), sqrtf(a.z / a.w), 1.f);
}

__global__ void k_pack_rays(const float* rays6, uint32_t n, PathQueue q) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = rays6 + 6 * (size_t)i;
    q.a[i] = mk4(r[0], r[1], r[2], r[3]);
    q.b[i] = make_float2(r[4], r[5]);
}

}  // namespace nrcu
