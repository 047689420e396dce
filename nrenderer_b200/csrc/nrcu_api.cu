// nrcu_api.cu — implementation of the C ABI in include/nrcu.h (libnrcuda.so).
//
// Host orchestration only: buffers in HBM, kernel launches on one stream, CUDA-event timing.
// No rendering arithmetic happens on the host except the handful of per-scene scalars the
// reference also computes once on the CPU (Camera ctor, per-node translation of the few explicit
// spheres/triangles/planes, Microfacet's constant Sampler(6) draws).  There is no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "nrcu.h"
#include "nrcu_kernels.cuh"
#include "nrcu_host_prep.hpp"

using namespace nrcu;

namespace {

thread_local std::string g_create_error;

// NRCU_TRACE_UPLOAD=1: print the wall time of the phases of nrcu_upload_scene to stderr (diagnostics)
struct PhaseTimer {
    bool on; std::chrono::steady_clock::time_point t0;
    PhaseTimer() : on(std::getenv("NRCU_TRACE_UPLOAD") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[nrcu upload] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    size_t skew = 0;   // as<T>() starts this many bytes into the allocation (see ensure_wave)
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    cudaError_t ensure(size_t n) {
        n += skew;
        if (n <= bytes && p) return cudaSuccess;
        release();
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n; else p = nullptr;
        return e;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(reinterpret_cast<char*>(p) + skew); }
};

static uint32_t env_u32(const char* name, uint32_t dflt) { const char* e = std::getenv(name); return e ? (uint32_t)std::atoi(e) : dflt; }
static uint32_t wave_skew_kb() { static uint32_t v = env_u32("NRCU_WAVE_SKEW_KB", 68); return v; }
// Words between two device counters of a wave.  Every counter is the target of lane-0 atomics from all warps of a
// kernel; the L2 serialises atomics per address and per slice (tools/micro/atomic_bench.cu: two addresses 128 B apart
// are no faster than one), so the counters that kernels of the concurrent waves hammer at the same time are kept on
// lines of their own.
static size_t counter_stride() { static uint32_t v = env_u32("NRCU_COUNTER_STRIDE", 32); return (size_t)(v < 1 ? 1 : (v > 1024 ? 1024 : v)); }

// sub-allocation of the BVH build arena
struct Carve { char* base; size_t off; void* take(size_t bytes) { void* p = base ? base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; } };
struct Sub { void* p = nullptr; template <typename T> T* as() const { return reinterpret_cast<T*>(p); } };

}  // namespace

#define NRCU_MAX_WAVES 4
#ifndef NRCU_OPT_RAYCOUNT
#define NRCU_OPT_RAYCOUNT 1
#endif
#ifndef NRCU_SHADE_POOL_DEFAULT
#define NRCU_SHADE_POOL_DEFAULT 1
#endif
#ifndef NRCU_QUEUE_REGIONS_DEFAULT
#define NRCU_QUEUE_REGIONS_DEFAULT 16
#endif
#ifndef NRCU_SCHED_DEFAULT
#define NRCU_SCHED_DEFAULT NRCU_SCHED_WAVES
#endif
struct nrcu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string error;
    bool have_scene = false;
    int mode = 0;
    uint32_t spp = 0;
    DScene ds{};
    // scene buffers
    DevBuf prim_geom, prim_shade, prim_box, prim_bound, prim_meta, nodes, leaf_prims, leaf_geom, leaf_box, big_geom, big_box, big_bound, big_meta, big_rect, live_px, dead_px, live_flag, live_count, materials, mat_head, area_lights, env, env_tab;
    // scene-prep sources kept for nrcu_download_primitives
    DevBuf src_a, src_b, sph_pos, sph_rad, sph_mat, tri_v, tri_n, tri_mat, pl_n, pl_p, pl_u, pl_v, pl_mat,
        mesh_voff, mesh_ioff, mesh_pos, mesh_idx, mesh_mat;
    PrimSources ps{};
    // wavefront state
    // wavefront state: one set per concurrent wave (NRCU_WAVES waves run side by side, each on its own stream)
    struct WaveSet { DevBuf qa[2], qb[2], qc[2], qd[2], sa, sb, sc, sd, hits, surv, L, counters; } ws[NRCU_MAX_WAVES];
    cudaStream_t extra_stream[NRCU_MAX_WAVES] = {nullptr, nullptr, nullptr, nullptr};   // [0] unused: wave 0 runs on `stream`
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_acc = nullptr;
    DevBuf accum_own, rgba_dev, build_scratch;
    DevBuf overflow;                       // DScene::overflow: [0] dropped traversal-stack pushes, [1] rays that did not fit the branching queue
    uint32_t* h_flags = nullptr;           // pinned: alive flags of the regeneration scheduler's batches, [wave][ring slot][NRCU_REGEN_FLAGS * stride]
    cudaEvent_t ev_batch[NRCU_MAX_WAVES][4] = {};
    // stats
    float ms_setup = 0.f;
    uint32_t bvh_nodes = 0, n_big = 0;
    uint64_t launches = 0;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
};

#define CTX_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ctx->error = std::string(#call) + ": " + cudaGetErrorString(_e);                        \
            return NRCU_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

#define CTX_LAUNCH_CHECK(name)                                                                      \
    do {                                                                                            \
        cudaError_t _e = cudaGetLastError();                                                        \
        ctx->launches++;                                                                            \
        if (_e != cudaSuccess) {                                                                    \
            ctx->error = std::string("launch of ") + name + ": " + cudaGetErrorString(_e);          \
            return NRCU_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)std::max<size_t>(1, (n + block - 1) / block); }

template <typename T>
static int upload(nrcu_ctx* ctx, DevBuf& buf, const T* src, size_t count) {
    CTX_CUDA(buf.ensure(std::max<size_t>(count, 1) * sizeof(T)));
    if (count) CTX_CUDA(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return NRCU_OK;
}

extern "C" {

int nrcu_abi_version(void) { return NRCU_ABI_VERSION; }

int nrcu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* nrcu_last_error(const nrcu_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int nrcu_create(int device, nrcu_ctx** out) {
    if (!out) { g_create_error = "nrcu_create: out is null"; return NRCU_ERR_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this library has no CPU fallback)";
        cudaGetLastError();
        return NRCU_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) { g_create_error = "nrcu_create: device index out of range"; return NRCU_ERR_INVALID; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return NRCU_ERR_CUDA; }
    nrcu_ctx* ctx = new nrcu_ctx();
    ctx->device = device;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_create_error = std::string("cudaStreamCreate: ") + cudaGetErrorString(e); delete ctx; return NRCU_ERR_CUDA; }
    ctx->own_stream = true;
    cudaEventCreate(&ctx->ev_begin); cudaEventCreate(&ctx->ev_end);
    if (ctx->overflow.ensure(2 * sizeof(uint32_t)) != cudaSuccess || cudaMemset(ctx->overflow.p, 0, 2 * sizeof(uint32_t)) != cudaSuccess) {
        g_create_error = "nrcu_create: device allocation failed"; cudaGetLastError(); nrcu_destroy(ctx); return NRCU_ERR_CUDA;
    }
    *out = ctx;
    return NRCU_OK;
}

int nrcu_destroy(nrcu_ctx* ctx) {
    if (!ctx) return NRCU_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_acc) cudaEventDestroy(ctx->ev_acc);
    for (auto& xs : ctx->extra_stream) if (xs) { cudaStreamSynchronize(xs); cudaStreamDestroy(xs); }
    for (auto& row : ctx->ev_batch) for (auto& e : row) if (e) cudaEventDestroy(e);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return NRCU_OK;
}

void* nrcu_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void nrcu_host_free(void* p) { if (p) cudaFreeHost(p); }

int nrcu_set_stream(nrcu_ctx* ctx, void* cuda_stream) {
    if (!ctx) return NRCU_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return NRCU_OK;
}

// Reads DScene::overflow after the stream has been synchronised; non-zero => the results of the calls since the last
// check cannot be trusted.  The counters are cleared so that the context stays usable.
static int check_overflow(nrcu_ctx* ctx) {
    uint32_t h[2] = {0, 0};
    CTX_CUDA(cudaMemcpy(h, ctx->overflow.p, sizeof(h), cudaMemcpyDeviceToHost));
    if (h[0] == 0 && h[1] == 0) return NRCU_OK;
    CTX_CUDA(cudaMemset(ctx->overflow.p, 0, sizeof(h)));
    char buf[256];
    if (h[0]) std::snprintf(buf, sizeof(buf), "BVH traversal stack overflow: %u entries did not fit %d-entry stacks; hits may have been missed", h[0], ctx->ds.stack_limit);
    else std::snprintf(buf, sizeof(buf), "ray queue overflow: %u rays did not fit their queue (branching glass mode: even at one sample per wave)", h[1]);
    ctx->error = buf;
    return NRCU_ERR_OVERFLOW;
}

int nrcu_synchronize(nrcu_ctx* ctx) {
    if (!ctx) return NRCU_ERR_INVALID;
    CTX_CUDA(cudaSetDevice(ctx->device));
    CTX_CUDA(cudaStreamSynchronize(ctx->stream));
    return check_overflow(ctx);   // asynchronous nrcu_render_accumulate calls report here
}

void nrcu_philox4x32(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
    u32x4 r = philox4x32_10(counter[0], counter[1], counter[2], counter[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// ---------------------------------------------------------------------------------------------
// scene upload
// ---------------------------------------------------------------------------------------------
static int build_bvh(nrcu_ctx* ctx, uint32_t n, float max_abs_coord);
static bool live_pixels() { static uint32_t v = env_u32("NRCU_LIVE_PIXELS", 1); return v != 0; }
static bool film_rects() { static uint32_t v = env_u32("NRCU_FILM_RECTS", 1); return v != 0; }

int nrcu_upload_scene(nrcu_ctx* ctx, const nrcu_scene* sc, int mode) {
    if (!ctx) return NRCU_ERR_INVALID;
    if (!sc || mode < NRCU_MODE_RAYCAST || mode > NRCU_MODE_ACC) { ctx->error = "nrcu_upload_scene: bad arguments"; return NRCU_ERR_INVALID; }
    CTX_CUDA(cudaSetDevice(ctx->device));
    ctx->have_scene = false;
    PhaseTimer pt;
    HostPrep hp;
    std::string why = host_prepare(sc, mode, hp);
    pt.mark("host_prepare");
    if (!why.empty()) { ctx->error = "nrcu_upload_scene: " + why; return NRCU_ERR_INVALID; }
    cudaEvent_t e0, e1;
    CTX_CUDA(cudaEventCreate(&e0)); CTX_CUDA(cudaEventCreate(&e1));
    CTX_CUDA(cudaEventRecord(e0, ctx->stream));
    const uint32_t n = (uint32_t)hp.src_a.size();

    // ---- upload sources -----------------------------------------------------------------------------
    const uint32_t zero_off = 0;
    int rc;
#define UP(buf, ptr, cnt) if ((rc = upload(ctx, ctx->buf, ptr, cnt)) != NRCU_OK) return rc
    UP(src_a, hp.src_a.data(), n); UP(src_b, hp.src_b.data(), n);
    UP(sph_pos, hp.sph.data(), hp.sph.size()); UP(sph_rad, sc->sphere_radius, sc->n_spheres); UP(sph_mat, sc->sphere_material, sc->n_spheres);
    UP(tri_v, hp.tri.data(), hp.tri.size()); UP(tri_n, sc->triangle_normal, 3 * (size_t)sc->n_triangles); UP(tri_mat, sc->triangle_material, sc->n_triangles);
    UP(pl_n, sc->plane_normal, 3 * (size_t)sc->n_planes); UP(pl_p, hp.pln.data(), hp.pln.size());
    UP(pl_u, sc->plane_u, 3 * (size_t)sc->n_planes); UP(pl_v, sc->plane_v, 3 * (size_t)sc->n_planes); UP(pl_mat, sc->plane_material, sc->n_planes);
    UP(mesh_voff, sc->n_meshes ? sc->mesh_vertex_offset : &zero_off, (size_t)sc->n_meshes + 1);
    UP(mesh_ioff, sc->n_meshes ? sc->mesh_index_offset : &zero_off, (size_t)sc->n_meshes + 1);
    UP(mesh_pos, sc->mesh_positions, 3 * (size_t)hp.total_vertices); UP(mesh_idx, sc->mesh_indices, hp.total_indices); UP(mesh_mat, sc->mesh_material, sc->n_meshes);
    UP(materials, hp.materials.data(), hp.materials.size());
    UP(mat_head, hp.mat_head.data(), hp.mat_head.size());
    UP(area_lights, hp.lights.data(), hp.lights.size());
#undef UP
    DScene& ds = ctx->ds;
    fill_scene_scalars(ds, sc, mode, hp);
    ds.overflow = ctx->overflow.as<uint32_t>();
    {   // NRCU_DEBUG_STACK_LIMIT: only for the test of the overflow report (tests/test_gpu_parity.py)
        int lim = (int)env_u32("NRCU_DEBUG_STACK_LIMIT", NRCU_LOCAL_STACK);
        ds.stack_limit = std::min(std::max(lim, NRCU_T3_STACK), NRCU_LOCAL_STACK);
    }
    if (sc->ambient_type == NRCU_AMBIENT_ENVIRONMENT_MAP && sc->ambient_environment_map >= 0 &&
        (uint32_t)sc->ambient_environment_map < sc->n_textures) {
        uint32_t ti = (uint32_t)sc->ambient_environment_map;
        size_t cnt = (size_t)sc->texture_width[ti] * sc->texture_height[ti];
        if (cnt) {
            if ((rc = upload(ctx, ctx->env, reinterpret_cast<const f4*>(sc->texture_rgba + sc->texture_offset[ti]), cnt)) != NRCU_OK) return rc;
            ds.env_rgba = ctx->env.as<f4>(); ds.env_w = (int)sc->texture_width[ti]; ds.env_h = (int)sc->texture_height[ti];
            // importance-sampling tables (NRCU_FLAG_ENV_IS): sin per row, marginal CDF, conditional CDFs + the total weight
            const size_t tab_floats = 2 * (size_t)ds.env_h + cnt;
            CTX_CUDA(ctx->env_tab.ensure(sizeof(float) * (tab_floats + 1)));
            float* tab = ctx->env_tab.as<float>();
            k_env_rows<<<grid_for((size_t)ds.env_h, 64), 64, 0, ctx->stream>>>(ds.env_rgba, ds.env_w, ds.env_h, tab);
            CTX_LAUNCH_CHECK("k_env_rows");
            k_env_marginal<<<1, 1, 0, ctx->stream>>>(ds.env_h, tab, tab + tab_floats);
            CTX_LAUNCH_CHECK("k_env_marginal");
            CTX_CUDA(cudaMemcpyAsync(&ds.env_total, tab + tab_floats, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            CTX_CUDA(cudaStreamSynchronize(ctx->stream));
            ds.env_tab = tab;
        }
    }
    // ---- mesh world transform on the device, once per MESH node, in node order --------------------------
    for (uint32_t e : hp.mesh_nodes) {
        uint32_t first = sc->mesh_vertex_offset[e], cnt = sc->mesh_vertex_offset[e + 1] - first;
        if (!cnt) continue;
        k_mesh_transform<<<grid_for(cnt, 256), 256, 0, ctx->stream>>>(ctx->mesh_pos.as<float>(), first, cnt);
        CTX_LAUNCH_CHECK("k_mesh_transform");
    }
    // ---- flatten into the packed primitive records ----------------------------------------------------
    PrimSources& ps = ctx->ps;
    ps.src_a = ctx->src_a.as<uint32_t>(); ps.src_b = ctx->src_b.as<uint32_t>();
    ps.sphere_position = ctx->sph_pos.as<float>(); ps.sphere_radius = ctx->sph_rad.as<float>(); ps.sphere_material = ctx->sph_mat.as<int>();
    ps.triangle_vertices = ctx->tri_v.as<float>(); ps.triangle_normal = ctx->tri_n.as<float>(); ps.triangle_material = ctx->tri_mat.as<int>();
    ps.plane_normal = ctx->pl_n.as<float>(); ps.plane_position = ctx->pl_p.as<float>(); ps.plane_u = ctx->pl_u.as<float>();
    ps.plane_v = ctx->pl_v.as<float>(); ps.plane_material = ctx->pl_mat.as<int>();
    ps.mesh_vertex_offset = ctx->mesh_voff.as<uint32_t>(); ps.mesh_index_offset = ctx->mesh_ioff.as<uint32_t>();
    ps.mesh_positions = ctx->mesh_pos.as<float>(); ps.mesh_indices = ctx->mesh_idx.as<uint32_t>(); ps.mesh_material = ctx->mesh_mat.as<int>();
    CTX_CUDA(ctx->prim_geom.ensure(sizeof(f4) * 3 * (size_t)std::max(n, 1u)));
    CTX_CUDA(ctx->prim_shade.ensure(sizeof(f4) * (size_t)std::max(n, 1u)));
    CTX_CUDA(ctx->prim_box.ensure(sizeof(f4) * 2 * (size_t)std::max(n, 1u)));
    CTX_CUDA(ctx->prim_bound.ensure(sizeof(f4) * 2 * (size_t)std::max(n, 1u)));
    CTX_CUDA(ctx->prim_meta.ensure(sizeof(uint32_t) * (size_t)std::max(n, 1u)));
    if (n) {
        k_build_prims<<<grid_for(n, 128), 128, 0, ctx->stream>>>(ps, n, mode == NRCU_MODE_RAYCAST, ctx->prim_geom.as<f4>(),
                                                                 ctx->prim_shade.as<f4>(), ctx->prim_box.as<f4>(), ctx->prim_bound.as<f4>(), ctx->prim_meta.as<uint32_t>(), nullptr);
        CTX_LAUNCH_CHECK("k_build_prims");
    }
    ds.prim_geom = ctx->prim_geom.as<f4>(); ds.prim_shade = ctx->prim_shade.as<f4>(); ds.prim_box = ctx->prim_box.as<f4>();
    ds.prim_meta = ctx->prim_meta.as<uint32_t>();
    ds.materials = ctx->materials.as<DMaterial>();
    ds.mat_head = ctx->mat_head.as<f4>();
    ds.area_lights = ctx->area_lights.as<f4>();
    ctx->spp = sc->samples_per_pixel;
    ctx->mode = mode;
    CTX_CUDA(cudaStreamSynchronize(ctx->stream));   // the pageable staging vectors in `hp` die with this scope
    pt.mark("h2d + flatten kernels");

    // ---- BVH ----------------------------------------------------------------------------------------
    ctx->bvh_nodes = 0;
    if (mode != NRCU_MODE_RAYCAST && n > 0) {
        if ((rc = build_bvh(ctx, n, hp.max_abs_coord)) != NRCU_OK) return rc;
    }
    pt.mark("build_bvh (incl. frees)");
    // ---- live pixels: camera rays are only generated where they can meet something (DScene::live_px) ----------------
    ds.live_px = nullptr; ds.live_flag = nullptr; ds.n_live = sc->width * sc->height; ds.dead_px = nullptr; ds.dead_env = 0;
    if (mode != NRCU_MODE_RAYCAST && ds.big_rect && ds.depth > 0 && live_pixels()) {
        const uint32_t npix = sc->width * sc->height;
        CTX_CUDA(ctx->live_px.ensure(sizeof(uint32_t) * (size_t)std::max(npix, 1u)));
        CTX_CUDA(ctx->live_flag.ensure((size_t)std::max(npix, 1u)));
        const unsigned nb = grid_for(npix, 256);
        CTX_CUDA(ctx->live_count.ensure(sizeof(uint32_t) * ((size_t)nb + 1)));   // [0] the total, [1 + b] block counts -> block offsets
        uint32_t* const cnt = ctx->live_count.as<uint32_t>();
        // the film rectangles of the BVH's bounds and of the lights go behind those of the wide primitives (same buffer)
        f4* const rects = ctx->big_rect.as<f4>();
        const uint32_t n_lights = std::min<uint32_t>(ds.n_area_lights, NRCU_LIVE_LIGHTS), n_rect = ds.n_big + 1u + n_lights;
        k_scene_rects<<<1, 64, 0, ctx->stream>>>(ds, rects); CTX_LAUNCH_CHECK("k_scene_rects");
        k_live_flags<<<nb, 256, 0, ctx->stream>>>(ds, rects, n_rect, ds.n_area_lights > NRCU_LIVE_LIGHTS ? 1 : 0, ctx->live_flag.as<unsigned char>(), cnt + 1); CTX_LAUNCH_CHECK("k_live_flags");
        k_live_scan<<<1, 1024, 0, ctx->stream>>>(cnt + 1, nb, cnt); CTX_LAUNCH_CHECK("k_live_scan");
        k_live_scatter<<<nb, 256, 0, ctx->stream>>>(ctx->live_flag.as<unsigned char>(), npix, cnt + 1, ctx->live_px.as<uint32_t>(), 1); CTX_LAUNCH_CHECK("k_live_scatter");
        const bool dead_env = ds.env_rgba && mode == NRCU_MODE_ACC;   // dead pixels see the environment map (k_env_dead)
        if (dead_env) {
            CTX_CUDA(ctx->dead_px.ensure(sizeof(uint32_t) * (size_t)std::max(npix, 1u)));
            k_live_scatter<<<nb, 256, 0, ctx->stream>>>(ctx->live_flag.as<unsigned char>(), npix, cnt + 1, ctx->dead_px.as<uint32_t>(), 0); CTX_LAUNCH_CHECK("k_live_scatter");
        }
        uint32_t n_live = npix;
        CTX_CUDA(cudaMemcpyAsync(&n_live, cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CTX_CUDA(cudaStreamSynchronize(ctx->stream));
        if (n_live < npix) {
            ds.live_px = ctx->live_px.as<uint32_t>(); ds.live_flag = ctx->live_flag.as<unsigned char>(); ds.n_live = n_live;
            if (dead_env) { ds.dead_px = ctx->dead_px.as<uint32_t>(); ds.dead_env = 1; }
        }
    }
    pt.mark("live pixels");
    CTX_CUDA(cudaEventRecord(e1, ctx->stream));
    CTX_CUDA(cudaEventSynchronize(e1));
    CTX_CUDA(cudaEventElapsedTime(&ctx->ms_setup, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctx->have_scene = true;
    return NRCU_OK;
}

static int build_bvh(nrcu_ctx* ctx, uint32_t n, float max_abs_coord) {
    const int cap = 2 * (int)n + 2;
    const int bin_nodes = std::max(64, (int)(0.4 * n) + 8);
    // Build scratch comes from ONE arena that lives in the context and only ever grows: cudaMalloc/cudaFree per
    // upload cost 5 ms in the good case and hundreds of ms when gigabytes of wave buffers are mapped (measured).
    Sub prim_node, nbox, cbox, ncount, nidmin, nidmax, nstate, nchild, nsplit_axis, nsplit_pos, ndepth, nleaf_first, nleaf_fill,
        nwide, bins, counters, nbin_slot, wide_tmp, big_count, big_cand;
    Sub* per_node[] = {&ncount, &nidmin, &nidmax, &nstate, &nchild, &nsplit_axis, &nsplit_pos, &ndepth, &nleaf_first, &nleaf_fill, &nwide, &nbin_slot};
    for (int pass = 0; pass < 2; pass++) {   // pass 0 sizes the arena, pass 1 hands out the pointers
        Carve cv{pass ? ctx->build_scratch.as<char>() : nullptr, 0};
        prim_node.p = cv.take(sizeof(int) * (size_t)n);
        nbox.p = cv.take(sizeof(int) * 6 * (size_t)cap); cbox.p = cv.take(sizeof(int) * 6 * (size_t)cap);
        for (Sub* b : per_node) b->p = cv.take(sizeof(int) * (size_t)cap);
        bins.p = cv.take(sizeof(int) * (size_t)bin_nodes * 3 * NRCU_NBINS * NRCU_BIN_WORDS);
        counters.p = cv.take(sizeof(int) * 8); big_count.p = cv.take(sizeof(int) * 2); big_cand.p = cv.take(sizeof(int) * NRCU_BIG_CAND_CAP);
        wide_tmp.p = cv.take(sizeof(f4) * NRCU_BVH_NODE_F4 * (size_t)std::max(1u, n));
        if (pass == 0) CTX_CUDA(ctx->build_scratch.ensure(cv.off));
    }
    CTX_CUDA(ctx->leaf_prims.ensure(sizeof(uint32_t) * (size_t)n));
    CTX_CUDA(ctx->leaf_geom.ensure(sizeof(f4) * 3 * (size_t)n)); CTX_CUDA(ctx->leaf_box.ensure(sizeof(f4) * 2 * (size_t)n));
    CTX_CUDA(ctx->big_geom.ensure(sizeof(f4) * 3 * NRCU_MAX_BIG)); CTX_CUDA(ctx->big_box.ensure(sizeof(f4) * 2 * NRCU_MAX_BIG));
    CTX_CUDA(ctx->big_meta.ensure(sizeof(uint32_t) * NRCU_MAX_BIG)); CTX_CUDA(ctx->big_bound.ensure(sizeof(f4) * 2 * NRCU_MAX_BIG));

    PhaseTimer pt;
    pt.mark("  bvh: allocations");
    BvhBuild b{};
    b.n_prims = n; b.prim_box = ctx->prim_box.as<f4>(); b.prim_bound = ctx->prim_bound.as<f4>(); b.prim_meta = ctx->prim_meta.as<uint32_t>();
    b.prim_node = prim_node.as<int>(); b.nbox = nbox.as<int>(); b.cbox = cbox.as<int>(); b.ncount = ncount.as<int>();
    b.nidmin = nidmin.as<int>(); b.nidmax = nidmax.as<int>(); b.nstate = nstate.as<int>(); b.nchild = nchild.as<int>();
    b.nsplit_axis = nsplit_axis.as<int>(); b.nsplit_pos = nsplit_pos.as<float>(); b.ndepth = ndepth.as<int>();
    b.nleaf_first = nleaf_first.as<int>(); b.nleaf_fill = nleaf_fill.as<int>(); b.nwide = nwide.as<int>();
    b.bins = bins.as<int>(); b.bin_nodes = bin_nodes; b.counters = counters.as<int>(); b.nbin_slot = nbin_slot.as<int>();
    b.leaf_prims = ctx->leaf_prims.as<uint32_t>(); b.wide_nodes = wide_tmp.as<f4>();
    b.prim_geom = ctx->prim_geom.as<f4>(); b.leaf_geom = ctx->leaf_geom.as<f4>(); b.leaf_box = ctx->leaf_box.as<f4>();
    b.big_geom = ctx->big_geom.as<f4>(); b.big_box = ctx->big_box.as<f4>(); b.big_bound = ctx->big_bound.as<f4>(); b.big_meta = ctx->big_meta.as<uint32_t>(); b.big_count = big_count.as<int>();
    b.big_cand = big_cand.as<int>(); b.big_cand_count = big_count.as<int>() + 1;
    b.inflate = max_abs_coord * (1.0f / 65536.0f);
    cudaStream_t st = ctx->stream;
    const int T = 128;

    int h_counters[8] = {1, 0, 0, 0, 0, 0, 0, 0};   // node count starts at 1 (the root)
    CTX_CUDA(cudaMemcpyAsync(counters.p, h_counters, sizeof(h_counters), cudaMemcpyHostToDevice, st));
    k_bvh_clear<<<grid_for(cap, T), T, 0, st>>>(b, 0, cap); CTX_LAUNCH_CHECK("k_bvh_clear");
    k_bvh_init_prim<<<grid_for(n, T), T, 0, st>>>(b, 0, (int)n); CTX_LAUNCH_CHECK("k_bvh_init_prim");
    // wide primitives leave the tree; node 0 is rebuilt from the rest
    CTX_CUDA(cudaMemsetAsync(big_count.p, 0, 2 * sizeof(int), st));
    k_bvh_big_candidate<<<grid_for(n, T), T, 0, st>>>(b, 0, (int)n); CTX_LAUNCH_CHECK("k_bvh_big_candidate");
    k_bvh_select_big<<<1, 1, 0, st>>>(b, 0, 1); CTX_LAUNCH_CHECK("k_bvh_select_big");
    k_bvh_clear<<<1, 1, 0, st>>>(b, 0, 1); CTX_LAUNCH_CHECK("k_bvh_clear");
    k_bvh_init_prim_rest<<<grid_for(n, T), T, 0, st>>>(b, 0, (int)n); CTX_LAUNCH_CHECK("k_bvh_init_prim_rest");
    int n_big = 0, root_box[6];
    CTX_CUDA(cudaMemcpyAsync(&n_big, big_count.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaMemcpyAsync(root_box, nbox.p, sizeof(root_box), cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaStreamSynchronize(st));
    pt.mark("  bvh: init + wide list");
    ctx->n_big = (uint32_t)n_big;
    DScene& ds = ctx->ds;
    ds.big_geom = ctx->big_geom.as<f4>(); ds.big_box = ctx->big_box.as<f4>(); ds.big_bound = ctx->big_bound.as<f4>(); ds.big_meta = ctx->big_meta.as<uint32_t>(); ds.n_big = (uint32_t)n_big;
    // film rectangles of the wide primitives for the camera rays of a pinhole camera (big_list_mask_film); NRCU_FILM_RECTS=0: off
    ds.big_rect = nullptr;
    if (n_big > 0 && ds.cam.lens_radius == 0.f && film_rects()) {
        CTX_CUDA(ctx->big_rect.ensure(sizeof(f4) * NRCU_LIVE_RECTS));   // + the BVH's bounds and the lights (live pixels)
        k_big_rects<<<1, NRCU_MAX_BIG, 0, st>>>(ds, ctx->big_rect.as<f4>()); CTX_LAUNCH_CHECK("k_big_rects");
        ds.big_rect = ctx->big_rect.as<f4>();
    }
    ds.nodes = nullptr; ds.leaf_prims = ctx->leaf_prims.as<uint32_t>(); ds.leaf_geom = ctx->leaf_geom.as<f4>(); ds.leaf_box = ctx->leaf_box.as<f4>();
    ds.root_ref = NRCU_REF_EMPTY; ds.bvh_lo = mk3(NRCU_INF); ds.bvh_hi = mk3(-NRCU_INF);
    ctx->bvh_nodes = 0;
    const uint32_t n_rest = n - (uint32_t)n_big;
    if (n_rest == 0) return NRCU_OK;   // everything is in the wide list
    bvh_padded_bounds(root_box, b.inflate, ds.bvh_lo, ds.bvh_hi);

    int begin = 0, end = 1;
    bool converged = false;
    for (int level = 0; level < 128; level++) {
        b.level_begin = begin; b.level_end = end;
        CTX_CUDA(cudaMemsetAsync(counters.as<int>() + 3, 0, 2 * sizeof(int), st));
        k_bvh_level_prepare<<<grid_for(end - begin, T), T, 0, st>>>(b, begin, end - begin); CTX_LAUNCH_CHECK("k_bvh_level_prepare");
        k_bvh_bin<<<grid_for(n, T), T, 0, st>>>(b, 0, (int)n); CTX_LAUNCH_CHECK("k_bvh_bin");
        k_bvh_split<<<grid_for(end - begin, T), T, 0, st>>>(b, begin, end - begin); CTX_LAUNCH_CHECK("k_bvh_split");
        CTX_CUDA(cudaMemcpyAsync(h_counters, counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
        CTX_CUDA(cudaStreamSynchronize(st));
        if (h_counters[3] == 0) { converged = true; break; }
        k_bvh_partition<<<grid_for(n, T), T, 0, st>>>(b, 0, (int)n); CTX_LAUNCH_CHECK("k_bvh_partition");
        begin = end; end = h_counters[0];
    }
    pt.mark("  bvh: level loop");
    if (!converged) {   // unreachable with the depth cap of bvh_split (<= 24 + 32 levels); never emit a tree with OPEN nodes
        ctx->error = "BVH build did not converge within 128 levels";
        return NRCU_ERR_INVALID;
    }
    const int n_nodes = h_counters[0];
    k_bvh_leaf_alloc<<<grid_for(n_nodes, T), T, 0, st>>>(b, 0, n_nodes); CTX_LAUNCH_CHECK("k_bvh_leaf_alloc");
    k_bvh_leaf_fill<<<grid_for(n, T), T, 0, st>>>(b, 0, (int)n); CTX_LAUNCH_CHECK("k_bvh_leaf_fill");
    k_bvh_leaf_sort<<<grid_for(n_nodes, T), T, 0, st>>>(b, 0, n_nodes); CTX_LAUNCH_CHECK("k_bvh_leaf_sort");
    k_bvh_leaf_gather<<<grid_for(n_rest, T), T, 0, st>>>(b, 0, (int)n_rest); CTX_LAUNCH_CHECK("k_bvh_leaf_gather");
    k_bvh_wide_index<<<grid_for(n_nodes, T), T, 0, st>>>(b, 0, n_nodes); CTX_LAUNCH_CHECK("k_bvh_wide_index");
    k_bvh_wide_emit<<<grid_for(n_nodes, T), T, 0, st>>>(b, 0, n_nodes); CTX_LAUNCH_CHECK("k_bvh_wide_emit");
    // root reference: wide index of node 0, or a leaf reference when everything left fits one leaf
    int root_state = 0, root_wide = -1, root_cnt = 0, root_first = 0;
    CTX_CUDA(cudaMemcpyAsync(h_counters, counters.p, sizeof(h_counters), cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaMemcpyAsync(&root_state, nstate.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaMemcpyAsync(&root_wide, nwide.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaMemcpyAsync(&root_cnt, ncount.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaMemcpyAsync(&root_first, nleaf_first.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaStreamSynchronize(st));
    const int n_wide = h_counters[2];
    CTX_CUDA(ctx->nodes.ensure(sizeof(f4) * NRCU_BVH_NODE_F4 * (size_t)std::max(1, n_wide)));
    if (n_wide) CTX_CUDA(cudaMemcpyAsync(ctx->nodes.p, wide_tmp.p, sizeof(f4) * NRCU_BVH_NODE_F4 * (size_t)n_wide, cudaMemcpyDeviceToDevice, st));
    CTX_CUDA(cudaStreamSynchronize(st));
    pt.mark("  bvh: leaves + wide emit");
    ds.nodes = ctx->nodes.as<f4>();
    ds.root_ref = root_state == BNODE_LEAF ? ~((root_first << 4) | (root_cnt - 1)) : root_wide;
    ctx->bvh_nodes = (uint32_t)n_wide;
    return NRCU_OK;
}

int nrcu_primitive_count(const nrcu_ctx* ctx, uint32_t* out) {
    if (!ctx || !out) return NRCU_ERR_INVALID;
    if (!ctx->have_scene) return NRCU_ERR_STATE;
    *out = ctx->ds.n_prims;
    return NRCU_OK;
}

int nrcu_download_primitives(const nrcu_ctx* cctx, uint32_t* kind, float* data16, int32_t* material) {
    nrcu_ctx* ctx = const_cast<nrcu_ctx*>(cctx);
    if (!ctx) return NRCU_ERR_INVALID;
    if (!ctx->have_scene) { ctx->error = "no scene uploaded"; return NRCU_ERR_STATE; }
    CTX_CUDA(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->ds.n_prims;
    if (!n) return NRCU_OK;
    if (data16) {
        DevBuf ex, g, s, bx, bd, mt;   // scratch outputs so the live scene is not disturbed
        CTX_CUDA(ex.ensure(sizeof(float) * 16 * (size_t)n)); CTX_CUDA(g.ensure(sizeof(f4) * 3 * (size_t)n));
        CTX_CUDA(s.ensure(sizeof(f4) * (size_t)n)); CTX_CUDA(bx.ensure(sizeof(f4) * 2 * (size_t)n)); CTX_CUDA(bd.ensure(sizeof(f4) * 2 * (size_t)n)); CTX_CUDA(mt.ensure(sizeof(uint32_t) * (size_t)n));
        k_build_prims<<<grid_for(n, 128), 128, 0, ctx->stream>>>(ctx->ps, n, ctx->mode == NRCU_MODE_RAYCAST, g.as<f4>(), s.as<f4>(), bx.as<f4>(), bd.as<f4>(), mt.as<uint32_t>(), ex.as<float>());
        CTX_LAUNCH_CHECK("k_build_prims");
        CTX_CUDA(cudaMemcpyAsync(data16, ex.p, sizeof(float) * 16 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CTX_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (kind || material) {
        std::vector<uint32_t> meta(n);
        CTX_CUDA(cudaMemcpyAsync(meta.data(), ctx->prim_meta.p, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CTX_CUDA(cudaStreamSynchronize(ctx->stream));
        for (uint32_t i = 0; i < n; i++) { if (kind) kind[i] = meta[i] & 3u; if (material) material[i] = (int32_t)(meta[i] >> 2); }
    }
    return NRCU_OK;
}

// ---------------------------------------------------------------------------------------------
// rendering
// ---------------------------------------------------------------------------------------------
enum { CNT_RAYS = 0 /* u64 */, CNT_HIGH_WATER = 2, CNT_QUEUE0 = 64 /* [depth+2] queue sizes, [depth+2] fetch cursors, [depth+2] survivor counts */ };
#define NRCU_MAX_REGIONS 32   /* queue regions (QRegions in nrcu_kernels.cuh): lane r of a warp keeps region r's count */
// Regions per queue (power of two, 1 = the plain compacted queue).  NRCU_QUEUE_REGIONS overrides.
static uint32_t queue_regions_log2() {
    static int v = -1;
    if (v < 0) { uint32_t k = env_u32("NRCU_QUEUE_REGIONS", NRCU_QUEUE_REGIONS_DEFAULT); v = 0; while ((2u << v) <= k && (2u << v) <= NRCU_MAX_REGIONS) v++; }
    return (uint32_t)v;
}

static int ensure_wave(nrcu_ctx* ctx, int set, uint32_t slots, uint32_t capacity, uint32_t depth, bool branch_bits, bool shadow_queue) {
    // The wave buffers are power-of-two sized (32 Mi x 16 B = 512 MiB) and the kernels stream through ten of
    // them at the same index; each buffer starts at its own skew inside its allocation so that the streams do
    // not share an HBM channel/bank phase (k_shade has been measured anywhere between 39 and 54 ms per 128 spp on
    // different boxes with identical code; the skew did not change that, it is kept as a cheap precaution).
    nrcu_ctx::WaveSet& w = ctx->ws[set];
    const size_t S = (size_t)wave_skew_kb() << 10;
    DevBuf* bufs[] = {&w.qa[0], &w.qb[0], &w.qc[0], &w.qa[1], &w.qb[1], &w.qc[1], &w.hits, &w.surv, &w.L};
    for (size_t j = 0; j < sizeof(bufs) / sizeof(bufs[0]); j++) bufs[j]->skew = (j + 1 + 9 * (size_t)set) * S;
    // the kernels read the queues in whole blocks of 32 entries (lanes past the live count load entries nobody uses): pad
    const size_t padded = (size_t)capacity + 32;
    for (int k = 0; k < 2; k++) {
        CTX_CUDA(w.qa[k].ensure(sizeof(f4) * padded));
        CTX_CUDA(w.qb[k].ensure(sizeof(float2) * padded));
        CTX_CUDA(w.qc[k].ensure(sizeof(f4) * padded));
    }
    CTX_CUDA(w.hits.ensure(sizeof(float2) * padded));
    if (branch_bits) for (int k = 0; k < 2; k++) CTX_CUDA(w.qd[k].ensure(sizeof(uint32_t) * padded));
    if (shadow_queue) {   // NEE: shadow rays of one bounce (ray, contribution + slot, light index)
        CTX_CUDA(w.sa.ensure(sizeof(f4) * (size_t)capacity)); CTX_CUDA(w.sb.ensure(sizeof(float2) * (size_t)capacity));
        CTX_CUDA(w.sc.ensure(sizeof(f4) * (size_t)capacity)); CTX_CUDA(w.sd.ensure(sizeof(uint32_t) * (size_t)capacity));
    }
    CTX_CUDA(w.surv.ensure(sizeof(uint32_t) * (size_t)capacity));
    CTX_CUDA(w.L.ensure(sizeof(f4) * (size_t)slots));
    CTX_CUDA(w.counters.ensure(sizeof(uint32_t) * (CNT_QUEUE0 + counter_stride() * ((6 + NRCU_MAX_REGIONS) * ((size_t)depth + 4)))));
    return NRCU_OK;
}

}  // extern "C"

static int sm_count(int device) {
    static int cached[64] = {0};
    if (device < 64 && cached[device]) return cached[device];
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    if (device < 64) cached[device] = n;
    return n;
}

// Traversal kernel variant and its refill threshold (tuning knobs; NRCU_TRACE_VARIANT=1 selects the
// first, batch-of-32 kernel kept for A/B measurements).
static int trace_variant() { static int v = -1; if (v < 0) { const char* e = std::getenv("NRCU_TRACE_VARIANT"); v = e ? std::atoi(e) : 2; } return v; }
static uint32_t trace_refill() { static int v = -1; if (v < 0) { const char* e = std::getenv("NRCU_TRACE_REFILL"); v = e ? std::atoi(e) : 8; if (v < 1) v = 1; if (v > 32) v = 32; } return (uint32_t)v; }
static uint32_t wave_slots_target() { static uint32_t v = env_u32("NRCU_WAVE_MSLOTS", 256) << 20; return v; }
static uint32_t wave_min_groups() { static uint32_t v = std::max<uint32_t>(1, env_u32("NRCU_WAVE_MIN_GROUPS", 1)); return v; }
static uint32_t trace_taper(int k) {
    static uint32_t v[2] = {0xffffffffu, 0xffffffffu}; static bool init = false;
    if (!init) { init = true; const char* e = std::getenv("NRCU_TRACE_TAPER"); unsigned a = 0xffffffffu, b = 0xffffffffu; if (e) std::sscanf(e, "%u,%u", &a, &b); v[0] = a; v[1] = b; }
    return v[k];
}
static uint32_t big_balanced() { static uint32_t v = env_u32("NRCU_BIG_BALANCED", 2); return v; }   // 0 per-lane k_big, 1 k_big_balanced, 2 k_big_balanced64 (needs NRCU_OPT_RAYCOUNT)
static bool shade_pool() { static uint32_t v = env_u32("NRCU_SHADE_POOL", NRCU_SHADE_POOL_DEFAULT); return v != 0; }   // k_shade_pool instead of k_shade where it applies
static unsigned dual_big64() { static uint32_t v = env_u32("NRCU_CONC_BIG64", 12); return v ? v : 1; }
static uint32_t trace_w_node() { static uint32_t v = env_u32("NRCU_TRACE_WNODE", 1); return v; }
static uint32_t trace_w_prim() { static uint32_t v = env_u32("NRCU_TRACE_WPRIM", 1); return v; }
static unsigned trace_blocks_per_sm() { static int v = -1; if (v < 0) { const char* e = std::getenv("NRCU_TRACE_BLOCKS"); v = e ? std::atoi(e) : 8; if (v < 1) v = 1; } return (unsigned)v; }

static int concurrent_waves() { static uint32_t v = env_u32("NRCU_WAVES", 2); return (int)std::min<uint32_t>(std::max<uint32_t>(v, 1), NRCU_MAX_WAVES); }
// CTAs per SM of each kernel when several waves share the machine (tuning knobs; full-size grids measured best)
static unsigned dual_big() { static uint32_t v = env_u32("NRCU_CONC_BIG", 8); return v ? v : 1; }
static unsigned dual_shade() { static uint32_t v = env_u32("NRCU_CONC_SHADE", 4); return v ? v : 1; }
static unsigned dual_trace() { static uint32_t v = env_u32("NRCU_CONC_TRACE", 8); return v ? v : 1; }

// Stage 2 of the closest hit: BVH traversal of the *n_surv rays listed in `surv`, refining hits[] in place.
template <bool GATE>
static void launch_stage2(nrcu_ctx* ctx, cudaStream_t st, unsigned share, const DScene& ds, PathQueue q, float2* hits, const uint32_t* surv, const uint32_t* n_surv,
                          uint32_t* fetch, unsigned long long* rays, uint32_t bounce = 0) {
    // Deep bounces hold few rays; a smaller persistent grid has a lower latency floor (fewer CTAs to start,
    // fewer warps contending for the fetch counter) - measured: it does not, smaller grids are simply slower (6,12: -5 %),
    // so tapering is off by default.  NRCU_TRACE_TAPER="a,b" halves the grid from bounce a and again from b.
    unsigned per_sm = share > 1 ? dual_trace() : trace_blocks_per_sm();   // share = 2: two waves run side by side, about half an SM each
    if (bounce >= trace_taper(0)) per_sm = std::max(1u, per_sm / 2);
    if (bounce >= trace_taper(1)) per_sm = std::max(1u, per_sm / 2);
    const unsigned grid = (unsigned)sm_count(ctx->device) * per_sm;
    if (trace_variant() == 3) k_trace3<GATE><<<grid, NRCU_TRACE_THREADS, 0, st>>>(ds, q, n_surv, surv, hits, fetch, rays, trace_refill(), trace_w_node(), trace_w_prim());
    else if (trace_variant() == 4) k_trace2<GATE, true><<<grid, NRCU_TRACE_THREADS, 0, st>>>(ds, q, n_surv, surv, hits, fetch, rays, trace_refill());
    else k_trace2<GATE, false><<<grid, NRCU_TRACE_THREADS, 0, st>>>(ds, q, n_surv, surv, hits, fetch, rays, trace_refill());
}

// Closest hit for the first *n_ptr entries of queue `q` into hits[] (the stand-alone form used by
// nrcu_trace_batch; the renderer fuses stage 1 into the kernels that generate the rays): stage 1 (wide
// primitives, every ray) then stage 2 (BVH traversal of the survivors).
template <bool GATE>
static void launch_closest_hit(nrcu_ctx* ctx, cudaStream_t st, unsigned share, const DScene& ds, PathQueue q, const uint32_t* n_ptr, float2* hits, uint32_t* surv,
                               uint32_t* n_surv, uint32_t* fetch, unsigned long long* rays, int* launches) {
    k_big<GATE><<<(unsigned)sm_count(ctx->device) * (share > 1 ? dual_big() : 8), 256, 0, st>>>(ds, q, QRegions{n_ptr, 0u, 1u, 0xffffffffu}, hits, surv, n_surv, rays);
    (*launches)++;
    if (ds.root_ref == NRCU_REF_EMPTY) return;
    launch_stage2<GATE>(ctx, st, share, ds, q, hits, surv, n_surv, fetch, rays);
    (*launches)++;
}

extern "C" {

static int direct_sampling_mode(const nrcu_ctx* ctx, const nrcu_render_params* params) {
    if (!params) return 0;
    if ((params->flags & NRCU_FLAG_ENV_IS) && ctx->mode == NRCU_MODE_ACC && ctx->ds.env_rgba && ctx->ds.env_tab && ctx->ds.env_total > 0.f) return 2;
    if ((params->flags & NRCU_FLAG_NEE) && ctx->ds.n_area_lights > 0) return 1;
    return 0;
}

static cudaEvent_t pool_event(nrcu_ctx* ctx, size_t i) {
    while (ctx->ev_pool.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); ctx->ev_pool.push_back(e); }
    return ctx->ev_pool[i];
}

#if NRCU_OPT_RAYCOUNT
#define NRCU_RC_K nullptr      /* closest-hit kernels: no per-warp ray-counter atomics ... */
#define NRCU_RC_A pp.d_rays    /* ... the wave's rays are summed from its queue sizes by k_accumulate */
#else
#define NRCU_RC_K pp.d_rays
#define NRCU_RC_A nullptr
#endif
static int render_waves(nrcu_ctx* ctx, const nrcu_render_params* params, f4* d_accum, nrcu_stats* stats) {
    // direct sampling at Lambertian vertices (extensions): the environment map when the scene has one and NRCU_FLAG_ENV_IS is
    // set, else the area lights under NRCU_FLAG_NEE; one shadow ray per vertex either way
    ctx->ds.nee = direct_sampling_mode(ctx, params);
    const bool nee = ctx->ds.nee != 0;
    const DScene& ds = ctx->ds;
    const uint32_t npix = ds.width * ds.height;
    uint32_t s0 = params ? params->sample_begin : 0, s1 = params ? params->sample_end : 0;
    if (s0 == 0 && s1 == 0) s1 = ctx->spp;
    if (s1 < s0) { ctx->error = "sample_end < sample_begin"; return NRCU_ERR_INVALID; }
    const uint64_t seed = params ? params->seed : 0;
    const int glass_branch = params && params->glass_mode == NRCU_GLASS_BRANCH;
    // Wave size: k samples of every live pixel; NRCU_WAVE_MSLOTS (default 256 Mi, ~29 GB of the 180 GB HBM) queue entries
    // are in flight, split over NRCU_WAVES (default 2) waves that run side by side on their own streams with full-size
    // persistent grids.  Bigger waves amortise the latency floor of the deep bounces (few rays; the longest single
    // traversal ends every stage-2 launch on a nearly idle machine), and the second wave's kernels run in those tails.
    // Measured on cfg3 (profiles/r2_history.md): one wave 3.0, two 3.4 Gpath-samples/s before the live pixels; round 1, with
    // the shading kernel waiting for its slot atomics, 2.14 against 2.68.  The accumulation order is fixed by events, so the
    // image does not depend on the number of waves.
    // camera rays exist for the live pixels only (DScene::live_px): a wave holds k x n_live queue entries and k x n_pixels radiance slots
    const uint32_t nlive = ds.live_px ? ds.n_live : npix;
    const bool explicit_k = params && params->samples_per_wave != 0;
    uint32_t k = explicit_k ? params->samples_per_wave : std::max<uint32_t>(1, wave_slots_target() / std::max(nlive, 1u));
    k = std::min<uint32_t>(k, std::max<uint32_t>(1, s1 - s0));
    if ((uint64_t)k * npix > 0x7fffffffull) k = std::max<uint32_t>(1, (uint32_t)(0x7fffffffull / npix));
    int NP = glass_branch ? 1 : std::max(1, std::min<int>(concurrent_waves(), (int)(s1 - s0)));
    if (NP > 1 && !explicit_k) k = std::max<uint32_t>(1, k / (uint32_t)NP);
    // NRCU_WAVE_MIN_GROUPS (default 1 = off): at least that many groups of concurrent waves where the samples allow it.
    // Measured on a 128-spp slice (what each GPU renders at N = 8) once the host no longer waits for statistics after
    // every frame: one group of two 64-spp waves 3109, two groups of 32-spp waves 3078-3088, three groups 3068
    // Mpath-samples/s - fewer, bigger launches win even for short slices (profiles/r2_history.md)
    if (!explicit_k && !glass_branch) k = std::max<uint32_t>(1, std::min<uint32_t>(k, (s1 - s0 + (uint32_t)NP * wave_min_groups() - 1) / ((uint32_t)NP * wave_min_groups())));
    // Equal groups: with k from the slot budget alone the samples left over after the last full group form a small group of
    // their own (512 spp at 1080p: 115 + 115, 115 + 115 and then 26 + 26 samples, each group with its own chain of ~60 launches
    // and its own depth-20 tail: -1.5 %).  The samples are spread evenly over the groups instead, and one group fewer is
    // taken when that costs at most 12 % more slots per wave than the budget.
    if (!explicit_k && !glass_branch && s1 > s0) {
        const uint32_t total = s1 - s0, per_group = (uint32_t)NP * k;
        uint32_t groups = (total + per_group - 1) / per_group;
        if (groups > 1 && (double)total <= 1.12 * (double)(groups - 1) * (double)per_group) groups--;
        const uint32_t kk = (total + groups * (uint32_t)NP - 1) / (groups * (uint32_t)NP);
        if ((uint64_t)kk * npix <= 0x7fffffffull) k = std::max<uint32_t>(1, kk);
    }
    NP = std::max(1, std::min<int>(NP, (int)((s1 - s0 + k - 1) / k)));   // also with no samples at all: one (idle) wave set
    const unsigned share = (unsigned)NP;
    uint32_t slots = k * npix, qslots = k * nlive;   // radiance slots / bounce-0 queue entries of a wave
    // branching glass mode: room for 4 rays per path slot, and never less than 4 Mi entries (small frames at many bounces)
    auto branch_capacity = [](uint32_t sl) { return (uint32_t)std::min<uint64_t>(0x7fffffffull, std::max<uint64_t>((uint64_t)sl * 4, 4ull << 20)); };
    uint32_t capacity = glass_branch ? branch_capacity(qslots) : qslots;
    // Queue regions (QRegions in nrcu_kernels.cuh): K counters per queue instead of one.  Region r of the queue that enters
    // bounce d + 1 receives the survivors of the input blocks pb = r (mod K), at most ceil(blocks / K) x 32 entries, so the
    // array extent grows by at most 32 K entries per bounce: `slack`.  The branching glass mode keeps the plain queue
    // (its overflow handling clamps ONE counter).
    const uint32_t logk = glass_branch ? 0u : queue_regions_log2(), K = 1u << logk;
    // k_shade_pool numbers its output rounds per warp: a region can be one round per shading warp longer than the others.
    const uint32_t pool_slack = (shade_pool() && !glass_branch && direct_sampling_mode(ctx, params) == 0) ? 32u * (K + 1u) * (uint32_t)sm_count(ctx->device) * std::max(dual_shade(), (unsigned)NRCU_SHADE_MINB) * 8u : 0u;
    const uint32_t slack = (K > 1 ? 32u * K * (ds.depth + 2) : 0u) + (K > 1 ? pool_slack : 0u);
    int rc;
    for (;;) {   // the default wave size assumes a B200's 180 GB; on a fuller or smaller device shrink the waves instead of failing
        rc = NRCU_OK;
        for (int p = 0; p < NP && rc == NRCU_OK; p++) rc = ensure_wave(ctx, p, slots, capacity + slack, ds.depth, glass_branch != 0, nee);
        if (rc == NRCU_OK) break;
        cudaError_t last = cudaGetLastError();
        (void)last;
        if (explicit_k || k == 1 || ctx->error.find("out of memory") == std::string::npos) return rc;
        for (int p = 0; p < NP; p++) {
            nrcu_ctx::WaveSet& w = ctx->ws[p];
            DevBuf* all[] = {&w.qa[0], &w.qa[1], &w.qb[0], &w.qb[1], &w.qc[0], &w.qc[1], &w.qd[0], &w.qd[1], &w.sa, &w.sb, &w.sc, &w.sd, &w.hits, &w.surv, &w.L};
            for (DevBuf* bfr : all) bfr->release();
        }
        k = std::max<uint32_t>(1, k / 2);
        slots = k * npix; qslots = k * nlive;
        capacity = glass_branch ? branch_capacity(qslots) : qslots;
    }
    cudaStream_t S[NRCU_MAX_WAVES] = {ctx->stream, ctx->stream, ctx->stream, ctx->stream};
    if (NP > 1) {
        for (int p = 1; p < NP; p++) {
            if (!ctx->extra_stream[p]) CTX_CUDA(cudaStreamCreateWithFlags(&ctx->extra_stream[p], cudaStreamNonBlocking));
            S[p] = ctx->extra_stream[p];
        }
        if (!ctx->ev_fork) { CTX_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)); CTX_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
                             CTX_CUDA(cudaEventCreateWithFlags(&ctx->ev_acc, cudaEventDisableTiming)); }
    }
    // the high-water mark lives in set 0's counter block; every wave counts its rays in its own block
    uint32_t* cnt0 = ctx->ws[0].counters.as<uint32_t>();
    const size_t CS = counter_stride();
    const size_t cnt_bytes = sizeof(uint32_t) * (CNT_QUEUE0 + CS * ((6 + (K > 1 ? K : 0)) * ((size_t)ds.depth + 2)));
    const int sms = sm_count(ctx->device);
    const unsigned shade_grid = (unsigned)sms * (NP > 1 ? dual_shade() : NRCU_SHADE_MINB), big_grid = (unsigned)sms * (NP > 1 ? dual_big() : 8);
    const bool want_stats = stats != nullptr, timing = want_stats && params && (params->flags & NRCU_FLAG_KERNEL_TIMES), gate = ctx->mode == NRCU_MODE_ACC, bvh = ds.root_ref != NRCU_REF_EMPTY;
    size_t ev_i = 0;
    struct Span { size_t a, b; int kind; };
    std::vector<Span> spans;
    for (int p = 0; p < NP; p++) CTX_CUDA(cudaMemsetAsync(ctx->ws[p].counters.p, 0, sizeof(uint32_t) * CNT_QUEUE0, S[0]));
    if (want_stats) CTX_CUDA(cudaEventRecord(ctx->ev_begin, S[0]));
    if (NP > 1) { CTX_CUDA(cudaEventRecord(ctx->ev_fork, S[0])); for (int p = 1; p < NP; p++) CTX_CUDA(cudaStreamWaitEvent(S[p], ctx->ev_fork, 0)); }
    const uint64_t launches0 = ctx->launches;
    bool acc_recorded = false;

    struct Pipe {
        bool live; uint32_t w0, kw, n_slots;
        PathQueue q[2], qs; uint32_t* cnt; uint32_t *d_qn, *d_fetch, *d_nsurv, *d_nshadow, *d_sfetch, *d_snsurv, *d_qr;
        float2* hb; uint32_t* surv; f4* L; cudaStream_t st; unsigned long long* d_rays;
    } P[NRCU_MAX_WAVES];
    for (int p = 0; p < NP; p++) {
        nrcu_ctx::WaveSet& w = ctx->ws[p];
        Pipe& pp = P[p];
        pp.live = false; pp.st = S[p];
        for (int j = 0; j < 2; j++) pp.q[j] = PathQueue{w.qa[j].as<f4>(), w.qb[j].as<float2>(), w.qc[j].as<f4>(), glass_branch ? w.qd[j].as<uint32_t>() : nullptr};
        pp.qs = nee ? PathQueue{w.sa.as<f4>(), w.sb.as<float2>(), w.sc.as<f4>(), w.sd.as<uint32_t>()} : PathQueue{nullptr, nullptr, nullptr, nullptr};
        pp.cnt = w.counters.as<uint32_t>();
        pp.d_rays = reinterpret_cast<unsigned long long*>(pp.cnt + CNT_RAYS);
        // counter (kind, bounce d) sits CS * (kind * (depth + 2) + d) words into the block
        pp.d_qn = pp.cnt + CNT_QUEUE0;                               // queue size entering bounce d
        pp.d_fetch = pp.cnt + CNT_QUEUE0 + CS * (ds.depth + 2);      // work-fetch cursor of bounce d
        pp.d_nsurv = pp.cnt + CNT_QUEUE0 + CS * 2 * (ds.depth + 2);  // stage-1 survivors of bounce d
        pp.d_nshadow = pp.cnt + CNT_QUEUE0 + CS * 3 * (ds.depth + 2);   // NEE: shadow rays of bounce d, their fetch cursors and survivors
        pp.d_sfetch = pp.cnt + CNT_QUEUE0 + CS * 4 * (ds.depth + 2);
        pp.d_snsurv = pp.cnt + CNT_QUEUE0 + CS * 5 * (ds.depth + 2);
        pp.d_qr = pp.cnt + CNT_QUEUE0 + CS * 6 * (ds.depth + 2);       // K > 1: region r of the queue entering bounce d counts at d_qr[CS (K d + r)]
        pp.hb = w.hits.as<float2>(); pp.surv = w.surv.as<uint32_t>(); pp.L = w.L.as<f4>();
    }

    uint32_t wave_retries = 0, bounce_rounds = 0;
    for (uint32_t g0 = s0; g0 < s1;) {
        // ---- camera rays (+ stage 1 of their closest hit) of the waves of this group ------------------------------
        for (int p = 0; p < NP; p++) {
            Pipe& pp = P[p];
            pp.w0 = g0 + (uint32_t)p * k; pp.live = pp.w0 < s1;
            if (!pp.live) continue;
            pp.kw = std::min(k, s1 - pp.w0); pp.n_slots = pp.kw * nlive;
            cudaStream_t st = pp.st;
            CTX_CUDA(cudaMemsetAsync(pp.cnt + CNT_QUEUE0, 0, cnt_bytes - sizeof(uint32_t) * CNT_QUEUE0, st));
            const unsigned gen_grid = std::min<unsigned>(grid_for(pp.n_slots, 256), (unsigned)sms * 8);
            if (timing) cudaEventRecord(pool_event(ctx, ev_i), st);
            // rays are counted from the queue sizes at the end of the wave (k_accumulate): no ray counter for the kernels
            if (gate) k_raygen<true, true><<<gen_grid, 256, 0, st>>>(ds, seed, pp.w0, pp.n_slots, pp.q[0], pp.L, pp.d_qn, pp.hb, pp.surv, pp.d_nsurv, NRCU_RC_K);
            else k_raygen<false, true><<<gen_grid, 256, 0, st>>>(ds, seed, pp.w0, pp.n_slots, pp.q[0], pp.L, pp.d_qn, pp.hb, pp.surv, pp.d_nsurv, NRCU_RC_K);
            CTX_LAUNCH_CHECK("k_raygen");
            if (ds.dead_env && ds.depth > 0 && pp.kw * (npix - nlive) > 0) {   // environment map: the dead pixels' samples are the map along the camera ray
                const uint32_t n_entries = pp.kw * (npix - nlive);
                k_env_dead<<<std::min<unsigned>(grid_for(n_entries, 256), (unsigned)sms * 8), 256, 0, st>>>(ds, seed, pp.w0, n_entries, pp.L);
                CTX_LAUNCH_CHECK("k_env_dead");
            }
            if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 1), st); spans.push_back({ev_i, ev_i + 1, 0}); ev_i += 2; }   // booked as closest-hit time: stage 1 of bounce 0 is most of this kernel
        }
        // ---- bounces: stage 1 (k_big; bounce 0's ran inside k_raygen), stage 2 (k_trace*) on the survivors, k_shade ---
        for (uint32_t d = 0; d < ds.depth; d++) {
            for (int p = 0; p < NP; p++) {
                Pipe& pp = P[p];
                if (!pp.live) continue;
                cudaStream_t st = pp.st;
                PathQueue qi = pp.q[d & 1], qo = pp.q[(d + 1) & 1];
                // the queue entering bounce 0 is dense (one counter, written by k_raygen); later queues come in K regions
                const uint32_t max_blocks = (capacity + slack + 31u) / 32u;   // the arrays are padded to whole blocks (ensure_wave)
                const QRegions rin = (K > 1 && d > 0) ? QRegions{pp.d_qr + CS * K * d, logk, (uint32_t)CS, max_blocks} : QRegions{pp.d_qn + CS * d, 0u, (uint32_t)CS, max_blocks};
                uint32_t* const cnt_out = K > 1 ? pp.d_qr + CS * K * (d + 1) : pp.d_qn + CS * (d + 1);
                if (timing) cudaEventRecord(pool_event(ctx, ev_i), st);
                if (d > 0) {
                    if (big_balanced() == 2 && NRCU_OPT_RAYCOUNT) {
                        const unsigned g64 = (unsigned)sms * dual_big64();
                        if (gate) k_big_balanced64<true><<<g64, 32 * NRCU_BIG64_WARPS, 0, st>>>(ds, qi, rin, pp.hb, pp.surv, pp.d_nsurv + CS * d);
                        else k_big_balanced64<false><<<g64, 32 * NRCU_BIG64_WARPS, 0, st>>>(ds, qi, rin, pp.hb, pp.surv, pp.d_nsurv + CS * d);
                    }
                    else if (big_balanced()) {
                        if (gate) k_big_balanced<true, false><<<big_grid, 32 * NRCU_BIGB_WARPS, 0, st>>>(ds, qi, rin.cnt, pp.hb, pp.surv, pp.d_nsurv + CS * d, NRCU_RC_K, 0u, rin.logk, rin.cs, rin.max_blocks);
                        else k_big_balanced<false, false><<<big_grid, 32 * NRCU_BIGB_WARPS, 0, st>>>(ds, qi, rin.cnt, pp.hb, pp.surv, pp.d_nsurv + CS * d, NRCU_RC_K, 0u, rin.logk, rin.cs, rin.max_blocks);
                    }
                    else if (gate) k_big<true><<<big_grid, 256, 0, st>>>(ds, qi, rin, pp.hb, pp.surv, pp.d_nsurv + CS * d, NRCU_RC_K);
                    else k_big<false><<<big_grid, 256, 0, st>>>(ds, qi, rin, pp.hb, pp.surv, pp.d_nsurv + CS * d, NRCU_RC_K);
                    CTX_LAUNCH_CHECK("k_big");
                }
                if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 3), st); spans.push_back({ev_i, ev_i + 3, 0}); }
                if (bvh) {
                    if (gate) launch_stage2<true>(ctx, st, share, ds, qi, pp.hb, pp.surv, pp.d_nsurv + CS * d, pp.d_fetch + CS * d, pp.d_rays, d);
                    else launch_stage2<false>(ctx, st, share, ds, qi, pp.hb, pp.surv, pp.d_nsurv + CS * d, pp.d_fetch + CS * d, pp.d_rays, d);
                    CTX_LAUNCH_CHECK("k_trace");
                }
                if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 1), st); spans.push_back({ev_i + 3, ev_i + 1, 2}); }
#define NRCU_SHADE(N, B) k_shade<N, B><<<shade_grid, 256, 0, st>>>(ds, seed, d, glass_branch, pp.w0, qi, rin, pp.hb, qo, cnt_out, logk, capacity + slack, pp.L, pp.qs, pp.d_nshadow + CS * d)
                if (shade_pool() && !nee && !glass_branch)
                    k_shade_pool<<<shade_grid, 256, 0, st>>>(ds, seed, d, pp.w0, qi, rin, pp.hb, qo, cnt_out, logk, capacity + slack, pp.L);
                else
#if NRCU_OPT_BRANCH_TEMPLATE
                if (glass_branch) { if (nee) NRCU_SHADE(true, true); else NRCU_SHADE(false, true); }
                else { if (nee) NRCU_SHADE(true, false); else NRCU_SHADE(false, false); }
#else
                if (nee) NRCU_SHADE(true, false); else NRCU_SHADE(false, false);
#endif
#undef NRCU_SHADE
                CTX_LAUNCH_CHECK("k_shade");
                if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 2), st); spans.push_back({ev_i + 1, ev_i + 2, 1}); ev_i += 4; }
                if (glass_branch) {   // only the branching mode can outgrow the queue
                    k_clamp_count<<<1, 1, 0, st>>>(pp.d_qn + CS * (d + 1), capacity, cnt0 + CNT_HIGH_WATER, ds.overflow + 1);
                    CTX_LAUNCH_CHECK("k_clamp_count");
                    if (nee) { k_clamp_count<<<1, 1, 0, st>>>(pp.d_nshadow + CS * d, capacity, cnt0 + CNT_HIGH_WATER, ds.overflow + 1); CTX_LAUNCH_CHECK("k_clamp_count"); }
                }
                if (nee && d + 1 < ds.depth) {   // the shadow rays of this bounce: same closest-hit kernels, then visibility + add
                    if (timing) cudaEventRecord(pool_event(ctx, ev_i), st);
                    int nl = 0;
                    if (gate) launch_closest_hit<true>(ctx, st, share, ds, pp.qs, pp.d_nshadow + CS * d, pp.hb, pp.surv, pp.d_snsurv + CS * d, pp.d_sfetch + CS * d, NRCU_RC_K, &nl);
                    else launch_closest_hit<false>(ctx, st, share, ds, pp.qs, pp.d_nshadow + CS * d, pp.hb, pp.surv, pp.d_snsurv + CS * d, pp.d_sfetch + CS * d, NRCU_RC_K, &nl);
                    ctx->launches += nl - 1;
                    CTX_LAUNCH_CHECK("k_big/k_trace (shadow)");
                    if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 1), st); spans.push_back({ev_i, ev_i + 1, 0}); }
                    k_shadow_resolve<<<shade_grid, 256, 0, st>>>(ds, pp.qs, pp.d_nshadow + CS * d, pp.hb, pp.L, glass_branch);
                    CTX_LAUNCH_CHECK("k_shadow_resolve");
                    if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 2), st); spans.push_back({ev_i + 1, ev_i + 2, 1}); ev_i += 3; }
                }
            }
        }
        bounce_rounds += ds.depth;
        if (glass_branch) {
            // Only the branching mode can outgrow its queue (2^bounces rays per path inside glass).  A wave that dropped
            // rays is NOT accumulated: it is rendered again with half the samples (the queue capacity stays, so the room
            // per sample doubles); at one sample per wave the call fails with NRCU_ERR_OVERFLOW.
            uint32_t h_over = 0;
            CTX_CUDA(cudaMemcpyAsync(&h_over, ds.overflow + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, P[0].st));
            CTX_CUDA(cudaStreamSynchronize(P[0].st));
            if (h_over) {
                if (k > 1) {
                    CTX_CUDA(cudaMemsetAsync(ds.overflow + 1, 0, sizeof(uint32_t), P[0].st));
                    k = std::max<uint32_t>(1, k / 2); wave_retries++;
                    continue;
                }
                CTX_CUDA(cudaStreamSynchronize(S[0]));
                return check_overflow(ctx);
            }
        }
        // ---- accum += the waves' samples, in sample order whatever stream a wave ran on -----------------------------
        for (int p = 0; p < NP; p++) {
            Pipe& pp = P[p];
            if (!pp.live) continue;
            if (NP > 1 && acc_recorded) CTX_CUDA(cudaStreamWaitEvent(pp.st, ctx->ev_acc, 0));
            k_accumulate<<<grid_for(npix, 256), 256, 0, pp.st>>>(pp.L, d_accum, npix, pp.kw, pp.kw, pp.d_qn, pp.d_qr, K, nee ? pp.d_nshadow : nullptr, (uint32_t)CS, ds.depth, NRCU_RC_A, ds.dead_env ? nullptr : ds.live_flag, npix - nlive);
            CTX_LAUNCH_CHECK("k_accumulate");
            if (NP > 1) { CTX_CUDA(cudaEventRecord(ctx->ev_acc, pp.st)); acc_recorded = true; }
        }
        g0 += k * (uint32_t)NP;
    }
    for (int p = 1; p < NP; p++) { CTX_CUDA(cudaEventRecord(ctx->ev_join, S[p])); CTX_CUDA(cudaStreamWaitEvent(S[0], ctx->ev_join, 0)); }
    if (want_stats) {
        cudaStream_t st = S[0];
        CTX_CUDA(cudaEventRecord(ctx->ev_end, st));
        uint32_t h_cnt[4], h_wave[NRCU_MAX_WAVES][4];
        for (int p = 0; p < NP; p++) CTX_CUDA(cudaMemcpyAsync(h_wave[p], ctx->ws[p].counters.p, sizeof(h_wave[p]), cudaMemcpyDeviceToHost, st));
        CTX_CUDA(cudaStreamSynchronize(st));
        std::memcpy(h_cnt, h_wave[0], sizeof(h_cnt));
        std::memset(stats, 0, sizeof(*stats));
        CTX_CUDA(cudaEventElapsedTime(&stats->ms_total, ctx->ev_begin, ctx->ev_end));
        for (auto& sp : spans) {   // with two streams the spans of the two waves overlap: the sums exceed ms_total
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev_pool[sp.a], ctx->ev_pool[sp.b]);
            if (sp.kind == 0) stats->ms_trace += ms; else if (sp.kind == 2) { stats->ms_trace += ms; stats->ms_stage2 += ms; } else stats->ms_shade += ms;
        }
        unsigned long long rays = 0;
        for (int p = 0; p < NP; p++) { unsigned long long r; std::memcpy(&r, h_wave[p] + CNT_RAYS, 8); rays += r; }
        stats->rays = rays; stats->paths = (uint64_t)npix * (s1 - s0);
        stats->kernel_launches = ctx->launches - launches0;
        stats->ms_setup = ctx->ms_setup; stats->bvh_nodes = ctx->bvh_nodes; stats->n_primitives = ds.n_prims;
        stats->max_queue = glass_branch ? h_cnt[CNT_HIGH_WATER] : std::min<uint32_t>(qslots, nlive * (s1 - s0));   // without branching the bounce-0 queue is the largest
        stats->scheduler = NRCU_SCHED_WAVES; stats->iterations = bounce_rounds; stats->wave_retries = wave_retries;
        stats->dead_pixels = ds.depth ? npix - nlive : 0u;
        return check_overflow(ctx);
    }
    return NRCU_OK;
}

// ---------------------------------------------------------------------------------------------
// Path-regeneration scheduler (kernels: k_regen_init, k_big_balanced<.., SLOTS>, k_trace2, k_shade_regen, k_accumulate_lanes)
// ---------------------------------------------------------------------------------------------
#define NRCU_REGEN_BATCH 8      // iterations enqueued between two looks at the alive flags
#define NRCU_REGEN_RING 4       // batches whose counters / flags are live at any time
static bool regen_fused() { static uint32_t v = env_u32("NRCU_REGEN_FUSED", 0); return v != 0; }
static unsigned fused_blocks() { static uint32_t v = env_u32("NRCU_FUSED_BLOCKS", NRCU_FUSED_MINB); return v ? v : 1; }
static uint32_t regen_slots_target() { static uint32_t v = env_u32("NRCU_REGEN_MSLOTS", 64) << 20; return v; }

static int render_regen(nrcu_ctx* ctx, const nrcu_render_params* params, f4* d_accum, nrcu_stats* stats, uint32_t s0, uint32_t s1) {
    ctx->ds.nee = 0;
    const DScene& ds = ctx->ds;
    const uint32_t npix = ds.width * ds.height, n_samples = s1 - s0;
    const uint64_t seed = params ? params->seed : 0;
    // K slots per pixel.  Automatic: as many as NRCU_REGEN_MSLOTS (default 64 Mi) slots allow; fewer slots per pixel mean
    // more iterations per frame, i.e. a shorter drain phase relative to the frame.
    uint32_t K = (params && params->scheduler == NRCU_SCHED_REGEN && params->samples_per_wave) ? params->samples_per_wave
                                                                                              : std::max<uint32_t>(1, regen_slots_target() / npix);
    K = std::min<uint32_t>(std::min<uint32_t>(K, n_samples), 32768u);
    if ((uint64_t)K * npix > 0x7fffffffull) K = std::max<uint32_t>(1, (uint32_t)(0x7fffffffull / npix));
    int NP = std::max(1, std::min<int>(concurrent_waves(), (int)K));
    const size_t CS = counter_stride();
    const size_t blk = 2 * CS + (size_t)NRCU_REGEN_FLAGS * NRCU_REGEN_FLAG_STRIDE;   // words per iteration: survivors, fetch cursor, alive flags
    const size_t ring_iters = (size_t)NRCU_REGEN_BATCH * NRCU_REGEN_RING;
    struct Part { uint32_t lane0, lanes, n_slots; PathQueue q; float2* hits; uint32_t* surv; f4* lacc; uint32_t* cnt; cudaStream_t st; bool done; uint32_t batches; } P[NRCU_MAX_WAVES];
    int rc = NRCU_OK;
    for (;;) {   // shrink K instead of failing when the device is smaller or fuller than a B200
        rc = NRCU_OK;
        for (int p = 0; p < NP && rc == NRCU_OK; p++) {
            nrcu_ctx::WaveSet& w = ctx->ws[p];
            P[p].lane0 = (uint32_t)((uint64_t)p * K / NP); P[p].lanes = (uint32_t)((uint64_t)(p + 1) * K / NP) - P[p].lane0;
            P[p].n_slots = P[p].lanes * npix;
            const size_t S = (size_t)wave_skew_kb() << 10;
            DevBuf* bufs[] = {&w.qa[0], &w.qb[0], &w.qc[0], &w.hits, &w.surv, &w.L};
            for (size_t j = 0; j < sizeof(bufs) / sizeof(bufs[0]); j++) bufs[j]->skew = (j + 1 + 9 * (size_t)p) * S;
            auto need = [&](DevBuf& b, size_t bytes) { if (rc == NRCU_OK && b.ensure(bytes) != cudaSuccess) { ctx->error = "device allocation failed: out of memory"; rc = NRCU_ERR_CUDA; } };
            need(w.qa[0], sizeof(f4) * (size_t)P[p].n_slots); need(w.qb[0], sizeof(float2) * (size_t)P[p].n_slots); need(w.qc[0], sizeof(f4) * (size_t)P[p].n_slots);
            need(w.hits, sizeof(float2) * (size_t)P[p].n_slots); need(w.surv, sizeof(uint32_t) * (size_t)P[p].n_slots); need(w.L, sizeof(f4) * (size_t)P[p].n_slots);
            need(w.counters, sizeof(uint32_t) * (CNT_QUEUE0 + blk * ring_iters));
        }
        if (rc == NRCU_OK) break;
        cudaGetLastError();
        if (K == 1) return rc;
        for (int p = 0; p < NP; p++) { nrcu_ctx::WaveSet& w = ctx->ws[p]; for (DevBuf* b : {&w.qa[0], &w.qb[0], &w.qc[0], &w.hits, &w.surv, &w.L}) b->release(); }
        K = std::max<uint32_t>(1, K / 2);
        NP = std::max(1, std::min<int>(NP, (int)K));
    }
    const uint32_t samples_per_lane = (n_samples + K - 1) / K;
    const bool fused = regen_fused();
    uint64_t t_max = (uint64_t)samples_per_lane * ds.depth + (fused ? 1 : 0);   // a lane renders its samples one after the other, <= depth rays each
    { static uint32_t dbg = env_u32("NRCU_REGEN_MAX_ITERS", 0); if (dbg) t_max = std::min<uint64_t>(t_max, dbg); }   // steady-state measurements only: truncates the frame
    if (!ctx->h_flags) CTX_CUDA(cudaHostAlloc(&ctx->h_flags, sizeof(uint32_t) * NRCU_MAX_WAVES * NRCU_REGEN_RING * NRCU_REGEN_FLAGS * NRCU_REGEN_FLAG_STRIDE, cudaHostAllocDefault));
    cudaStream_t S[NRCU_MAX_WAVES] = {ctx->stream, ctx->stream, ctx->stream, ctx->stream};
    for (int p = 1; p < NP; p++) {
        if (!ctx->extra_stream[p]) CTX_CUDA(cudaStreamCreateWithFlags(&ctx->extra_stream[p], cudaStreamNonBlocking));
        S[p] = ctx->extra_stream[p];
    }
    if (NP > 1 && !ctx->ev_fork) {
        CTX_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)); CTX_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        CTX_CUDA(cudaEventCreateWithFlags(&ctx->ev_acc, cudaEventDisableTiming));
    }
    for (int p = 0; p < NP; p++) for (int r = 0; r < NRCU_REGEN_RING; r++)
        if (!ctx->ev_batch[p][r]) CTX_CUDA(cudaEventCreateWithFlags(&ctx->ev_batch[p][r], cudaEventDisableTiming));
    const int sms = sm_count(ctx->device);
    const unsigned share = (unsigned)NP;
    const unsigned big_grid = (unsigned)sms * (NP > 1 ? dual_big() : 8);
    const bool want_stats = stats != nullptr, timing = want_stats && params && (params->flags & NRCU_FLAG_KERNEL_TIMES), gate = ctx->mode == NRCU_MODE_ACC, bvh = ds.root_ref != NRCU_REF_EMPTY;
    size_t ev_i = 0;
    struct Span { size_t a, b; int kind; };
    std::vector<Span> spans;
    CTX_CUDA(cudaMemsetAsync(ctx->ws[0].counters.p, 0, sizeof(uint32_t) * CNT_QUEUE0, S[0]));
    if (want_stats) CTX_CUDA(cudaEventRecord(ctx->ev_begin, S[0]));
    if (NP > 1) { CTX_CUDA(cudaEventRecord(ctx->ev_fork, S[0])); for (int p = 1; p < NP; p++) CTX_CUDA(cudaStreamWaitEvent(S[p], ctx->ev_fork, 0)); }
    const uint64_t launches0 = ctx->launches;
    for (int p = 0; p < NP; p++) {
        nrcu_ctx::WaveSet& w = ctx->ws[p];
        P[p].q = PathQueue{w.qa[0].as<f4>(), w.qb[0].as<float2>(), w.qc[0].as<f4>(), nullptr};
        P[p].hits = w.hits.as<float2>(); P[p].surv = w.surv.as<uint32_t>(); P[p].lacc = w.L.as<f4>(); P[p].cnt = w.counters.as<uint32_t>();
        P[p].st = S[p]; P[p].done = false; P[p].batches = 0;
    }
    const size_t flag_words = (size_t)NRCU_REGEN_FLAGS * NRCU_REGEN_FLAG_STRIDE;
    auto iter_block = [&](const Part& pp, uint64_t t) { return pp.cnt + CNT_QUEUE0 + blk * (size_t)(t % ring_iters); };
    uint32_t iterations = 0;
    const dim3 shade_block(256);
    // Batches of NRCU_REGEN_BATCH iterations are enqueued two ahead of the last batch whose alive flags the host has seen:
    // the GPU never waits for the host, and at most two batches of (immediately returning) kernels are launched in vain.
    for (uint64_t b = 0;; b++) {
        bool any = false;
        for (int p = 0; p < NP; p++) {
            Part& pp = P[p];
            if (pp.done) continue;
            if (b >= 2) {   // look at batch b-2
                const int r = (int)((b - 2) % NRCU_REGEN_RING);
                CTX_CUDA(cudaEventSynchronize(ctx->ev_batch[p][r]));
                const uint32_t* hf = ctx->h_flags + ((size_t)p * NRCU_REGEN_RING + r) * flag_words;
                uint32_t alive = 0;
                for (int f = 0; f < NRCU_REGEN_FLAGS; f++) alive |= hf[(size_t)f * NRCU_REGEN_FLAG_STRIDE];
                if (!alive) { pp.done = true; continue; }
            }
            const uint64_t t_first = b * NRCU_REGEN_BATCH;
            if (t_first >= t_max) { pp.done = true; continue; }
            any = true;
            cudaStream_t st = pp.st;
            CTX_CUDA(cudaMemsetAsync(iter_block(pp, t_first), 0, sizeof(uint32_t) * blk * NRCU_REGEN_BATCH, st));
            const uint64_t t_last = std::min<uint64_t>(t_first + NRCU_REGEN_BATCH, t_max) - 1;
            for (uint64_t t = t_first; t <= t_last; t++) {
                uint32_t* ib = iter_block(pp, t);
                uint32_t *n_surv = ib, *fetch = ib + CS, *flags = ib + 2 * CS;
                const uint32_t* flags_prev = t ? iter_block(pp, t - 1) + 2 * CS : nullptr;
                if (fused && t > 0) {   // NRCU_REGEN_FUSED: shade (hits of iteration t-1) + stage 1 of the new rays in one kernel, then stage 2
                    if (timing) cudaEventRecord(pool_event(ctx, ev_i), st);
                    const unsigned fgrid = (unsigned)sms * fused_blocks();
                    if (gate) k_regen_fused<true><<<fgrid, 32 * NRCU_BIGB_WARPS, 0, st>>>(ds, seed, s0, n_samples, K, pp.lane0, pp.n_slots, pp.q, pp.hits, pp.lacc, pp.surv, n_surv, flags_prev, flags);
                    else k_regen_fused<false><<<fgrid, 32 * NRCU_BIGB_WARPS, 0, st>>>(ds, seed, s0, n_samples, K, pp.lane0, pp.n_slots, pp.q, pp.hits, pp.lacc, pp.surv, n_surv, flags_prev, flags);
                    CTX_LAUNCH_CHECK("k_regen_fused");
                    if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 1), st); spans.push_back({ev_i, ev_i + 1, 1}); }
                    if (bvh) {
                        if (gate) launch_stage2<true>(ctx, st, share, ds, pp.q, pp.hits, pp.surv, n_surv, fetch, nullptr);
                        else launch_stage2<false>(ctx, st, share, ds, pp.q, pp.hits, pp.surv, n_surv, fetch, nullptr);
                        CTX_LAUNCH_CHECK("k_trace");
                    }
                    if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 2), st); spans.push_back({ev_i + 1, ev_i + 2, 2}); ev_i += 3; }
                    if (p == 0) iterations++;
                    continue;
                }
                if (timing) cudaEventRecord(pool_event(ctx, ev_i), st);
                if (t == 0) {
                    const unsigned gen_grid = std::min<unsigned>(grid_for(pp.n_slots, 256), (unsigned)sms * 8);
                    if (gate) k_regen_init<true><<<gen_grid, 256, 0, st>>>(ds, seed, s0, n_samples, pp.lane0, pp.n_slots, pp.q, pp.lacc, pp.hits, pp.surv, n_surv);
                    else k_regen_init<false><<<gen_grid, 256, 0, st>>>(ds, seed, s0, n_samples, pp.lane0, pp.n_slots, pp.q, pp.lacc, pp.hits, pp.surv, n_surv);
                    CTX_LAUNCH_CHECK("k_regen_init");
                } else {
                    if (gate) k_big_balanced<true, true><<<big_grid, 32 * NRCU_BIGB_WARPS, 0, st>>>(ds, pp.q, flags_prev, pp.hits, pp.surv, n_surv, nullptr, pp.n_slots, 0u, 1u, 0xffffffffu);
                    else k_big_balanced<false, true><<<big_grid, 32 * NRCU_BIGB_WARPS, 0, st>>>(ds, pp.q, flags_prev, pp.hits, pp.surv, n_surv, nullptr, pp.n_slots, 0u, 1u, 0xffffffffu);
                    CTX_LAUNCH_CHECK("k_big_balanced (slots)");
                }
                if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 1), st); spans.push_back({ev_i, ev_i + 1, 0}); }
                if (bvh) {
                    if (gate) launch_stage2<true>(ctx, st, share, ds, pp.q, pp.hits, pp.surv, n_surv, fetch, nullptr);
                    else launch_stage2<false>(ctx, st, share, ds, pp.q, pp.hits, pp.surv, n_surv, fetch, nullptr);
                    CTX_LAUNCH_CHECK("k_trace");
                }
                if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 2), st); spans.push_back({ev_i + 1, ev_i + 2, 2}); }
                if (fused) {   // iteration 0 of the fused form: first rays + stage 1 + stage 2 only; everybody is alive
                    CTX_CUDA(cudaMemsetAsync(flags, 0xff, sizeof(uint32_t), st));
                    if (timing) ev_i += 4;
                    if (p == 0) iterations++;
                    continue;
                }
                k_shade_regen<<<dim3(grid_for(npix, 256), pp.lanes), shade_block, 0, st>>>(ds, seed, s0, n_samples, K, pp.lane0, pp.q, pp.hits, pp.lacc, flags_prev, flags);
                CTX_LAUNCH_CHECK("k_shade_regen");
                if (timing) { cudaEventRecord(pool_event(ctx, ev_i + 3), st); spans.push_back({ev_i + 2, ev_i + 3, 1}); ev_i += 4; }
                if (p == 0) iterations++;
            }
            const int r = (int)(b % NRCU_REGEN_RING);
            CTX_CUDA(cudaMemcpyAsync(ctx->h_flags + ((size_t)p * NRCU_REGEN_RING + r) * flag_words, iter_block(pp, t_last) + 2 * CS, sizeof(uint32_t) * flag_words, cudaMemcpyDeviceToHost, st));
            CTX_CUDA(cudaEventRecord(ctx->ev_batch[p][r], st));
        }
        if (!any) break;
    }
    for (int p = 1; p < NP; p++) { CTX_CUDA(cudaEventRecord(ctx->ev_join, S[p])); CTX_CUDA(cudaStreamWaitEvent(S[0], ctx->ev_join, 0)); }
    unsigned long long* d_rays = reinterpret_cast<unsigned long long*>(ctx->ws[0].counters.as<uint32_t>() + CNT_RAYS);
    for (int p = 0; p < NP; p++) {   // lane order = partition order: the summation tree does not depend on NRCU_WAVES
        k_accumulate_lanes<<<grid_for(npix, 256), 256, 0, S[0]>>>(P[p].lacc, d_accum, npix, P[p].lanes, p == 0 ? n_samples : 0u, d_rays);
        CTX_LAUNCH_CHECK("k_accumulate_lanes");
    }
    if (want_stats) {
        cudaStream_t st = S[0];
        CTX_CUDA(cudaEventRecord(ctx->ev_end, st));
        unsigned long long rays = 0;
        CTX_CUDA(cudaMemcpyAsync(&rays, d_rays, sizeof(rays), cudaMemcpyDeviceToHost, st));
        CTX_CUDA(cudaStreamSynchronize(st));
        std::memset(stats, 0, sizeof(*stats));
        CTX_CUDA(cudaEventElapsedTime(&stats->ms_total, ctx->ev_begin, ctx->ev_end));
        for (auto& sp : spans) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev_pool[sp.a], ctx->ev_pool[sp.b]);
            if (sp.kind == 0) stats->ms_trace += ms; else if (sp.kind == 2) { stats->ms_trace += ms; stats->ms_stage2 += ms; } else stats->ms_shade += ms;
        }
        stats->rays = rays; stats->paths = (uint64_t)npix * n_samples;
        stats->kernel_launches = ctx->launches - launches0;
        stats->ms_setup = ctx->ms_setup; stats->bvh_nodes = ctx->bvh_nodes; stats->n_primitives = ds.n_prims;
        stats->max_queue = K * npix; stats->scheduler = NRCU_SCHED_REGEN; stats->iterations = iterations;
        return check_overflow(ctx);
    }
    return NRCU_OK;
}

static int default_scheduler() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("NRCU_SCHED"); v = (e && std::string(e) == "waves") ? NRCU_SCHED_WAVES : ((e && std::string(e) == "regen") ? NRCU_SCHED_REGEN : NRCU_SCHED_DEFAULT); }
    return v;
}

static int render_pt(nrcu_ctx* ctx, const nrcu_render_params* params, f4* d_accum, nrcu_stats* stats) {
    uint32_t s0 = params ? params->sample_begin : 0, s1 = params ? params->sample_end : 0;
    if (s0 == 0 && s1 == 0) s1 = ctx->spp;
    if (s1 < s0) { ctx->error = "sample_end < sample_begin"; return NRCU_ERR_INVALID; }
    int sched = params ? (int)params->scheduler : NRCU_SCHED_AUTO;
    if (sched != NRCU_SCHED_WAVES && sched != NRCU_SCHED_REGEN) sched = (params && params->samples_per_wave) ? NRCU_SCHED_WAVES : default_scheduler();
    const bool nee = direct_sampling_mode(ctx, params) != 0;
    const bool branch = params && params->glass_mode == NRCU_GLASS_BRANCH;
    const uint32_t depth = ctx->ds.depth;
    // the regeneration scheduler carries one ray per slot: no shadow rays, no path splitting; its state packs the bounce
    // into 12 bits and counts a slot's rays in an fp32 (exact up to 2^24)
    const bool can_regen = !nee && !branch && depth >= 1 && depth < (1u << NRCU_SLOT_BOUNCE_BITS) && s1 > s0 &&
                           (uint64_t)(s1 - s0) * depth < (1ull << 24);
    if (sched == NRCU_SCHED_REGEN && can_regen) return render_regen(ctx, params, d_accum, stats, s0, s1);
    return render_waves(ctx, params, d_accum, stats);
}

int nrcu_render_accumulate(nrcu_ctx* ctx, const nrcu_render_params* params, float* d_accum, nrcu_stats* stats) {
    if (!ctx) return NRCU_ERR_INVALID;
    if (!ctx->have_scene) { ctx->error = "nrcu_render_accumulate: no scene uploaded"; return NRCU_ERR_STATE; }
    if (!d_accum) { ctx->error = "d_accum is null"; return NRCU_ERR_INVALID; }
    if (ctx->mode == NRCU_MODE_RAYCAST) { ctx->error = "RayCast mode has no sample accumulation; use nrcu_render"; return NRCU_ERR_STATE; }
    CTX_CUDA(cudaSetDevice(ctx->device));
    return render_pt(ctx, params, reinterpret_cast<f4*>(d_accum), stats);
}

int nrcu_resolve(nrcu_ctx* ctx, const float* d_accum, float* d_rgba) {
    if (!ctx || !d_accum || !d_rgba) return NRCU_ERR_INVALID;
    if (!ctx->have_scene) { ctx->error = "nrcu_resolve: no scene uploaded"; return NRCU_ERR_STATE; }
    CTX_CUDA(cudaSetDevice(ctx->device));
    const uint32_t npix = ctx->ds.width * ctx->ds.height;
    k_resolve<<<grid_for(npix, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const f4*>(d_accum), reinterpret_cast<f4*>(d_rgba), npix);
    CTX_LAUNCH_CHECK("k_resolve");
    return NRCU_OK;
}

int nrcu_render(nrcu_ctx* ctx, const nrcu_render_params* params, float* rgba_out, nrcu_stats* stats) {
    if (!ctx) return NRCU_ERR_INVALID;
    if (!ctx->have_scene) { ctx->error = "nrcu_render: no scene uploaded"; return NRCU_ERR_STATE; }
    if (!rgba_out) { ctx->error = "rgba_out is null"; return NRCU_ERR_INVALID; }
    CTX_CUDA(cudaSetDevice(ctx->device));
    const DScene& ds = ctx->ds;
    const uint32_t npix = ds.width * ds.height;
    cudaStream_t st = ctx->stream;
    CTX_CUDA(ctx->rgba_dev.ensure(sizeof(f4) * (size_t)npix));
    if (ctx->mode == NRCU_MODE_RAYCAST) {
        CTX_CUDA(ctx->ws[0].counters.ensure(sizeof(uint32_t) * 64));
        unsigned long long* d_rays = ctx->ws[0].counters.as<unsigned long long>();
        CTX_CUDA(cudaMemsetAsync(d_rays, 0, 8, st));
        uint64_t launches0 = ctx->launches;
        CTX_CUDA(cudaEventRecord(ctx->ev_begin, st));
        k_raycast<<<grid_for(npix, 128), 128, 0, st>>>(ds, ctx->rgba_dev.as<f4>(), d_rays);
        CTX_LAUNCH_CHECK("k_raycast");
        CTX_CUDA(cudaEventRecord(ctx->ev_end, st));
        CTX_CUDA(cudaMemcpyAsync(rgba_out, ctx->rgba_dev.p, sizeof(f4) * (size_t)npix, cudaMemcpyDeviceToHost, st));
        unsigned long long rays = 0;
        CTX_CUDA(cudaMemcpyAsync(&rays, d_rays, 8, cudaMemcpyDeviceToHost, st));
        CTX_CUDA(cudaStreamSynchronize(st));
        if (stats) {
            std::memset(stats, 0, sizeof(*stats));
            CTX_CUDA(cudaEventElapsedTime(&stats->ms_total, ctx->ev_begin, ctx->ev_end));
            stats->ms_trace = stats->ms_total; stats->rays = rays; stats->paths = npix;
            stats->kernel_launches = ctx->launches - launches0; stats->ms_setup = ctx->ms_setup; stats->n_primitives = ds.n_prims;
        }
        return NRCU_OK;   // brute force: no traversal stack, nothing that can overflow
    }
    CTX_CUDA(ctx->accum_own.ensure(sizeof(f4) * (size_t)npix));
    CTX_CUDA(cudaMemsetAsync(ctx->accum_own.p, 0, sizeof(f4) * (size_t)npix, st));
    nrcu_render_params p{};
    if (params) p = *params;
    p.sample_begin = 0; p.sample_end = 0;   // whole frame
    int rc = render_pt(ctx, &p, ctx->accum_own.as<f4>(), stats);
    if (rc != NRCU_OK) return rc;
    k_resolve<<<grid_for(npix, 256), 256, 0, st>>>(ctx->accum_own.as<f4>(), ctx->rgba_dev.as<f4>(), npix);
    CTX_LAUNCH_CHECK("k_resolve");
    if (stats) stats->kernel_launches++;
    CTX_CUDA(cudaMemcpyAsync(rgba_out, ctx->rgba_dev.p, sizeof(f4) * (size_t)npix, cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaStreamSynchronize(st));
    return check_overflow(ctx);
}

int nrcu_render_progressive(nrcu_ctx* ctx, const nrcu_render_params* params, uint32_t samples_per_update,
                            float* rgba_out, nrcu_update_fn on_update, void* user, nrcu_stats* stats) {
    if (!ctx) return NRCU_ERR_INVALID;
    if (!ctx->have_scene) { ctx->error = "nrcu_render_progressive: no scene uploaded"; return NRCU_ERR_STATE; }
    if (!rgba_out) { ctx->error = "rgba_out is null"; return NRCU_ERR_INVALID; }
    if (ctx->mode == NRCU_MODE_RAYCAST || ctx->spp == 0) {   // one deterministic pass: a single update
        int rc = nrcu_render(ctx, params, rgba_out, stats);
        if (rc == NRCU_OK && on_update) on_update(user, rgba_out, 1, 1);
        return rc;
    }
    CTX_CUDA(cudaSetDevice(ctx->device));
    const uint32_t npix = ctx->ds.width * ctx->ds.height, spp = ctx->spp;
    cudaStream_t st = ctx->stream;
    CTX_CUDA(ctx->rgba_dev.ensure(sizeof(f4) * (size_t)npix));
    CTX_CUDA(ctx->accum_own.ensure(sizeof(f4) * (size_t)npix));
    CTX_CUDA(cudaMemsetAsync(ctx->accum_own.p, 0, sizeof(f4) * (size_t)npix, st));
    uint32_t step = samples_per_update ? samples_per_update : std::max<uint32_t>(1, wave_slots_target() / npix);
    nrcu_stats total{}; std::memset(&total, 0, sizeof(total));
    for (uint32_t s0 = 0; s0 < spp; s0 += step) {
        nrcu_render_params p{};
        if (params) p = *params;
        p.sample_begin = s0; p.sample_end = std::min(spp, s0 + step);
        nrcu_stats part{};
        int rc = render_pt(ctx, &p, ctx->accum_own.as<f4>(), &part);
        if (rc != NRCU_OK) return rc;
        k_resolve<<<grid_for(npix, 256), 256, 0, st>>>(ctx->accum_own.as<f4>(), ctx->rgba_dev.as<f4>(), npix);
        CTX_LAUNCH_CHECK("k_resolve");
        CTX_CUDA(cudaMemcpyAsync(rgba_out, ctx->rgba_dev.p, sizeof(f4) * (size_t)npix, cudaMemcpyDeviceToHost, st));
        CTX_CUDA(cudaStreamSynchronize(st));
        total.paths += part.paths; total.rays += part.rays; total.kernel_launches += part.kernel_launches + 1;
        total.ms_total += part.ms_total; total.ms_trace += part.ms_trace; total.ms_shade += part.ms_shade; total.ms_stage2 += part.ms_stage2;
        total.max_queue = std::max(total.max_queue, part.max_queue);
        total.ms_setup = part.ms_setup; total.bvh_nodes = part.bvh_nodes; total.n_primitives = part.n_primitives;
        if (on_update && on_update(user, rgba_out, p.sample_end, spp) != 0) break;
    }
    if (stats) *stats = total;
    return NRCU_OK;
}

int nrcu_render_multi(nrcu_ctx* const* ctxs, int n_ctx, const nrcu_render_params* params, float* rgba_out, nrcu_stats* stats) {
    if (!ctxs || n_ctx < 1 || !ctxs[0]) return NRCU_ERR_INVALID;
    nrcu_ctx* ctx = ctxs[0];   // root: errors are reported here
    if (n_ctx == 1 || ctx->mode == NRCU_MODE_RAYCAST) return nrcu_render(ctx, params, rgba_out, stats);
    if (n_ctx > NRCU_MAX_DEVICES) { ctx->error = "nrcu_render_multi: too many contexts"; return NRCU_ERR_INVALID; }
    if (!rgba_out) { ctx->error = "rgba_out is null"; return NRCU_ERR_INVALID; }
    for (int g = 0; g < n_ctx; g++) {
        nrcu_ctx* c = ctxs[g];
        if (!c || !c->have_scene) { ctx->error = "nrcu_render_multi: a context has no scene"; return NRCU_ERR_STATE; }
        if (c->mode != ctx->mode || c->spp != ctx->spp || c->ds.width != ctx->ds.width || c->ds.height != ctx->ds.height || c->ds.n_prims != ctx->ds.n_prims) {
            ctx->error = "nrcu_render_multi: the contexts hold different scenes"; return NRCU_ERR_INVALID;
        }
        for (int h = 0; h < g; h++) if (ctxs[h]->device == c->device) { ctx->error = "nrcu_render_multi: two contexts on one device"; return NRCU_ERR_INVALID; }
    }
    const uint32_t npix = ctx->ds.width * ctx->ds.height, spp = ctx->spp;
    std::vector<int> rc(n_ctx, NRCU_OK);
    std::vector<nrcu_stats> st(n_ctx);
    std::vector<std::thread> pool;
    auto wall0 = std::chrono::steady_clock::now();
    for (int g = 0; g < n_ctx; g++) {
        pool.emplace_back([&, g]() {
            nrcu_ctx* c = ctxs[g];
            std::memset(&st[g], 0, sizeof(nrcu_stats));
            if (cudaSetDevice(c->device) != cudaSuccess) { c->error = "cudaSetDevice failed"; rc[g] = NRCU_ERR_CUDA; return; }
            if (c->accum_own.ensure(sizeof(f4) * (size_t)npix) != cudaSuccess ||
                cudaMemsetAsync(c->accum_own.p, 0, sizeof(f4) * (size_t)npix, c->stream) != cudaSuccess) { c->error = "accum allocation failed"; rc[g] = NRCU_ERR_CUDA; return; }
            nrcu_render_params p{};
            if (params) p = *params;
            p.sample_begin = (uint32_t)((uint64_t)g * spp / n_ctx); p.sample_end = (uint32_t)((uint64_t)(g + 1) * spp / n_ctx);
            if (p.sample_end > p.sample_begin) rc[g] = render_pt(c, &p, c->accum_own.as<f4>(), &st[g]);
            if (rc[g] == NRCU_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) { c->error = "synchronize failed"; rc[g] = NRCU_ERR_CUDA; }
        });
    }
    for (auto& t : pool) t.join();
    for (int g = 0; g < n_ctx; g++) if (rc[g] != NRCU_OK) { if (g) ctx->error = "device " + std::to_string(ctxs[g]->device) + ": " + ctxs[g]->error; return rc[g]; }
    // ---- reduce + resolve on the root, reading the peers' partial frames in place ------------------------------
    CTX_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s0 = ctx->stream;
    CTX_CUDA(ctx->rgba_dev.ensure(sizeof(f4) * (size_t)npix));
    PartialFrames pf{}; pf.n = n_ctx; pf.part[0] = ctx->accum_own.as<f4>();
    std::vector<DevBuf> staged(n_ctx);
    for (int g = 1; g < n_ctx; g++) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, ctx->device, ctxs[g]->device);
        if (can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[g]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            can = e == cudaSuccess;
            if (!can) cudaGetLastError();
        }
        if (can) pf.part[g] = ctxs[g]->accum_own.as<f4>();
        else {   // no peer mapping: stage the partial frame through a copy
            CTX_CUDA(staged[g].ensure(sizeof(f4) * (size_t)npix));
            cudaError_t e = cudaMemcpyPeerAsync(staged[g].p, ctx->device, ctxs[g]->accum_own.p, ctxs[g]->device, sizeof(f4) * (size_t)npix, s0);
            if (e != cudaSuccess) {
                ctx->error = "device " + std::to_string(ctxs[g]->device) + ": partial frame neither peer-mapped nor copied: " + cudaGetErrorString(e);
                cudaGetLastError();
                return NRCU_ERR_PEER;
            }
            pf.part[g] = staged[g].as<f4>();
        }
    }
    k_resolve_multi<<<grid_for(npix, 256), 256, 0, s0>>>(pf, nullptr, ctx->rgba_dev.as<f4>(), npix);
    CTX_LAUNCH_CHECK("k_resolve_multi");
    CTX_CUDA(cudaMemcpyAsync(rgba_out, ctx->rgba_dev.p, sizeof(f4) * (size_t)npix, cudaMemcpyDeviceToHost, s0));
    CTX_CUDA(cudaStreamSynchronize(s0));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (int g = 0; g < n_ctx; g++) {
            stats->paths += st[g].paths; stats->rays += st[g].rays; stats->kernel_launches += st[g].kernel_launches;
            stats->ms_trace = std::max(stats->ms_trace, st[g].ms_trace); stats->ms_shade = std::max(stats->ms_shade, st[g].ms_shade);
            stats->ms_stage2 = std::max(stats->ms_stage2, st[g].ms_stage2);
            stats->max_queue = std::max(stats->max_queue, st[g].max_queue);
        }
        stats->kernel_launches += 1;
        stats->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - wall0).count();
        stats->ms_setup = ctx->ms_setup; stats->bvh_nodes = ctx->bvh_nodes; stats->n_primitives = ctx->ds.n_prims;
    }
    return NRCU_OK;
}

int nrcu_render_mlt(nrcu_ctx* ctx, const nrcu_mlt_params* params, float* rgba_out, nrcu_stats* stats) {
    if (!ctx) return NRCU_ERR_INVALID;
    if (!ctx->have_scene) { ctx->error = "nrcu_render_mlt: no scene uploaded"; return NRCU_ERR_STATE; }
    if (!rgba_out) { ctx->error = "rgba_out is null"; return NRCU_ERR_INVALID; }
    if (ctx->mode == NRCU_MODE_RAYCAST) { ctx->error = "nrcu_render_mlt needs a scene uploaded in a path-tracing mode"; return NRCU_ERR_STATE; }
    const DScene& ds = ctx->ds;
    if (ds.depth > NRCU_MLT_MAX_DEPTH) { ctx->error = "nrcu_render_mlt: depth > 32"; return NRCU_ERR_INVALID; }
    CTX_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t npix = ds.width * ds.height;
    const uint64_t seed = params ? params->seed : 0;
    const uint32_t mpp = params && params->mutations_per_pixel ? params->mutations_per_pixel : ctx->spp;
    const uint64_t total = (uint64_t)mpp * npix;
    uint32_t chains = params && params->chains ? params->chains : (uint32_t)std::min<uint64_t>(1u << 18, std::max<uint64_t>(1, total / 64));
    chains = (uint32_t)std::min<uint64_t>(chains, std::max<uint64_t>(total, 1));
    const uint32_t per_chain = (uint32_t)std::min<uint64_t>(0x7fffffffu, (total + chains - 1) / std::max(chains, 1u));
    const uint32_t n_init = params && params->n_init ? params->n_init : (1u << 18);
    const float p_large = params && params->large_step_prob > 0.f ? std::min(params->large_step_prob, 1.f) : 0.3f;
    const int tone = params ? (int)params->tone_map : 0;
    CTX_CUDA(ctx->rgba_dev.ensure(sizeof(f4) * (size_t)npix));
    CTX_CUDA(ctx->accum_own.ensure(sizeof(f4) * (size_t)npix));
    CTX_CUDA(ctx->ws[0].counters.ensure(sizeof(uint32_t) * 64));
    unsigned long long* d_cnt = ctx->ws[0].counters.as<unsigned long long>();         // [0] rays, [1] accepted mutations
    DevBuf& scratch = ctx->ws[0].hits;                                                // b-estimation: n_init floats + n_init doubles
    CTX_CUDA(scratch.ensure(sizeof(double) * (size_t)n_init + sizeof(float) * (size_t)n_init + 16));
    double* d_cdf = scratch.as<double>();
    float* d_scalar = reinterpret_cast<float*>(d_cdf + n_init);
    CTX_CUDA(cudaMemsetAsync(ctx->ws[0].counters.p, 0, 32, st));
    CTX_CUDA(cudaMemsetAsync(ctx->accum_own.p, 0, sizeof(f4) * (size_t)npix, st));
    const uint64_t launches0 = ctx->launches;
    const bool gate = ctx->mode == NRCU_MODE_ACC;
    CTX_CUDA(cudaEventRecord(ctx->ev_begin, st));
    if (total > 0) {
        if (gate) k_mlt_b<true><<<grid_for(n_init, 128), 128, 0, st>>>(ds, seed, n_init, d_scalar, d_cnt);
        else k_mlt_b<false><<<grid_for(n_init, 128), 128, 0, st>>>(ds, seed, n_init, d_scalar, d_cnt);
        CTX_LAUNCH_CHECK("k_mlt_b");
        k_mlt_cdf<<<1, 1024, 0, st>>>(d_scalar, n_init, d_cdf);
        CTX_LAUNCH_CHECK("k_mlt_cdf");
        if (gate) k_mlt_chains<true><<<grid_for(chains, 128), 128, 0, st>>>(ds, seed, chains, per_chain, d_cdf, n_init, p_large, ctx->accum_own.as<f4>(), d_cnt);
        else k_mlt_chains<false><<<grid_for(chains, 128), 128, 0, st>>>(ds, seed, chains, per_chain, d_cdf, n_init, p_large, ctx->accum_own.as<f4>(), d_cnt);
        CTX_LAUNCH_CHECK("k_mlt_chains");
    }
    const double n_mut = (double)chains * per_chain;
    const float scale = n_mut > 0 ? (float)(((double)ds.width + 2.0) * ((double)ds.height + 2.0) / (4.0 * n_mut)) : 0.f;
    k_mlt_resolve<<<grid_for(npix, 256), 256, 0, st>>>(ctx->accum_own.as<f4>(), ctx->rgba_dev.as<f4>(), npix, scale, tone);
    CTX_LAUNCH_CHECK("k_mlt_resolve");
    CTX_CUDA(cudaEventRecord(ctx->ev_end, st));
    CTX_CUDA(cudaMemcpyAsync(rgba_out, ctx->rgba_dev.p, sizeof(f4) * (size_t)npix, cudaMemcpyDeviceToHost, st));
    unsigned long long h_cnt[2] = {0, 0};
    CTX_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        CTX_CUDA(cudaEventElapsedTime(&stats->ms_total, ctx->ev_begin, ctx->ev_end));
        stats->paths = total ? (uint64_t)n_mut : 0; stats->rays = h_cnt[0]; stats->kernel_launches = ctx->launches - launches0;
        stats->ms_trace = stats->ms_total; stats->ms_setup = ctx->ms_setup; stats->bvh_nodes = ctx->bvh_nodes; stats->n_primitives = ds.n_prims;
        stats->max_queue = chains; stats->iterations = per_chain; stats->wave_retries = (uint32_t)std::min<unsigned long long>(0xffffffffull, h_cnt[1] >> 10);
    }
    return check_overflow(ctx);
}

int nrcu_trace_batch(nrcu_ctx* ctx, const float* rays, uint32_t n, int32_t* prim_id, float* t) {
    if (!ctx) return NRCU_ERR_INVALID;
    if (!ctx->have_scene) { ctx->error = "nrcu_trace_batch: no scene uploaded"; return NRCU_ERR_STATE; }
    if (n == 0) return NRCU_OK;
    if (!rays) { ctx->error = "rays is null"; return NRCU_ERR_INVALID; }
    CTX_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevBuf d_rays, qa, qb, hits, surv, cnt;
    CTX_CUDA(d_rays.ensure(sizeof(float) * 6 * (size_t)n)); CTX_CUDA(qa.ensure(sizeof(f4) * (size_t)n)); CTX_CUDA(qb.ensure(sizeof(float2) * (size_t)n));
    CTX_CUDA(hits.ensure(sizeof(float2) * (size_t)n)); CTX_CUDA(surv.ensure(sizeof(uint32_t) * (size_t)n)); CTX_CUDA(cnt.ensure(32));
    CTX_CUDA(cudaMemcpyAsync(d_rays.p, rays, sizeof(float) * 6 * (size_t)n, cudaMemcpyHostToDevice, st));
    uint32_t h_cnt[8] = {0, 0, n, 0, 0, 0, 0, 0};   // [0..1] ray counter, [2] n, [3] fetch cursor, [4] survivors
    CTX_CUDA(cudaMemcpyAsync(cnt.p, h_cnt, sizeof(h_cnt), cudaMemcpyHostToDevice, st));
    PathQueue q{qa.as<f4>(), qb.as<float2>(), nullptr, nullptr};
    k_pack_rays<<<grid_for(n, 256), 256, 0, st>>>(d_rays.as<float>(), n, q);
    CTX_LAUNCH_CHECK("k_pack_rays");
    uint32_t* c = cnt.as<uint32_t>();
    int nl = 0;
    if (ctx->mode == NRCU_MODE_RAYCAST) { k_trace_linear_rc<<<grid_for(n, 128), 128, 0, st>>>(ctx->ds, q, n, hits.as<float2>()); nl = 1; }
    else if (ctx->mode == NRCU_MODE_ACC) launch_closest_hit<true>(ctx, st, 1, ctx->ds, q, c + 2, hits.as<float2>(), surv.as<uint32_t>(), c + 4, c + 3, reinterpret_cast<unsigned long long*>(c), &nl);
    else launch_closest_hit<false>(ctx, st, 1, ctx->ds, q, c + 2, hits.as<float2>(), surv.as<uint32_t>(), c + 4, c + 3, reinterpret_cast<unsigned long long*>(c), &nl);
    ctx->launches += nl - 1;
    CTX_LAUNCH_CHECK("k_big/k_trace");
    std::vector<float2> h(n);
    CTX_CUDA(cudaMemcpyAsync(h.data(), hits.p, sizeof(float2) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CTX_CUDA(cudaStreamSynchronize(st));
    if (int orc = check_overflow(ctx)) return orc;
    for (uint32_t i = 0; i < n; i++) {
        int id; std::memcpy(&id, &h[i].y, 4);
        if (prim_id) prim_id[i] = id;
        if (t) t[i] = h[i].x;
    }
    return NRCU_OK;
}

}  // extern "C"
