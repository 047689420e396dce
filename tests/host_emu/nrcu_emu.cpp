// nrcu_emu.cpp — CPU emulation of the device code, TEST INFRASTRUCTURE ONLY.
//
// The __host__ __device__ bodies that the CUDA kernels wrap (nrenderer_b200/csrc/nrcu_*.cuh:
// scene flattening, the binned-SAH BVH build steps, wide-BVH traversal, exact primitive tests,
// camera rays, shading, the per-bounce path update) are compiled here with g++ and run
// sequentially, one "thread" after another, so that `pytest -m "not gpu"` can check the kernel
// logic against the oracle on a machine without a GPU.  This library is never loaded by
// nrenderer_b200/ and is not a fallback: the product fails loudly without a CUDA device.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nrcu.h"
#include "nrcu_bvh.cuh"
#include "nrcu_prep.cuh"
#include "nrcu_shade.cuh"
#include "nrcu_host_prep.hpp"

using namespace nrcu;

struct EmuScene {
    DScene ds;
    HostPrep hp;
    uint32_t spp;
    std::vector<float> mesh_pos;
    std::vector<f4> geom, shade, box, nodes, env, leaf_geom, leaf_box, big_geom, big_box, big_bound, bound;
    std::vector<float> env_tab;
    std::vector<uint32_t> big_meta;
    int n_big = 0;
    std::vector<uint32_t> meta, leaf_prims;
    std::vector<float> export16;
    int n_binary_nodes = 0, n_wide = 0, levels = 0, max_leaf = 0;
    std::string error;
};

static void emu_build_bvh(EmuScene& es) {
    const uint32_t n = es.ds.n_prims;
    const int cap = 2 * (int)n + 2;
    const int bin_nodes = std::max(64, (int)(0.4 * n) + 8);
    std::vector<int> prim_node(n), nbox(6 * (size_t)cap), cbox(6 * (size_t)cap), ncount(cap), nidmin(cap), nidmax(cap), nstate(cap), nchild(cap),
        nsplit_axis(cap), ndepth(cap), nleaf_first(cap), nleaf_fill(cap), nwide(cap), nbin_slot(cap), counters(8, 0),
        bins((size_t)bin_nodes * 3 * NRCU_NBINS * NRCU_BIN_WORDS);
    std::vector<float> nsplit_pos(cap);
    es.leaf_prims.assign(n, 0);
    std::vector<f4> wide((size_t)std::max(1u, n) * NRCU_BVH_NODE_F4);
    BvhBuild b{};
    b.n_prims = n; b.prim_box = es.box.data(); b.prim_bound = es.bound.data(); b.prim_meta = es.meta.data();
    b.prim_node = prim_node.data(); b.nbox = nbox.data(); b.cbox = cbox.data(); b.ncount = ncount.data();
    b.nidmin = nidmin.data(); b.nidmax = nidmax.data(); b.nstate = nstate.data(); b.nchild = nchild.data();
    b.nsplit_axis = nsplit_axis.data(); b.nsplit_pos = nsplit_pos.data(); b.ndepth = ndepth.data();
    b.nleaf_first = nleaf_first.data(); b.nleaf_fill = nleaf_fill.data(); b.nwide = nwide.data();
    b.bins = bins.data(); b.bin_nodes = bin_nodes; b.counters = counters.data(); b.nbin_slot = nbin_slot.data();
    b.leaf_prims = es.leaf_prims.data(); b.wide_nodes = wide.data();
    es.leaf_geom.assign(3 * (size_t)std::max(1u, n), mk4(0, 0, 0, 0)); es.leaf_box.assign(2 * (size_t)std::max(1u, n), mk4(0, 0, 0, 0));
    b.prim_geom = es.geom.data(); b.leaf_geom = es.leaf_geom.data(); b.leaf_box = es.leaf_box.data();
    es.big_geom.assign(3 * NRCU_MAX_BIG, mk4(0, 0, 0, 0)); es.big_box.assign(2 * NRCU_MAX_BIG, mk4(0, 0, 0, 0)); es.big_bound.assign(2 * NRCU_MAX_BIG, mk4(0, 0, 0, 0)); es.big_meta.assign(NRCU_MAX_BIG, 0);
    b.big_geom = es.big_geom.data(); b.big_box = es.big_box.data(); b.big_bound = es.big_bound.data(); b.big_meta = es.big_meta.data(); b.big_count = &es.n_big;
    b.inflate = es.hp.max_abs_coord * (1.0f / 65536.0f);
    // same orchestration as build_bvh() in nrcu_api.cu
    counters[0] = 1;
    for (int i = 0; i < cap; i++) node_clear(b, i);
    for (uint32_t i = 0; i < n; i++) bvh_init_prim(b, (int)i);
    std::vector<int> big_cand(NRCU_BIG_CAND_CAP); int big_cand_count = 0;
    b.big_cand = big_cand.data(); b.big_cand_count = &big_cand_count;
    for (uint32_t i = 0; i < n; i++) bvh_big_candidate(b, (int)i);
    bvh_select_big(b, 0);
    node_clear(b, 0);
    for (uint32_t i = 0; i < n; i++) bvh_init_prim_rest(b, (int)i);
    DScene& ds = es.ds;
    ds.big_geom = es.big_geom.data(); ds.big_box = es.big_box.data(); ds.big_bound = es.big_bound.data(); ds.big_meta = es.big_meta.data(); ds.n_big = (uint32_t)es.n_big;
    ds.leaf_prims = es.leaf_prims.data(); ds.leaf_geom = es.leaf_geom.data(); ds.leaf_box = es.leaf_box.data();
    ds.root_ref = NRCU_REF_EMPTY; ds.bvh_lo = mk3(NRCU_INF); ds.bvh_hi = mk3(-NRCU_INF);
    const uint32_t n_rest = n - (uint32_t)es.n_big;
    if (n_rest == 0) return;
    bvh_padded_bounds(nbox.data(), b.inflate, ds.bvh_lo, ds.bvh_hi);
    int begin = 0, end = 1;
    for (int level = 0; level < 128; level++) {
        b.level_begin = begin; b.level_end = end;
        counters[3] = counters[4] = 0;
        for (int i = begin; i < end; i++) bvh_level_prepare(b, i);
        for (uint32_t i = 0; i < n; i++) bvh_bin(b, (int)i);
        for (int i = begin; i < end; i++) bvh_split(b, i);
        es.levels = level + 1;
        if (counters[3] == 0) break;
        for (uint32_t i = 0; i < n; i++) bvh_partition(b, (int)i);
        begin = end; end = counters[0];
    }
    const int n_nodes = counters[0];
    for (int i = 0; i < n_nodes; i++) bvh_leaf_alloc(b, i);
    for (uint32_t i = 0; i < n; i++) bvh_leaf_fill(b, (int)i);
    for (int i = 0; i < n_nodes; i++) bvh_leaf_sort(b, i);
    for (uint32_t i = 0; i < n_rest; i++) bvh_leaf_gather(b, (int)i);
    for (int i = 0; i < n_nodes; i++) bvh_wide_index(b, i);
    for (int i = 0; i < n_nodes; i++) bvh_wide_emit(b, i);
    es.n_binary_nodes = n_nodes; es.n_wide = counters[2];
    for (int i = 0; i < n_nodes; i++) if (nstate[i] == BNODE_LEAF) es.max_leaf = std::max(es.max_leaf, ncount[i]);
    es.nodes.assign(wide.begin(), wide.begin() + (size_t)std::max(1, es.n_wide) * NRCU_BVH_NODE_F4);
    es.ds.nodes = es.nodes.data();
    es.ds.leaf_prims = es.leaf_prims.data();
    es.ds.leaf_geom = es.leaf_geom.data(); es.ds.leaf_box = es.leaf_box.data();
    es.ds.root_ref = nstate[0] == BNODE_LEAF ? ~((nleaf_first[0] << 4) | (ncount[0] - 1)) : nwide[0];
}

extern "C" {

const char* emu_last_error(EmuScene* es) { return es ? es->error.c_str() : ""; }

EmuScene* emu_create(const nrcu_scene* sc, int mode) {
    EmuScene* es = new EmuScene();
    std::string why = host_prepare(sc, mode, es->hp);
    if (!why.empty()) { es->error = why; return es; }
    HostPrep& hp = es->hp;
    fill_scene_scalars(es->ds, sc, mode, hp);
    es->spp = sc->samples_per_pixel;
    es->mesh_pos.assign(sc->mesh_positions, sc->mesh_positions + 3 * (size_t)hp.total_vertices);
    for (uint32_t e : hp.mesh_nodes)
        for (uint32_t v = sc->mesh_vertex_offset[e]; v < sc->mesh_vertex_offset[e + 1]; v++) mesh_transform_vertex(es->mesh_pos.data(), v);
    const uint32_t n = es->ds.n_prims;
    PrimSources ps{};
    const uint32_t zero = 0;
    ps.src_a = hp.src_a.data(); ps.src_b = hp.src_b.data();
    ps.sphere_position = hp.sph.data(); ps.sphere_radius = sc->sphere_radius; ps.sphere_material = sc->sphere_material;
    ps.triangle_vertices = hp.tri.data(); ps.triangle_normal = sc->triangle_normal; ps.triangle_material = sc->triangle_material;
    ps.plane_normal = sc->plane_normal; ps.plane_position = hp.pln.data(); ps.plane_u = sc->plane_u; ps.plane_v = sc->plane_v; ps.plane_material = sc->plane_material;
    ps.mesh_vertex_offset = sc->n_meshes ? sc->mesh_vertex_offset : &zero; ps.mesh_index_offset = sc->n_meshes ? sc->mesh_index_offset : &zero;
    ps.mesh_positions = es->mesh_pos.data(); ps.mesh_indices = sc->mesh_indices; ps.mesh_material = sc->mesh_material;
    es->geom.assign(3 * (size_t)std::max(n, 1u), mk4(0, 0, 0, 0)); es->shade.assign(std::max(n, 1u), mk4(0, 0, 0, 0));
    es->box.assign(2 * (size_t)std::max(n, 1u), mk4(0, 0, 0, 0)); es->bound.assign(2 * (size_t)std::max(n, 1u), mk4(0, 0, 0, 0)); es->meta.assign(std::max(n, 1u), 0); es->export16.assign(16 * (size_t)std::max(n, 1u), 0.f);
    for (uint32_t i = 0; i < n; i++) build_prim(ps, i, mode == NRCU_MODE_RAYCAST, es->geom.data(), es->shade.data(), es->box.data(), es->bound.data(), es->meta.data(), es->export16.data());
    DScene& ds = es->ds;
    ds.prim_geom = es->geom.data(); ds.prim_shade = es->shade.data(); ds.prim_box = es->box.data(); ds.prim_meta = es->meta.data();
    ds.materials = hp.materials.data(); ds.mat_head = hp.mat_head.data(); ds.area_lights = hp.lights.data();
    if (sc->ambient_type == NRCU_AMBIENT_ENVIRONMENT_MAP && sc->ambient_environment_map >= 0 && (uint32_t)sc->ambient_environment_map < sc->n_textures) {
        uint32_t ti = (uint32_t)sc->ambient_environment_map;
        size_t cnt = (size_t)sc->texture_width[ti] * sc->texture_height[ti];
        if (cnt) {
            es->env.resize(cnt);
            std::memcpy(es->env.data(), sc->texture_rgba + sc->texture_offset[ti], cnt * sizeof(f4));
            ds.env_rgba = es->env.data(); ds.env_w = (int)sc->texture_width[ti]; ds.env_h = (int)sc->texture_height[ti];
            es->env_tab.assign(2 * (size_t)ds.env_h + cnt, 0.f);   // the k_env_rows / k_env_marginal work items, one after the other
            for (int y = 0; y < ds.env_h; y++) env_table_row(ds.env_rgba, ds.env_w, ds.env_h, y, es->env_tab.data());
            ds.env_total = env_table_marginal(ds.env_h, es->env_tab.data());
            ds.env_tab = es->env_tab.data();
        }
    }
    if (mode != NRCU_MODE_RAYCAST && n > 0) emu_build_bvh(*es);
    return es;
}

void emu_destroy(EmuScene* es) { delete es; }
uint32_t emu_primitive_count(EmuScene* es) { return es->ds.n_prims; }

void emu_primitives(EmuScene* es, uint32_t* kind, float* data16, int32_t* material, float* box6) {
    for (uint32_t i = 0; i < es->ds.n_prims; i++) {
        if (kind) kind[i] = es->meta[i] & 3u;
        if (material) material[i] = (int32_t)(es->meta[i] >> 2);
        if (data16) std::memcpy(data16 + 16 * (size_t)i, es->export16.data() + 16 * (size_t)i, 64);
        if (box6) { box6[6 * i] = es->box[2 * i].x; box6[6 * i + 1] = es->box[2 * i].y; box6[6 * i + 2] = es->box[2 * i].z;
                    box6[6 * i + 3] = es->box[2 * i + 1].x; box6[6 * i + 4] = es->box[2 * i + 1].y; box6[6 * i + 5] = es->box[2 * i + 1].z; }
    }
}

void emu_camera(EmuScene* es, float* cam18, float* lens_radius) {
    const DCamera& c = es->ds.cam;
    st3(cam18, c.position); st3(cam18 + 3, c.lower_left); st3(cam18 + 6, c.horizontal); st3(cam18 + 9, c.vertical); st3(cam18 + 12, c.u); st3(cam18 + 15, c.v);
    *lens_radius = c.lens_radius;
}

// out[0..5]: binary nodes, wide nodes, build levels, max leaf size, leaf prim slots used, root ref
void emu_bvh_stats(EmuScene* es, int32_t* out) {
    out[0] = es->n_binary_nodes; out[1] = es->n_wide; out[2] = es->levels; out[3] = es->max_leaf;
    out[4] = (int32_t)es->leaf_prims.size(); out[5] = es->ds.root_ref; out[6] = es->n_big;
}

void emu_trace_batch(EmuScene* es, const float* rays, uint32_t n, int32_t* prim_id, float* t, int use_linear) {
    for (uint32_t i = 0; i < n; i++) {
        Ray r; r.o = ld3(rays + 6 * (size_t)i); r.d = ld3(rays + 6 * (size_t)i + 3);
        float tt; int id;
        if (es->ds.mode == MODE_RAYCAST) closest_hit_linear<true>(es->ds, r, tt, id);
        else if (use_linear) closest_hit_linear<false>(es->ds, r, tt, id);
        else { LocalStack st; if (es->ds.mode == MODE_ACC) closest_hit_bvh<true>(es->ds, r, st, tt, id); else closest_hit_bvh<false>(es->ds, r, st, tt, id); }
        prim_id[i] = id; t[i] = tt;
    }
}

void emu_render_raycast(EmuScene* es, float* rgba) {
    uint32_t n = es->ds.width * es->ds.height, rays = 0;
    for (uint32_t p = 0; p < n; p++) {
        vec3 c = raycast_pixel(es->ds, p, &rays);
        rgba[4 * (size_t)p] = c.x; rgba[4 * (size_t)p + 1] = c.y; rgba[4 * (size_t)p + 2] = c.z; rgba[4 * (size_t)p + 3] = 1.f;
    }
}

// Same dataflow as the wavefront kernels, one path at a time: raygen -> (trace -> shade)* ; paths
// that split (glass branch mode) go through a small stack.  accum: w*h*4 (or n_pixels*4), sums + count.
void emu_render_pt(EmuScene* es, uint64_t seed, uint32_t s0, uint32_t s1, int glass_branch,
                   const uint32_t* pixels, uint32_t n_pixels, float* accum4, uint64_t* rays_out, uint32_t flags) {
    es->ds.nee = ((flags & NRCU_FLAG_ENV_IS) && es->ds.mode == MODE_ACC && es->ds.env_tab && es->ds.env_total > 0.f) ? 2
               : (((flags & NRCU_FLAG_NEE) && es->ds.n_area_lights > 0) ? 1 : 0);
    const DScene& ds = es->ds;
    if (s0 == 0 && s1 == 0) s1 = es->spp;
    uint64_t rays = 0;
    struct Work { Ray r; vec3 thr; uint32_t d, branch; bool skip_light; };
    std::vector<Work> stack;
    uint32_t total = pixels ? n_pixels : ds.width * ds.height;
    for (uint32_t q = 0; q < total; q++) {
        uint32_t p = pixels ? pixels[q] : q;
        float* a = accum4 + 4 * (size_t)q;
        for (uint32_t k = s0; k < s1; k++) {
            vec3 L = mk3(0.f);
            if (ds.depth == 0) L = ds.ambient;
            else {
                stack.clear();
                stack.push_back({pt_camera_ray(ds, seed, p, k), mk3(1.f), 0u, 0u, false});
                while (!stack.empty()) {
                    Work w = stack.back(); stack.pop_back();
                    for (;;) {
                        float t; int id;
                        LocalStack st;
                        if (ds.mode == MODE_ACC) closest_hit_bvh<true>(ds, w.r, st, t, id); else closest_hit_bvh<false>(ds, w.r, st, t, id);
                        rays++;
                        PathStep ps = ds.nee ? path_vertex<true>(ds, seed, p, k, w.d, w.branch, w.r, w.thr, t, id, glass_branch, w.skip_light)
                                             : path_vertex<false>(ds, seed, p, k, w.d, w.branch, w.r, w.thr, t, id, glass_branch, false);
                        if (ps.nee) {   // shadow ray: same closest-hit query, then visibility of the sampled light
                            float st; int sid; LocalStack sst;
                            if (ds.mode == MODE_ACC) closest_hit_bvh<true>(ds, ps.shadow, sst, st, sid); else closest_hit_bvh<false>(ds, ps.shadow, sst, st, sid);
                            rays++;
                            if (nee_visible(ds, ps.shadow, ps.nee_light, st, sid)) L = L + ps.nee_contrib;
                        }
                        if (ps.action == PATH_TERMINATE) { L = L + ps.radiance; break; }
                        if (ps.action == PATH_SPLIT) stack.push_back({ps.next2, ps.thr2, w.d + 1, w.branch | (1u << (w.d & 31u)), ps.next_skips_light});
                        w.r = ps.next; w.thr = ps.thr; w.d++; w.skip_light = ps.next_skips_light;
                    }
                }
            }
            a[0] += L.x; a[1] += L.y; a[2] += L.z;
        }
        a[3] += (float)(s1 - s0);
    }
    if (rays_out) *rays_out = rays;
}

void emu_camera_ray(EmuScene* es, uint64_t seed, uint32_t pixel, uint32_t sample, float* out6) {
    Ray r = pt_camera_ray(es->ds, seed, pixel, sample); st3(out6, r.o); st3(out6 + 3, r.d);
}

void emu_philox4x32(const uint32_t c[4], const uint32_t k[2], uint32_t out[4]) {
    u32x4 r = philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // extern "C"
