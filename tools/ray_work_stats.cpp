// ray_work_stats.cpp — research tool (not product, not a test): per-ray traversal work of the wide BVH on
// real path-traced ray populations, bounce by bounce, computed with the CPU emulation of the device code.
//   g++ -O2 -std=c++17 -ffp-contract=off -DNRCU_HOST_EMU=1 -Iinclude -Inrenderer_b200/csrc tools/ray_work_stats.cpp -o /tmp/ray_work_stats
// Reads a flat scene (.nrsc is parsed by Python; this tool is driven through ctypes from tools/ray_work_stats.py).
#include "../tests/host_emu/nrcu_emu.cpp"

struct CountStack : LocalStack {};

template <bool GATE>
static void traverse_count(const DScene& s, const Ray& ray, float& best_t, int& best_id, int& nodes, int& leaves, int& prims) {
    best_t = NRCU_INF; best_id = -1; nodes = leaves = prims = 0;
    RayPrep rp = prep_ray(ray);
    vec3 ginv = gate_inverse(ray);
    big_list_step<GATE>(s, s.big_geom, s.big_box, s.big_bound, s.big_meta, ray, rp, ginv, best_t, best_id);
    if (!bvh_reachable(s, rp, best_t)) return;
    int cur = s.root_ref;
    LocalStack stack;
    for (;;) {
        while (cur >= 0) { cur = node_step(s, rp, cur, best_t, stack); nodes++; }
        if (cur == NRCU_REF_DONE) return;
        leaves++; prims += (int)(((uint32_t)(~cur)) & 15u) + 1;
        leaf_step<GATE>(s, ray, ginv, cur, best_t, best_id);
        cur = pop_next(stack, best_t);
    }
}

extern "C" {
// For the given pixels (in queue order) follow 1 sample per pixel through the path tracer and record per ray:
// out[k*4..] = (bounce, nodes, leaves, prims) in wavefront queue order bounce by bounce.  Returns the ray count.
uint32_t emu_ray_work(EmuScene* es, uint64_t seed, const uint32_t* pixels, uint32_t n_pixels, uint32_t sample, int32_t* out, uint32_t cap) {
    const DScene& ds = es->ds;
    struct P { Ray r; vec3 thr; uint32_t pixel; };
    std::vector<P> cur, nxt;
    for (uint32_t q = 0; q < n_pixels; q++) cur.push_back({pt_camera_ray(ds, seed, pixels[q], sample), mk3(1.f), pixels[q]});
    uint32_t k = 0;
    for (uint32_t d = 0; d < ds.depth && !cur.empty(); d++) {
        nxt.clear();
        for (auto& p : cur) {
            float t; int id, nn, nl, np;
            if (ds.mode == MODE_ACC) traverse_count<true>(ds, p.r, t, id, nn, nl, np); else traverse_count<false>(ds, p.r, t, id, nn, nl, np);
            if (k < cap) { out[4 * k] = (int)d; out[4 * k + 1] = nn; out[4 * k + 2] = nl; out[4 * k + 3] = np; k++; }
            PathStep ps = path_vertex(ds, seed, p.pixel, sample, d, 0, p.r, p.thr, t, id, 0, false);
            if (ps.action == PATH_CONTINUE) nxt.push_back({ps.next, ps.thr, p.pixel});
        }
        cur.swap(nxt);
    }
    return k;
}

// Stage-1 statistics for the same ray population: out[k*2..] = (bounce, candidate bit mask of the wide list after the
// conservative slab test) in wavefront queue order.  Returns the ray count.
uint32_t emu_stage1_masks(EmuScene* es, uint64_t seed, const uint32_t* pixels, uint32_t n_pixels, uint32_t sample, uint32_t* out, uint32_t cap) {
    const DScene& ds = es->ds;
    struct P { Ray r; vec3 thr; uint32_t pixel; };
    std::vector<P> cur, nxt;
    for (uint32_t q = 0; q < n_pixels; q++) cur.push_back({pt_camera_ray(ds, seed, pixels[q], sample), mk3(1.f), pixels[q]});
    uint32_t k = 0;
    for (uint32_t d = 0; d < ds.depth && !cur.empty(); d++) {
        nxt.clear();
        for (auto& p : cur) {
            RayPrep rp = prep_ray(p.r);
            const vec3 ainv = mk3(fabsf(rp.inv.x), fabsf(rp.inv.y), fabsf(rp.inv.z));
            uint32_t mask = 0;
            for (uint32_t b = 0; b < ds.n_big; b++) {
                float tn, tf;
                slab_center_extent(ds.big_bound[2 * b], ds.big_bound[2 * b + 1], rp, ainv, -rp.oinv.x, tn, tf);
                if (tn <= tf) mask |= 1u << b;
            }
            if (k < cap) { out[2 * k] = d; out[2 * k + 1] = mask; k++; }
            float t; int id, nn, nl, np;
            if (ds.mode == MODE_ACC) traverse_count<true>(ds, p.r, t, id, nn, nl, np); else traverse_count<false>(ds, p.r, t, id, nn, nl, np);
            PathStep ps = path_vertex(ds, seed, p.pixel, sample, d, 0, p.r, p.thr, t, id, 0, false);
            if (ps.action == PATH_CONTINUE) nxt.push_back({ps.next, ps.thr, p.pixel});
        }
        cur.swap(nxt);
    }
    return k;
}
}
