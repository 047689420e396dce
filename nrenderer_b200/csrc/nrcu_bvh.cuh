// nrcu_bvh.cuh — GPU binned-SAH BVH build + emission of the 4-wide node layout.
//
// Replaces BVHTree::build (reference code/components/acc_path_tracing/include/BVH.hpp:166-222:
// recursive median split, one primitive per leaf, never pruned by t).  Input: the per-primitive
// boxes (the reference's own Bounds3 definitions, Bounds3.hpp:35-103).  The build is breadth
// first, one level per round of three kernels, primitives are never moved — each carries the id
// of the binary node it currently belongs to:
//   bin       every primitive of a node with > LEAF_MAX primitives adds its box to 3 x NBINS bins
//             of that node (atomic min/max on order-preserving int keys)
//   split     one thread per node of the level sweeps the bins, picks the minimum-SAH plane over
//             the three axes (falls back to a primitive-id median when all centroids coincide)
//             and allocates two children
//   partition every primitive moves to its child and grows the child's box / centroid box
// Then leaves get contiguous ranges of `leaf_prims`, and every binary inner node at even depth
// becomes one BVH4 node whose slots are its grandchildren (or children that are leaves).
// Each step body is a __host__ __device__ function of the work-item index so that
// tests/host_emu can run the identical code sequentially on the CPU.
#pragma once
#include "nrcu_scene.cuh"

namespace nrcu {

#define NRCU_NBINS 16
#define NRCU_SAH_MAX_DEPTH 24
#define NRCU_BIN_WORDS 7   // 6 encoded bounds + count

// order-preserving float <-> int key (so that integer atomicMin/Max order floats)
NR_HD int fkey(float f) { int i = f2i(f); return i >= 0 ? i : (i ^ 0x7fffffff); }
NR_HD float fkey_inv(int k) { return i2f(k >= 0 ? k : (k ^ 0x7fffffff)); }
#define NRCU_KEY_POS_INF 0x7f800000
#define NRCU_KEY_NEG_INF ((int)(0xff800000u ^ 0x7fffffffu))

NR_HD int atomic_add_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    int o = *p; *p = o + v; return o;
#endif
}
NR_HD void atomic_min_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
NR_HD void atomic_max_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}

enum { BNODE_OPEN = 0, BNODE_LEAF = 1, BNODE_INNER = 2 };

struct BvhBuild {
    uint32_t n_prims;
    const f4* prim_box;        // 2 per primitive: the reference's leaf box (gate)
    const f4* prim_bound;      // 2 per primitive: true bounds (tree construction)
    const uint32_t* prim_meta; // kind | material << 2
    int* prim_node;            // [n_prims] binary node of each primitive
    // binary nodes, capacity 2*n_prims + 2
    int* nbox;                 // [cap*6] encoded bounds  (min xyz, max xyz)
    int* cbox;                 // [cap*6] encoded centroid bounds
    int* ncount;               // [cap]
    int* nidmin; int* nidmax;  // [cap] primitive id range (fallback split)
    int* nstate;               // [cap] BNODE_*
    int* nchild;               // [cap] index of the left child (right = +1)
    int* nsplit_axis;          // [cap] 0..2 spatial, 3 = id median
    float* nsplit_pos;         // [cap] centroid threshold: left iff centroid[axis] < pos (or id <= pos bits)
    int* ndepth;               // [cap]
    int* nleaf_first;          // [cap] first slot in leaf_prims
    int* nleaf_fill;           // [cap]
    int* nwide;                // [cap] wide node index of even-depth inner nodes
    int* bins;                 // [bin_nodes * 3 * NBINS * BIN_WORDS]
    int bin_nodes;             // capacity of `bins` in nodes
    int* counters;             // [0] node count, [1] leaf cursor, [2] wide count, [3] splits in this level, [4] bin-slot cursor
    int* nbin_slot;            // [cap] bin slot of a node in the current level
    int level_begin, level_end;
    // outputs
    uint32_t* leaf_prims;      // [n_prims]
    const f4* prim_geom;       // 3 per primitive (gather source)
    f4* leaf_geom;             // [n_prims * 3] prim_geom in leaf order
    f4* leaf_box;              // [n_prims * 2] prim_box in leaf order
    // wide-primitive list (outputs of bvh_select_big)
    f4* big_geom; f4* big_box; f4* big_bound; uint32_t* big_meta;   // capacity NRCU_MAX_BIG; big_bound = padded true bounds as (centre, half extent)
    int* big_count;            // [1]
    int* big_cand;             // [NRCU_BIG_CAND_CAP] ids whose box is large enough (unordered), filled by bvh_big_candidate
    int* big_cand_count;       // [1]
    f4* wide_nodes;            // [wide capacity * 7]
    float inflate;             // absolute padding of wide-node boxes
};

NR_HD vec3 box_centroid(f4 lo, f4 hi) { return mk3(0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z); }

NR_HD void node_clear(const BvhBuild& b, int n) {
    for (int k = 0; k < 3; k++) {
        b.nbox[n * 6 + k] = NRCU_KEY_POS_INF; b.nbox[n * 6 + 3 + k] = NRCU_KEY_NEG_INF;
        b.cbox[n * 6 + k] = NRCU_KEY_POS_INF; b.cbox[n * 6 + 3 + k] = NRCU_KEY_NEG_INF;
    }
    b.ncount[n] = 0; b.nidmin[n] = 0x7fffffff; b.nidmax[n] = -1; b.nstate[n] = BNODE_OPEN; b.nchild[n] = -1;
    b.nsplit_axis[n] = 0; b.nsplit_pos[n] = 0.f; b.nleaf_first[n] = 0; b.nleaf_fill[n] = 0; b.nwide[n] = -1; b.nbin_slot[n] = -1;
    if (n == 0) b.ndepth[n] = 0;
}

NR_HD void node_add_prim(const BvhBuild& b, int n, int i) {
    f4 lo = b.prim_bound[2 * i], hi = b.prim_bound[2 * i + 1];
    vec3 c = box_centroid(lo, hi);
    atomic_min_i(&b.nbox[n * 6 + 0], fkey(lo.x)); atomic_min_i(&b.nbox[n * 6 + 1], fkey(lo.y)); atomic_min_i(&b.nbox[n * 6 + 2], fkey(lo.z));
    atomic_max_i(&b.nbox[n * 6 + 3], fkey(hi.x)); atomic_max_i(&b.nbox[n * 6 + 4], fkey(hi.y)); atomic_max_i(&b.nbox[n * 6 + 5], fkey(hi.z));
    atomic_min_i(&b.cbox[n * 6 + 0], fkey(c.x)); atomic_min_i(&b.cbox[n * 6 + 1], fkey(c.y)); atomic_min_i(&b.cbox[n * 6 + 2], fkey(c.z));
    atomic_max_i(&b.cbox[n * 6 + 3], fkey(c.x)); atomic_max_i(&b.cbox[n * 6 + 4], fkey(c.y)); atomic_max_i(&b.cbox[n * 6 + 5], fkey(c.z));
    atomic_add_i(&b.ncount[n], 1);
    atomic_min_i(&b.nidmin[n], i); atomic_max_i(&b.nidmax[n], i);
}

// step 0: one work item per primitive
NR_HD void bvh_init_prim(const BvhBuild& b, int i) { b.prim_node[i] = 0; node_add_prim(b, 0, i); }

NR_HD float half_area(const float* lo, const float* hi) {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

#define NRCU_PRIM_EXCLUDED (-2)
NR_HD float box_half_area(f4 lo, f4 hi) {
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}
// step 0b (ONE work item, after steps 0 and 0a): pick the "wide" primitives — box
// surface area >= NRCU_BIG_AREA_FRACTION of the scene box's, the NRCU_MAX_BIG largest if there are more (ties to
// the lower id) — copy their records to the wide list in ascending id order and mark them excluded from the BVH.
// A primitive this large is entered by most rays whatever the tree looks like (its SAH probability is ~1), so it
// is cheaper to test it once per ray in a warp-uniform loop than to let it bloat the boxes of the top BVH levels.
#define NRCU_BIG_CAND_CAP 1024
// step 0a (per primitive, after step 0): collect the ids whose box is large enough, in any order
NR_HD void bvh_big_candidate(const BvhBuild& b, int i) {
    float rlo[3], rhi[3];
    for (int a = 0; a < 3; a++) { rlo[a] = fkey_inv(b.nbox[a]); rhi[a] = fkey_inv(b.nbox[3 + a]); }
    const float root_area = half_area(rlo, rhi);
    if (!(root_area > 0.f && root_area < NRCU_INF)) return;
    if (!(box_half_area(b.prim_bound[2 * i], b.prim_bound[2 * i + 1]) >= NRCU_BIG_AREA_FRACTION * root_area)) return;
    int k = atomic_add_i(b.big_cand_count, 1);
    if (k < NRCU_BIG_CAND_CAP) b.big_cand[k] = i;
}
NR_HD void bvh_select_big(const BvhBuild& b, int) {
    float rlo[3], rhi[3];
    for (int a = 0; a < 3; a++) { rlo[a] = fkey_inv(b.nbox[a]); rhi[a] = fkey_inv(b.nbox[3 + a]); }
    const float root_area = half_area(rlo, rhi);
    int ids[NRCU_MAX_BIG]; float areas[NRCU_MAX_BIG]; int cnt = 0;
    if (root_area > 0.f && root_area < NRCU_INF) {
        const float thr = NRCU_BIG_AREA_FRACTION * root_area;
        // the candidate list (any order) when it did not overflow, else every primitive; the selection is a total
        // order (area descending, id ascending), so the result does not depend on the order of the candidates
        const int n_cand = *b.big_cand_count;
        const bool listed = n_cand <= NRCU_BIG_CAND_CAP;
        const int n_scan = listed ? n_cand : (int)b.n_prims;
        for (int q = 0; q < n_scan; q++) {
            const int i = listed ? b.big_cand[q] : q;
            float ar = box_half_area(b.prim_bound[2 * i], b.prim_bound[2 * i + 1]);
            if (!(ar >= thr)) continue;
            int pos = cnt;
            while (pos > 0 && (areas[pos - 1] < ar || (areas[pos - 1] == ar && ids[pos - 1] > i))) pos--;
            if (pos >= NRCU_MAX_BIG) continue;
            int last = cnt < NRCU_MAX_BIG ? cnt : NRCU_MAX_BIG - 1;
            for (int k = last; k > pos; k--) { ids[k] = ids[k - 1]; areas[k] = areas[k - 1]; }
            ids[pos] = i; areas[pos] = ar;
            if (cnt < NRCU_MAX_BIG) cnt++;
        }
    }
    // order: planes, then triangles, then spheres (so that the lanes of a warp run the same test most of the
    // time), ascending id inside each class; ties between equal-t hits are settled by id, not by list order
    for (int i = 1; i < cnt; i++) {
        int v = ids[i], j = i - 1;
        #define NRCU_KCLASS(id_) ((b.prim_meta[id_] & 3u) == KIND_PLANE ? 0 : ((b.prim_meta[id_] & 3u) == KIND_SPHERE ? 2 : 1))
        while (j >= 0 && (NRCU_KCLASS(ids[j]) > NRCU_KCLASS(v) || (NRCU_KCLASS(ids[j]) == NRCU_KCLASS(v) && ids[j] > v))) { ids[j + 1] = ids[j]; j--; }
        #undef NRCU_KCLASS
        ids[j + 1] = v;
    }
    for (int k = 0; k < cnt; k++) {
        int id = ids[k];
        for (int q = 0; q < 3; q++) b.big_geom[3 * k + q] = b.prim_geom[3 * (size_t)id + q];
        for (int q = 0; q < 2; q++) b.big_box[2 * k + q] = b.prim_box[2 * (size_t)id + q];
        f4 lo = b.prim_bound[2 * (size_t)id], hi = b.prim_bound[2 * (size_t)id + 1];
        float px = b.inflate + 1e-6f * fmaxf(fabsf(lo.x), fabsf(hi.x)), py = b.inflate + 1e-6f * fmaxf(fabsf(lo.y), fabsf(hi.y)),
              pz = b.inflate + 1e-6f * fmaxf(fabsf(lo.z), fabsf(hi.z));
        // stored as (centre, half extent) for the three-FFMA slab test (slab_center_extent, nrcu_intersect.cuh); the half
        // extent is rounded outwards so that [c - h, c + h] contains the padded box
        const float l3[3] = {lo.x - px, lo.y - py, lo.z - pz}, h3[3] = {hi.x + px, hi.y + py, hi.z + pz};
        float c3[3], e3[3];
        for (int q = 0; q < 3; q++) { c3[q] = 0.5f * l3[q] + 0.5f * h3[q]; e3[q] = fmaxf(h3[q] - c3[q], c3[q] - l3[q]) * 1.000001f + 1e-30f; }
        b.big_bound[2 * k] = mk4(c3[0], c3[1], c3[2], 0.f); b.big_bound[2 * k + 1] = mk4(e3[0], e3[1], e3[2], 0.f);
        b.big_meta[k] = ((uint32_t)id << 2) | (b.prim_meta[id] & 3u);
        b.prim_node[id] = NRCU_PRIM_EXCLUDED;
    }
    *b.big_count = cnt;
}
// step 0c: one work item per primitive — rebuild node 0 from the primitives that stay in the BVH
// (node 0 must have been cleared again with node_clear first)
NR_HD void bvh_init_prim_rest(const BvhBuild& b, int i) { if (b.prim_node[i] != NRCU_PRIM_EXCLUDED) node_add_prim(b, 0, i); }

// step A (per node of the level): decide leaf / open and hand out a bin slot
NR_HD void bvh_level_prepare(const BvhBuild& b, int n) {
    if (b.ncount[n] <= NRCU_LEAF_MAX) { b.nstate[n] = BNODE_LEAF; return; }
    int slot = atomic_add_i(&b.counters[4], 1);
    b.nbin_slot[n] = slot;
    if (slot < b.bin_nodes) {
        int* bn = b.bins + (size_t)slot * 3 * NRCU_NBINS * NRCU_BIN_WORDS;
        for (int k = 0; k < 3 * NRCU_NBINS; k++) {
            int* w = bn + k * NRCU_BIN_WORDS;
            w[0] = w[1] = w[2] = NRCU_KEY_POS_INF; w[3] = w[4] = w[5] = NRCU_KEY_NEG_INF; w[6] = 0;
        }
    }
}

NR_HD int bin_of(float c, float lo, float hi) {
    float ext = hi - lo;
    if (!(ext > 0.f)) return 0;
    int k = (int)((c - lo) * ((float)NRCU_NBINS / ext));
    return k < 0 ? 0 : (k > NRCU_NBINS - 1 ? NRCU_NBINS - 1 : k);
}

// step B (per primitive)
NR_HD void bvh_bin(const BvhBuild& b, int i) {
    int n = b.prim_node[i];
    if (n < 0 || n < b.level_begin || b.nstate[n] != BNODE_OPEN) return;
    int slot = b.nbin_slot[n];
    if (slot < 0 || slot >= b.bin_nodes) return;
    f4 lo = b.prim_bound[2 * i], hi = b.prim_bound[2 * i + 1];
    vec3 c = box_centroid(lo, hi);
    int* bn = b.bins + (size_t)slot * 3 * NRCU_NBINS * NRCU_BIN_WORDS;
    for (int a = 0; a < 3; a++) {
        float clo = fkey_inv(b.cbox[n * 6 + a]), chi = fkey_inv(b.cbox[n * 6 + 3 + a]);
        int k = bin_of(comp(c, a), clo, chi);
        int* w = bn + (a * NRCU_NBINS + k) * NRCU_BIN_WORDS;
        atomic_min_i(&w[0], fkey(lo.x)); atomic_min_i(&w[1], fkey(lo.y)); atomic_min_i(&w[2], fkey(lo.z));
        atomic_max_i(&w[3], fkey(hi.x)); atomic_max_i(&w[4], fkey(hi.y)); atomic_max_i(&w[5], fkey(hi.z));
        atomic_add_i(&w[6], 1);
    }
}


// step C (per node of the level): choose the split and allocate the children
NR_HD void bvh_split(const BvhBuild& b, int n) {
    if (b.nstate[n] != BNODE_OPEN) return;
    int slot = b.nbin_slot[n];
    int best_axis = -1, best_k = 0;
    float best_cost = NRCU_INF;
    // Depth cap: SAH planes may peel off one primitive per level on adversarial inputs.  From NRCU_SAH_MAX_DEPTH on, nodes
    // are split at the median of their primitive-id range, which at least halves that range per level: the binary tree
    // is at most NRCU_SAH_MAX_DEPTH + 32 levels deep, i.e. <= (24 + 32) / 2 = 28 BVH4 levels x 3 pushes = 84 stack
    // entries < NRCU_LOCAL_STACK (96), whatever the scene.
    if (slot >= 0 && slot < b.bin_nodes && b.ndepth[n] < NRCU_SAH_MAX_DEPTH) {
        const int* bn = b.bins + (size_t)slot * 3 * NRCU_NBINS * NRCU_BIN_WORDS;
        for (int a = 0; a < 3; a++) {
            float clo = fkey_inv(b.cbox[n * 6 + a]), chi = fkey_inv(b.cbox[n * 6 + 3 + a]);
            if (!(chi > clo)) continue;
            // suffix sweep: right_area[k] = SAH term of bins [k, NBINS)
            float r_area[NRCU_NBINS]; int r_cnt[NRCU_NBINS];
            float lo[3] = {NRCU_INF, NRCU_INF, NRCU_INF}, hi[3] = {-NRCU_INF, -NRCU_INF, -NRCU_INF};
            int cnt = 0;
            for (int k = NRCU_NBINS - 1; k >= 1; k--) {
                const int* w = bn + (a * NRCU_NBINS + k) * NRCU_BIN_WORDS;
                if (w[6] > 0) {
                    for (int c = 0; c < 3; c++) { lo[c] = fminf(lo[c], fkey_inv(w[c])); hi[c] = fmaxf(hi[c], fkey_inv(w[3 + c])); }
                    cnt += w[6];
                }
                r_cnt[k] = cnt; r_area[k] = cnt > 0 ? half_area(lo, hi) : 0.f;
            }
            float llo[3] = {NRCU_INF, NRCU_INF, NRCU_INF}, lhi[3] = {-NRCU_INF, -NRCU_INF, -NRCU_INF};
            int lcnt = 0;
            for (int k = 0; k < NRCU_NBINS - 1; k++) {   // split after bin k
                const int* w = bn + (a * NRCU_NBINS + k) * NRCU_BIN_WORDS;
                if (w[6] > 0) {
                    for (int c = 0; c < 3; c++) { llo[c] = fminf(llo[c], fkey_inv(w[c])); lhi[c] = fmaxf(lhi[c], fkey_inv(w[3 + c])); }
                    lcnt += w[6];
                }
                int rc = r_cnt[k + 1];
                if (lcnt == 0 || rc == 0) continue;
                float cost = half_area(llo, lhi) * (float)lcnt + r_area[k + 1] * (float)rc;
                if (cost < best_cost) { best_cost = cost; best_axis = a; best_k = k; }
            }
        }
    }
    int c = atomic_add_i(&b.counters[0], 2);
    b.nchild[n] = c; b.nstate[n] = BNODE_INNER;
    b.ndepth[c] = b.ndepth[n] + 1; b.ndepth[c + 1] = b.ndepth[n] + 1;
    atomic_add_i(&b.counters[3], 1);
    if (best_axis >= 0) {
        float clo = fkey_inv(b.cbox[n * 6 + best_axis]), chi = fkey_inv(b.cbox[n * 6 + 3 + best_axis]);
        b.nsplit_axis[n] = best_axis;
        b.nsplit_pos[n] = (float)(best_k + 1);   // bin boundary: left iff bin_of(c) <= best_k
        (void)clo; (void)chi;
    } else {
        b.nsplit_axis[n] = 3;
        b.nsplit_pos[n] = i2f((b.nidmin[n] >> 1) + (b.nidmax[n] >> 1));   // id median, stored as bits
    }
}

// step D (per primitive)
NR_HD void bvh_partition(const BvhBuild& b, int i) {
    int n = b.prim_node[i];
    if (n < 0 || n < b.level_begin || b.nstate[n] != BNODE_INNER) return;
    int side;
    int a = b.nsplit_axis[n];
    if (a == 3) side = (i <= f2i(b.nsplit_pos[n])) ? 0 : 1;
    else {
        f4 lo = b.prim_bound[2 * i], hi = b.prim_bound[2 * i + 1];
        vec3 c = box_centroid(lo, hi);
        float clo = fkey_inv(b.cbox[n * 6 + a]), chi = fkey_inv(b.cbox[n * 6 + 3 + a]);
        side = ((float)bin_of(comp(c, a), clo, chi) < b.nsplit_pos[n]) ? 0 : 1;
    }
    int child = b.nchild[n] + side;
    b.prim_node[i] = child;
    node_add_prim(b, child, i);
}

// leaves: per node, then per primitive, then per node (sort ids for a deterministic layout)
NR_HD void bvh_leaf_alloc(const BvhBuild& b, int n) {
    if (b.nstate[n] != BNODE_LEAF) return;
    b.nleaf_first[n] = atomic_add_i(&b.counters[1], b.ncount[n]);
}
NR_HD void bvh_leaf_fill(const BvhBuild& b, int i) {
    int n = b.prim_node[i];
    if (n < 0) return;   // wide primitive, not in the tree
    int pos = b.nleaf_first[n] + atomic_add_i(&b.nleaf_fill[n], 1);
    b.leaf_prims[pos] = ((uint32_t)i << 2) | (b.prim_meta[i] & 3u);
}
NR_HD void bvh_leaf_sort(const BvhBuild& b, int n) {
    if (b.nstate[n] != BNODE_LEAF) return;
    uint32_t* p = b.leaf_prims + b.nleaf_first[n];
    int c = b.ncount[n];
    for (int i = 1; i < c; i++) { uint32_t v = p[i]; int j = i - 1; while (j >= 0 && p[j] > v) { p[j + 1] = p[j]; j--; } p[j + 1] = v; }
}

// per leaf slot: copy the primitive's intersection record and reference box next to its neighbours in the leaf
NR_HD void bvh_leaf_gather(const BvhBuild& b, int slot) {
    uint32_t id = b.leaf_prims[slot] >> 2;
    for (int k = 0; k < 3; k++) b.leaf_geom[3 * (size_t)slot + k] = b.prim_geom[3 * (size_t)id + k];
    for (int k = 0; k < 2; k++) b.leaf_box[2 * (size_t)slot + k] = b.prim_box[2 * (size_t)id + k];
}

NR_HD int leaf_ref(const BvhBuild& b, int n) { return ~((b.nleaf_first[n] << 4) | (b.ncount[n] - 1)); }

// Padded bounds of the whole tree from node 0's encoded box (same padding as the wide-node boxes).
NR_HD void bvh_padded_bounds(const int* root_box_keys, float inflate, vec3& lo, vec3& hi) {
    float l[3], h[3];
    for (int a = 0; a < 3; a++) {
        float lv = fkey_inv(root_box_keys[a]), hv = fkey_inv(root_box_keys[3 + a]);
        float pad = inflate + 1e-6f * fmaxf(fabsf(lv), fabsf(hv));
        l[a] = lv - pad; h[a] = hv + pad;
    }
    lo = mk3(l[0], l[1], l[2]); hi = mk3(h[0], h[1], h[2]);
}

// wide emission, pass 1 (per binary node): hand out wide indices
NR_HD void bvh_wide_index(const BvhBuild& b, int n) {
    if (b.nstate[n] == BNODE_INNER && (b.ndepth[n] & 1) == 0) b.nwide[n] = atomic_add_i(&b.counters[2], 1);
}
// pass 2 (per binary node): fill the node
NR_HD void bvh_wide_emit(const BvhBuild& b, int n) {
    if (b.nwide[n] < 0) return;
    int slots[4]; int ns = 0;
    for (int s = 0; s < 2; s++) {
        int c = b.nchild[n] + s;
        if (b.nstate[c] == BNODE_LEAF) { if (b.ncount[c] > 0) slots[ns++] = c; }
        else { for (int g = 0; g < 2; g++) { int gc = b.nchild[c] + g; if (b.nstate[gc] != BNODE_LEAF || b.ncount[gc] > 0) slots[ns++] = gc; } }
    }
    float lo[3][4], hi[3][4]; int ref[4];
    for (int k = 0; k < 4; k++) {
        if (k < ns) {
            int c = slots[k];
            for (int a = 0; a < 3; a++) {
                float l = fkey_inv(b.nbox[c * 6 + a]), h = fkey_inv(b.nbox[c * 6 + 3 + a]);
                float pad = b.inflate + 1e-6f * fmaxf(fabsf(l), fabsf(h));
                // stored as centre (lo[][]) and half extent rounded outwards (hi[][]): see NRCU_SLAB in node_step
                const float pl = l - pad, ph = h + pad, c0 = 0.5f * pl + 0.5f * ph;
                lo[a][k] = c0; hi[a][k] = fmaxf(ph - c0, c0 - pl) * 1.000001f + 1e-30f;
            }
            ref[k] = b.nstate[c] == BNODE_LEAF ? leaf_ref(b, c) : b.nwide[c];
        } else {
            for (int a = 0; a < 3; a++) { lo[a][k] = NRCU_INF; hi[a][k] = 0.f; }   // centre at infinity: never entered (see closest_hit_bvh)
            ref[k] = NRCU_REF_EMPTY;
        }
    }
    f4* w = b.wide_nodes + (size_t)b.nwide[n] * NRCU_BVH_NODE_F4;
    for (int a = 0; a < 3; a++) {
        w[2 * a] = mk4(lo[a][0], lo[a][1], lo[a][2], lo[a][3]);
        w[2 * a + 1] = mk4(hi[a][0], hi[a][1], hi[a][2], hi[a][3]);
    }
    w[6] = mk4(i2f(ref[0]), i2f(ref[1]), i2f(ref[2]), i2f(ref[3]));
}

}  // namespace nrcu
