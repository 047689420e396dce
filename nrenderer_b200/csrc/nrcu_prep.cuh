// nrcu_prep.cuh — scene preparation bodies (__host__ __device__): mesh world transform and
// flattening of spheres / triangles / planes / mesh triangles into the packed device records.
#pragma once
#include "nrcu_intersect.cuh"

namespace nrcu {

// ---------------------------------------------------------------------------------------------
// scene preparation
// ---------------------------------------------------------------------------------------------
// VertexTransformer::exec, MESH branch (acc_path_tracing/src/VertexTransformer.cpp:26-51):
// p <- M * (p,1), M = diag(600) + translation (40,-305,920); glm mat4*vec4 order (type_mat4x4.inl:561-571).
NR_HD void mesh_transform_vertex(float* pos, uint32_t vertex) {
    float* p = pos + 3 * (size_t)vertex;
    float x = p[0], y = p[1], z = p[2];
    const float s = 600.0f;
    p[0] = (s * x + 0.0f * y) + (0.0f * z + 40.0f * 1.0f);
    p[1] = (0.0f * x + s * y) + (0.0f * z + -305.0f * 1.0f);
    p[2] = (0.0f * x + 0.0f * y) + (s * z + 920.0f * 1.0f);
}

struct PrimSources {
    // per primitive: src_a = kind | entity << 2, src_b = triangle index inside the mesh (KIND_MESH)
    const uint32_t* src_a; const uint32_t* src_b;
    const float* sphere_position; const float* sphere_radius; const int* sphere_material;
    const float* triangle_vertices; const float* triangle_normal; const int* triangle_material;
    const float* plane_normal; const float* plane_position; const float* plane_u; const float* plane_v; const int* plane_material;
    const uint32_t* mesh_vertex_offset; const uint32_t* mesh_index_offset; const float* mesh_positions;
    const uint32_t* mesh_indices; const int* mesh_material;
};

// Flattening (SimplePathTracer.cpp:57-78, BVH.hpp:34-60) + Bounds3 constructors (Bounds3.hpp:35-103):
// one thread per output primitive writes its intersection record, shading record, leaf box and meta word
// (and, for nrcu_download_primitives, the 16-float world-space description when export16 != null).
// `box` is the reference's Bounds3 of the primitive (what its leaf gate tests); `bound` is a box that really
// contains the primitive (the reference's plane box can leave corners outside, Bounds3.hpp:52-78) and is
// what the BVH and the wide-primitive pre-test are built from.
NR_HD void build_prim(const PrimSources& ps, uint32_t i, int raycast, f4* geom, f4* shade, f4* box, f4* bound, uint32_t* meta, float* export16) {
    uint32_t a = ps.src_a[i], kind = a & 3u, e = a >> 2;
    int material;
    vec3 lo, hi, blo, bhi;
    bool own_bound = false;
    if (kind == KIND_SPHERE) {
        vec3 c = ld3(ps.sphere_position + 3 * e); float r = ps.sphere_radius[e];
        material = ps.sphere_material[e];
        geom[3 * i] = mk4(c.x, c.y, c.z, r); geom[3 * i + 1] = mk4(0, 0, 0, 0); geom[3 * i + 2] = mk4(0, 0, 0, 0);
        shade[i] = mk4(0, 0, 0, i2f((int)(((uint32_t)material << 2) | kind)));
        lo = c - r; hi = c + r;
        if (export16) { float* o = export16 + 16 * (size_t)i; for (int k = 0; k < 16; k++) o[k] = 0.f; st3(o, c); o[3] = r; }
    } else if (kind == KIND_PLANE) {
        vec3 n0 = ld3(ps.plane_normal + 3 * e), p = ld3(ps.plane_position + 3 * e), u = ld3(ps.plane_u + 3 * e), v = ld3(ps.plane_v + 3 * e);
        material = ps.plane_material[e];
        vec3 nn = raycast ? normalize(n0) : n0;   // ray_cast/.../intersections.cpp:55 normalises per test
        float r0[3], r1[3];
        quad_inverse_rows(u, v, r0, r1);
        geom[3 * i] = mk4(nn.x, nn.y, nn.z, p.x); geom[3 * i + 1] = mk4(p.y, p.z, r0[0], r0[1]); geom[3 * i + 2] = mk4(r0[2], r1[0], r1[1], r1[2]);
        shade[i] = mk4(nn.x, nn.y, nn.z, i2f((int)(((uint32_t)material << 2) | kind)));
        // Bounds3(Plane*), Bounds3.hpp:52-78 (uses the stored, un-normalised normal)
        vec3 p1 = p, p2 = p + u, p3 = p + v, p4 = p + u + v, en = 0.01f * n0;
        p1 = p1 - en; p2 = p2 - en; p3 = p3 + en; p4 = p4 + en;
        lo = vmin(p1, vmin(p2, vmin(p3, p4))); hi = vmax(p1, vmax(p2, vmax(p3, p4)));
        vec3 c2 = p + u, c3 = p + v, c4 = p + u + v;
        blo = vmin(lo, vmin(p, vmin(c2, vmin(c3, c4)))); bhi = vmax(hi, vmax(p, vmax(c2, vmax(c3, c4)))); own_bound = true;
        if (export16) { float* o = export16 + 16 * (size_t)i; for (int k = 0; k < 16; k++) o[k] = 0.f; st3(o, nn); st3(o + 3, p); st3(o + 6, u); st3(o + 9, v); }
    } else {
        vec3 v1, v2, v3, nrm;
        if (kind == KIND_TRIANGLE) {
            const float* t = ps.triangle_vertices + 9 * (size_t)e;
            v1 = ld3(t); v2 = ld3(t + 3); v3 = ld3(t + 6);
            nrm = ld3(ps.triangle_normal + 3 * e);
            material = ps.triangle_material[e];
        } else {
            const uint32_t* ix = ps.mesh_indices + ps.mesh_index_offset[e] + 3 * (size_t)ps.src_b[i];
            const float* base = ps.mesh_positions + 3 * (size_t)ps.mesh_vertex_offset[e];
            v1 = ld3(base + 3 * (size_t)ix[0]); v2 = ld3(base + 3 * (size_t)ix[1]); v3 = ld3(base + 3 * (size_t)ix[2]);
            nrm = normalize(cross(v2 - v1, v3 - v1));   // SimplePathTracer.cpp:72, Bounds3.hpp:96
            material = ps.mesh_material[e];
        }
        if (raycast) nrm = normalize(nrm);              // ray_cast/.../intersections.cpp:9
        vec3 e1 = v2 - v1, e2 = v3 - v1;
        geom[3 * i] = mk4(v1.x, v1.y, v1.z, e1.x); geom[3 * i + 1] = mk4(e1.y, e1.z, e2.x, e2.y); geom[3 * i + 2] = mk4(e2.z, 0.f, 0.f, 0.f);
        shade[i] = mk4(nrm.x, nrm.y, nrm.z, i2f((int)(((uint32_t)material << 2) | kind)));
        lo = vmin(v1, vmin(v2, v3)); hi = vmax(v1, vmax(v2, v3));
        if (export16) { float* o = export16 + 16 * (size_t)i; for (int k = 0; k < 16; k++) o[k] = 0.f; st3(o, v1); st3(o + 3, v2); st3(o + 6, v3); st3(o + 9, nrm); }
    }
    box[2 * i] = mk4(lo.x, lo.y, lo.z, 0.f); box[2 * i + 1] = mk4(hi.x, hi.y, hi.z, 0.f);
    if (!own_bound) { blo = lo; bhi = hi; }
    bound[2 * i] = mk4(blo.x, blo.y, blo.z, 0.f); bound[2 * i + 1] = mk4(bhi.x, bhi.y, bhi.z, 0.f);
    meta[i] = kind | ((uint32_t)material << 2);
}

}  // namespace nrcu
