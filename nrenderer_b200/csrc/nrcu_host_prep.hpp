// nrcu_host_prep.hpp — the per-scene scalars and small tables the host prepares before upload.
//
// Pure C++ (no CUDA): validation of the nrcu_scene, the sequential per-node translation of the
// explicit spheres / triangles / planes (VertexTransformer::exec,
// reference code/components/acc_path_tracing/src/VertexTransformer.cpp:6-25 — in place and in
// scene.nodes order, so an entity referenced twice moves twice, exactly like the reference), the
// primitive order of each mode, material defaults, area-light records, the Camera constructor
// (ray_cast/include/Camera.hpp:25-46) and Microfacet's constant Sampler(6) draws.  Everything
// per-vertex / per-primitive runs on the device (nrcu_prep.cuh).  Shared by nrcu_api.cu and by the
// CPU emulation under tests/host_emu, so both see identical inputs.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "nrcu.h"
#include "nrcu_intersect.cuh"

namespace nrcu {

struct HostPrep {
    std::vector<float> sph, tri, pln;        // translated copies of sphere_position / triangle_vertices / plane_position
    std::vector<uint32_t> src_a, src_b;      // primitive order: kind | entity << 2, mesh triangle index
    std::vector<uint32_t> mesh_nodes;        // entity of every MESH node in node order (one transform launch each)
    std::vector<DMaterial> materials;
    std::vector<f4> mat_head;                // DScene::mat_head
    std::vector<f4> lights;                  // NRCU_LIGHT_F4 float4 per area light
    DCamera cam;
    float mf_u1, mf_u2, mf_cos_phi, mf_sin_phi;
    float max_abs_coord;
    uint32_t total_vertices, total_indices;
};

// Returns an empty string on success, else the reason the scene is rejected.
inline std::string host_prepare(const nrcu_scene* sc, int mode, HostPrep& hp) {
    if (sc->width == 0 || sc->height == 0 || (uint64_t)sc->width * sc->height > 0x7fffffffull) return "bad resolution";
    for (uint32_t i = 0; i < sc->n_nodes; i++) {
        uint32_t t = sc->node_type[i], e = sc->node_entity[i];
        uint32_t lim = t == NRCU_NODE_SPHERE ? sc->n_spheres : t == NRCU_NODE_TRIANGLE ? sc->n_triangles : t == NRCU_NODE_PLANE ? sc->n_planes : sc->n_meshes;
        if (t > NRCU_NODE_MESH || e >= lim || sc->node_model[i] >= sc->n_models) return "node references a missing entity or model";
    }
    auto bad_mat = [&](const int32_t* m, uint32_t n) { for (uint32_t i = 0; i < n; i++) if (m[i] < 0 || (uint32_t)m[i] >= sc->n_materials) return true; return false; };
    // SceneBuilder::build returns nullptr when a node has no material (SceneBuilder.cpp:25-55)
    if (bad_mat(sc->sphere_material, sc->n_spheres) || bad_mat(sc->triangle_material, sc->n_triangles) ||
        bad_mat(sc->plane_material, sc->n_planes) || bad_mat(sc->mesh_material, sc->n_meshes)) return "an entity has no (valid) material";
    for (uint32_t m = 0; m < sc->n_meshes; m++) {
        uint32_t nv = sc->mesh_vertex_offset[m + 1] - sc->mesh_vertex_offset[m];
        for (uint32_t k = sc->mesh_index_offset[m]; k < sc->mesh_index_offset[m + 1]; k++)
            if (sc->mesh_indices[k] >= nv) return "mesh index out of range";
    }
    hp.total_vertices = sc->n_meshes ? sc->mesh_vertex_offset[sc->n_meshes] : 0;
    hp.total_indices = sc->n_meshes ? sc->mesh_index_offset[sc->n_meshes] : 0;

    // VertexTransformer::exec for explicit primitives.  glm: t*Vec4{v,1}, t = translate(I, tr):
    // (1*x + 0*y) + (0*z + tr.x*1)   (type_mat4x4.inl:561-571)
    hp.sph.assign(sc->sphere_position, sc->sphere_position + 3 * (size_t)sc->n_spheres);
    hp.tri.assign(sc->triangle_vertices, sc->triangle_vertices + 9 * (size_t)sc->n_triangles);
    hp.pln.assign(sc->plane_position, sc->plane_position + 3 * (size_t)sc->n_planes);
    auto xlate = [](float* v, const float* t) {
        float x = v[0], y = v[1], z = v[2];
        v[0] = (1.0f * x + 0.0f * y) + (0.0f * z + t[0] * 1.0f);
        v[1] = (0.0f * x + 1.0f * y) + (0.0f * z + t[1] * 1.0f);
        v[2] = (0.0f * x + 0.0f * y) + (1.0f * z + t[2] * 1.0f);
    };
    hp.mesh_nodes.clear();
    for (uint32_t i = 0; i < sc->n_nodes; i++) {
        const float* t = sc->model_translation + 3 * (size_t)sc->node_model[i];
        uint32_t e = sc->node_entity[i];
        if (sc->node_type[i] == NRCU_NODE_TRIANGLE) for (int k = 0; k < 3; k++) xlate(&hp.tri[9 * (size_t)e + 3 * k], t);
        else if (sc->node_type[i] == NRCU_NODE_SPHERE) xlate(&hp.sph[3 * (size_t)e], t);
        else if (sc->node_type[i] == NRCU_NODE_PLANE) xlate(&hp.pln[3 * (size_t)e], t);
        else if (mode != NRCU_MODE_RAYCAST) hp.mesh_nodes.push_back(e);   // RayCast's VertexTransformer ignores meshes
    }
    // primitive order
    hp.src_a.clear(); hp.src_b.clear();
    auto push_mesh = [&](uint32_t e) {
        uint32_t nt = (sc->mesh_index_offset[e + 1] - sc->mesh_index_offset[e]) / 3;
        for (uint32_t q = 0; q < nt; q++) { hp.src_a.push_back(KIND_MESH | (e << 2)); hp.src_b.push_back(q); }
    };
    if (mode == NRCU_MODE_ACC) {   // BVHNode::buildBounds, BVH.hpp:34-60
        for (uint32_t i = 0; i < sc->n_nodes; i++) {
            uint32_t t = sc->node_type[i], e = sc->node_entity[i];
            if (t == NRCU_NODE_MESH) push_mesh(e); else { hp.src_a.push_back(t | (e << 2)); hp.src_b.push_back(0); }
        }
    } else {                       // typed-buffer loops, RayCastRenderer.cpp:66-91 / SimplePathTracer.cpp:57-78,104-129
        for (uint32_t e = 0; e < sc->n_spheres; e++) { hp.src_a.push_back(KIND_SPHERE | (e << 2)); hp.src_b.push_back(0); }
        for (uint32_t e = 0; e < sc->n_triangles; e++) { hp.src_a.push_back(KIND_TRIANGLE | (e << 2)); hp.src_b.push_back(0); }
        if (mode == NRCU_MODE_SIMPLE)
            for (uint32_t i = 0; i < sc->n_nodes; i++) if (sc->node_type[i] == NRCU_NODE_MESH) push_mesh(sc->node_entity[i]);
        for (uint32_t e = 0; e < sc->n_planes; e++) { hp.src_a.push_back(KIND_PLANE | (e << 2)); hp.src_b.push_back(0); }
    }
    if (hp.src_a.size() >= (1u << 27)) return "too many primitives";

    // materials with the shader-constructor defaults (Lambertian.cpp:8-14, Phong.cpp:8-23, Microfacet.cpp:151-166;
    // Conductor.hpp:17-26 / Glass.hpp:16-22 leave absent members uninitialised: zero here)
    static_assert(sizeof(DMaterial) == sizeof(nrcu_material), "material layouts must agree");
    hp.materials.assign(std::max(sc->n_materials, 1u), DMaterial{});
    for (uint32_t i = 0; i < sc->n_materials; i++) {
        std::memcpy(&hp.materials[i], &sc->materials[i], sizeof(DMaterial));
        DMaterial& m = hp.materials[i];
        if (!(m.present & NRCU_MP_DIFFUSE_COLOR)) m.diffuse_color[0] = m.diffuse_color[1] = m.diffuse_color[2] = 1.f;
        if (!(m.present & NRCU_MP_SPECULAR_COLOR)) m.specular_color[0] = m.specular_color[1] = m.specular_color[2] = 1.f;
        if (!(m.present & NRCU_MP_SPECULAR_EX)) m.specular_ex = 1.f;
        if (m.type == 3) {
            if (!(m.present & NRCU_MP_ALBEDO)) m.albedo[0] = m.albedo[1] = m.albedo[2] = 1.f;
            if (!(m.present & NRCU_MP_ROUGHNESS)) m.roughness = 0.2f;
            if (!(m.present & NRCU_MP_F0)) m.f0 = 0.04;
        }
    }
    hp.mat_head.resize(hp.materials.size());
    for (size_t i = 0; i < hp.materials.size(); i++) {
        const DMaterial& m = hp.materials[i];
        const float pi = 3.1415926535898f;   // acc_path_tracing/include/shaders/Shader.hpp:17 (NRCU_PT_PI)
        hp.mat_head[i] = mk4(m.diffuse_color[0] / pi, m.diffuse_color[1] / pi, m.diffuse_color[2] / pi, i2f((int)m.type));
    }
    // area lights: quad record with n = cross(u,v) (xAreaLight, intersections.cpp:74-93) + radiance
    hp.lights.assign(std::max(sc->n_area_lights, 1u) * NRCU_LIGHT_F4, mk4(0, 0, 0, 0));
    for (uint32_t i = 0; i < sc->n_area_lights; i++) {
        vec3 u = ld3(sc->area_u + 3 * i), v = ld3(sc->area_v + 3 * i), p = ld3(sc->area_position + 3 * i), nn = cross(u, v);
        float r0[3], r1[3];
        quad_inverse_rows(u, v, r0, r1);
        f4* Lr = &hp.lights[NRCU_LIGHT_F4 * (size_t)i];
        Lr[0] = mk4(nn.x, nn.y, nn.z, p.x); Lr[1] = mk4(p.y, p.z, r0[0], r0[1]);
        Lr[2] = mk4(r0[2], r1[0], r1[1], r1[2]);
        Lr[3] = mk4(sc->area_radiance[3 * i], sc->area_radiance[3 * i + 1], sc->area_radiance[3 * i + 2], 0.f);
        Lr[4] = mk4(u.x, u.y, u.z, 0.f); Lr[5] = mk4(v.x, v.y, v.z, 0.f);   // edges, for light sampling (NEE extension)
    }
    // Camera ctor (ray_cast/include/Camera.hpp:25-46): same libm tanf as the reference
    {
        DCamera& c = hp.cam;
        c.position = ld3(sc->cam_position);
        c.lens_radius = sc->cam_aperture / 2.f;
        float vfov = sc->cam_fov;
        if (vfov > 160.f) vfov = 160.f; else if (vfov < 20.f) vfov = 20.f;   // clamp(x, max, min), geometry/vec.hpp:86-91
        float theta = vfov * 0.01745329251994329576923690768489f;           // glm::radians
        float half_h = tanf(theta / 2.f), half_w = sc->cam_aspect * half_h;
        vec3 w = normalize(c.position - ld3(sc->cam_look_at));
        c.u = normalize(cross(ld3(sc->cam_up), w));
        c.v = cross(w, c.u);
        float f = sc->cam_focus_distance;
        c.lower_left = c.position - (half_w * f) * c.u - (half_h * f) * c.v - f * w;
        c.horizontal = (2 * half_w * f) * c.u;
        c.vertical = (2 * half_h * f) * c.v;
    }
    {   // Sampler(6): minstd_rand seeded with 6, two uniform_real_distribution<float>(0,1) draws (Microfacet.cpp:65-70)
        uint64_t x = 6; float range = (float)2147483646.0L;
        x = (48271u * x) % 2147483647u; hp.mf_u1 = (float)(uint32_t)(x - 1) / range;
        x = (48271u * x) % 2147483647u; hp.mf_u2 = (float)(uint32_t)(x - 1) / range;
        float phi = 2.0f * 3.1415926535898f * hp.mf_u2;
        hp.mf_cos_phi = cosf(phi); hp.mf_sin_phi = sinf(phi);
    }
    // largest coordinate magnitude: sizes the conservative padding of the BVH boxes
    float max_abs = 1.f;
    for (int k = 0; k < 3; k++) max_abs = std::max(max_abs, std::fabs(sc->cam_position[k]));
    for (uint32_t i = 0; i < sc->n_spheres; i++) for (int k = 0; k < 3; k++) max_abs = std::max(max_abs, std::fabs(hp.sph[3 * i + k]) + sc->sphere_radius[i]);
    for (float v : hp.tri) max_abs = std::max(max_abs, std::fabs(v));
    for (uint32_t i = 0; i < sc->n_planes; i++) for (int k = 0; k < 3; k++)
        max_abs = std::max(max_abs, std::fabs(hp.pln[3 * i + k]) + std::fabs(sc->plane_u[3 * i + k]) + std::fabs(sc->plane_v[3 * i + k]));
    if (mode != NRCU_MODE_RAYCAST)
        for (uint32_t i = 0; i < 3 * hp.total_vertices; i++) max_abs = std::max(max_abs, 600.f * std::fabs(sc->mesh_positions[i]) + 920.f);
    hp.max_abs_coord = max_abs;
    return "";
}

// Fills the pointer-free part of the device scene descriptor.
inline void fill_scene_scalars(DScene& ds, const nrcu_scene* sc, int mode, const HostPrep& hp) {
    std::memset(&ds, 0, sizeof(ds));
    ds.mode = mode; ds.width = sc->width; ds.height = sc->height; ds.depth = sc->depth;
    ds.n_prims = (uint32_t)hp.src_a.size();
    ds.root_ref = NRCU_REF_EMPTY;
    ds.n_materials = sc->n_materials;
    ds.n_area_lights = sc->n_area_lights;
    ds.n_point_lights = sc->n_point_lights;
    if (sc->n_point_lights) { ds.point_position = ld3(sc->point_position); ds.point_intensity = ld3(sc->point_intensity); }
    ds.cam = hp.cam;
    ds.ambient = ld3(sc->ambient_constant);
    ds.mf_u1 = hp.mf_u1; ds.mf_u2 = hp.mf_u2; ds.mf_cos_phi = hp.mf_cos_phi; ds.mf_sin_phi = hp.mf_sin_phi;
}

}  // namespace nrcu
