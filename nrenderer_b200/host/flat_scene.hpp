// flat_scene.hpp — owning, reference-header-free mirror of `nrcu_scene` plus its on-disk form.
//
// A FlatScene is the POD restatement of NRenderer::Scene (reference
// code/include/scene/Scene.hpp:40-67) in model-local coordinates.  It is what the plugin
// adapter hands to libnrcuda.so, what the headless harness can dump/load (`.nrsc` files,
// used as test fixtures on machines where /root/reference is absent) and what the C oracle
// reads.  File layout: 8-byte magic "NRSC0001", then records
//   u32 name_len | name bytes | u32 dtype (0=f32,1=u32,2=i32,3=u64) | u64 count | payload
// until EOF.  Little endian.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "nrcu.h"

namespace nrb200 {

struct FlatScene {
    uint32_t width = 500, height = 500, depth = 4, samples_per_pixel = 16;
    float cam_position[3] = {0, 0, 10}, cam_up[3] = {0, 1, 0}, cam_look_at[3] = {0, 0, 1000};
    float cam_fov = 40, cam_aperture = 0, cam_focus_distance = 0.1f, cam_aspect = 1;
    uint32_t ambient_type = 0;
    float ambient_constant[3] = {0, 0, 0};
    int32_t ambient_environment_map = -1;

    std::vector<float> model_translation;
    std::vector<uint32_t> node_type, node_entity, node_model;
    std::vector<float> sphere_position, sphere_radius;
    std::vector<int32_t> sphere_material;
    std::vector<float> triangle_vertices, triangle_normal;
    std::vector<int32_t> triangle_material;
    std::vector<float> plane_normal, plane_position, plane_u, plane_v;
    std::vector<int32_t> plane_material;
    std::vector<uint32_t> mesh_vertex_offset{0}, mesh_index_offset{0};
    std::vector<float> mesh_positions;
    std::vector<uint32_t> mesh_indices;
    std::vector<int32_t> mesh_material;
    std::vector<nrcu_material> materials;
    std::vector<float> point_intensity, point_position;
    std::vector<float> area_radiance, area_position, area_u, area_v;
    std::vector<uint32_t> texture_width, texture_height;
    std::vector<uint64_t> texture_offset;
    std::vector<float> texture_rgba;

    // Non-owning C view; valid while *this is alive and unmodified.
    nrcu_scene view() const {
        nrcu_scene s;
        std::memset(&s, 0, sizeof(s));
        s.width = width; s.height = height; s.depth = depth; s.samples_per_pixel = samples_per_pixel;
        for (int i = 0; i < 3; i++) {
            s.cam_position[i] = cam_position[i]; s.cam_up[i] = cam_up[i]; s.cam_look_at[i] = cam_look_at[i];
            s.ambient_constant[i] = ambient_constant[i];
        }
        s.cam_fov = cam_fov; s.cam_aperture = cam_aperture;
        s.cam_focus_distance = cam_focus_distance; s.cam_aspect = cam_aspect;
        s.ambient_type = ambient_type; s.ambient_environment_map = ambient_environment_map;
        s.n_models = (uint32_t)(model_translation.size() / 3); s.model_translation = model_translation.data();
        s.n_nodes = (uint32_t)node_type.size();
        s.node_type = node_type.data(); s.node_entity = node_entity.data(); s.node_model = node_model.data();
        s.n_spheres = (uint32_t)sphere_radius.size();
        s.sphere_position = sphere_position.data(); s.sphere_radius = sphere_radius.data();
        s.sphere_material = sphere_material.data();
        s.n_triangles = (uint32_t)triangle_material.size();
        s.triangle_vertices = triangle_vertices.data(); s.triangle_normal = triangle_normal.data();
        s.triangle_material = triangle_material.data();
        s.n_planes = (uint32_t)plane_material.size();
        s.plane_normal = plane_normal.data(); s.plane_position = plane_position.data();
        s.plane_u = plane_u.data(); s.plane_v = plane_v.data(); s.plane_material = plane_material.data();
        s.n_meshes = (uint32_t)mesh_material.size();
        s.mesh_vertex_offset = mesh_vertex_offset.data(); s.mesh_index_offset = mesh_index_offset.data();
        s.mesh_positions = mesh_positions.data(); s.mesh_indices = mesh_indices.data();
        s.mesh_material = mesh_material.data();
        s.n_materials = (uint32_t)materials.size(); s.materials = materials.data();
        s.n_point_lights = (uint32_t)(point_position.size() / 3);
        s.point_intensity = point_intensity.data(); s.point_position = point_position.data();
        s.n_area_lights = (uint32_t)(area_position.size() / 3);
        s.area_radiance = area_radiance.data(); s.area_position = area_position.data();
        s.area_u = area_u.data(); s.area_v = area_v.data();
        s.n_textures = (uint32_t)texture_width.size();
        s.texture_width = texture_width.data(); s.texture_height = texture_height.data();
        s.texture_offset = texture_offset.data(); s.texture_rgba = texture_rgba.data();
        return s;
    }
};

namespace detail {
template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr uint32_t v = 0; };
template <> struct dtype_of<uint32_t> { static constexpr uint32_t v = 1; };
template <> struct dtype_of<int32_t> { static constexpr uint32_t v = 2; };
template <> struct dtype_of<uint64_t> { static constexpr uint32_t v = 3; };

template <typename T>
inline void put(FILE* f, const char* name, const T* p, size_t n) {
    uint32_t nl = (uint32_t)std::strlen(name), dt = dtype_of<T>::v;
    uint64_t cnt = n;
    std::fwrite(&nl, 4, 1, f); std::fwrite(name, 1, nl, f);
    std::fwrite(&dt, 4, 1, f); std::fwrite(&cnt, 8, 1, f);
    if (n) std::fwrite(p, sizeof(T), n, f);
}
template <typename T>
inline void put(FILE* f, const char* name, const std::vector<T>& v) { put(f, name, v.data(), v.size()); }

struct Record { uint32_t dtype; std::vector<char> bytes; };
template <typename T>
inline void get(const std::map<std::string, Record>& m, const char* name, std::vector<T>& out) {
    auto it = m.find(name);
    if (it == m.end()) return;
    if (it->second.dtype != dtype_of<T>::v) throw std::runtime_error(std::string("nrsc: dtype mismatch for ") + name);
    out.resize(it->second.bytes.size() / sizeof(T));
    if (!out.empty()) std::memcpy(out.data(), it->second.bytes.data(), out.size() * sizeof(T));
}
template <typename T>
inline void get_fixed(const std::map<std::string, Record>& m, const char* name, T* out, size_t n) {
    std::vector<T> v; get(m, name, v);
    if (v.empty()) return;
    if (v.size() != n) throw std::runtime_error(std::string("nrsc: bad length for ") + name);
    std::memcpy(out, v.data(), n * sizeof(T));
}
}  // namespace detail

inline void save_flat_scene(const FlatScene& s, const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    std::fwrite("NRSC0001", 1, 8, f);
    using detail::put;
    uint32_t opt[4] = {s.width, s.height, s.depth, s.samples_per_pixel};
    put(f, "render_option", opt, 4);
    float cam[13] = {s.cam_position[0], s.cam_position[1], s.cam_position[2], s.cam_up[0], s.cam_up[1], s.cam_up[2],
                     s.cam_look_at[0], s.cam_look_at[1], s.cam_look_at[2],
                     s.cam_fov, s.cam_aperture, s.cam_focus_distance, s.cam_aspect};
    put(f, "camera", cam, 13);
    put(f, "ambient_type", &s.ambient_type, 1);
    put(f, "ambient_constant", s.ambient_constant, 3);
    put(f, "ambient_environment_map", &s.ambient_environment_map, 1);
    put(f, "model_translation", s.model_translation);
    put(f, "node_type", s.node_type); put(f, "node_entity", s.node_entity); put(f, "node_model", s.node_model);
    put(f, "sphere_position", s.sphere_position); put(f, "sphere_radius", s.sphere_radius);
    put(f, "sphere_material", s.sphere_material);
    put(f, "triangle_vertices", s.triangle_vertices); put(f, "triangle_normal", s.triangle_normal);
    put(f, "triangle_material", s.triangle_material);
    put(f, "plane_normal", s.plane_normal); put(f, "plane_position", s.plane_position);
    put(f, "plane_u", s.plane_u); put(f, "plane_v", s.plane_v); put(f, "plane_material", s.plane_material);
    put(f, "mesh_vertex_offset", s.mesh_vertex_offset); put(f, "mesh_index_offset", s.mesh_index_offset);
    put(f, "mesh_positions", s.mesh_positions); put(f, "mesh_indices", s.mesh_indices);
    put(f, "mesh_material", s.mesh_material);
    // materials: u32 (type, present) pairs + 22 floats each
    std::vector<uint32_t> mt; std::vector<float> mf;
    for (auto& m : s.materials) {
        mt.push_back(m.type); mt.push_back(m.present);
        const float* p = m.diffuse_color;
        mf.insert(mf.end(), p, p + 22);
    }
    put(f, "material_type_present", mt); put(f, "material_params", mf);
    put(f, "point_intensity", s.point_intensity); put(f, "point_position", s.point_position);
    put(f, "area_radiance", s.area_radiance); put(f, "area_position", s.area_position);
    put(f, "area_u", s.area_u); put(f, "area_v", s.area_v);
    put(f, "texture_width", s.texture_width); put(f, "texture_height", s.texture_height);
    put(f, "texture_offset", s.texture_offset); put(f, "texture_rgba", s.texture_rgba);
    std::fclose(f);
}

inline FlatScene load_flat_scene(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot read " + path);
    char magic[8];
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "NRSC0001", 8) != 0) {
        std::fclose(f); throw std::runtime_error("not an NRSC0001 file: " + path);
    }
    std::map<std::string, detail::Record> m;
    for (;;) {
        uint32_t nl;
        if (std::fread(&nl, 4, 1, f) != 1) break;
        std::string name(nl, '\0');
        uint32_t dt; uint64_t cnt;
        if (std::fread(name.data(), 1, nl, f) != nl || std::fread(&dt, 4, 1, f) != 1 || std::fread(&cnt, 8, 1, f) != 1) {
            std::fclose(f); throw std::runtime_error("truncated record in " + path);
        }
        size_t esz = dt == 3 ? 8 : 4;
        detail::Record r; r.dtype = dt; r.bytes.resize(cnt * esz);
        if (cnt && std::fread(r.bytes.data(), esz, cnt, f) != cnt) {
            std::fclose(f); throw std::runtime_error("truncated payload in " + path);
        }
        m[name] = std::move(r);
    }
    std::fclose(f);
    FlatScene s;
    using detail::get; using detail::get_fixed;
    uint32_t opt[4] = {s.width, s.height, s.depth, s.samples_per_pixel};
    get_fixed(m, "render_option", opt, 4);
    s.width = opt[0]; s.height = opt[1]; s.depth = opt[2]; s.samples_per_pixel = opt[3];
    float cam[13];
    std::vector<float> camv; get(m, "camera", camv);
    if (camv.size() == 13) {
        std::memcpy(cam, camv.data(), sizeof(cam));
        for (int i = 0; i < 3; i++) { s.cam_position[i] = cam[i]; s.cam_up[i] = cam[3 + i]; s.cam_look_at[i] = cam[6 + i]; }
        s.cam_fov = cam[9]; s.cam_aperture = cam[10]; s.cam_focus_distance = cam[11]; s.cam_aspect = cam[12];
    }
    get_fixed(m, "ambient_type", &s.ambient_type, 1);
    get_fixed(m, "ambient_constant", s.ambient_constant, 3);
    get_fixed(m, "ambient_environment_map", &s.ambient_environment_map, 1);
    get(m, "model_translation", s.model_translation);
    get(m, "node_type", s.node_type); get(m, "node_entity", s.node_entity); get(m, "node_model", s.node_model);
    get(m, "sphere_position", s.sphere_position); get(m, "sphere_radius", s.sphere_radius);
    get(m, "sphere_material", s.sphere_material);
    get(m, "triangle_vertices", s.triangle_vertices); get(m, "triangle_normal", s.triangle_normal);
    get(m, "triangle_material", s.triangle_material);
    get(m, "plane_normal", s.plane_normal); get(m, "plane_position", s.plane_position);
    get(m, "plane_u", s.plane_u); get(m, "plane_v", s.plane_v); get(m, "plane_material", s.plane_material);
    get(m, "mesh_vertex_offset", s.mesh_vertex_offset); get(m, "mesh_index_offset", s.mesh_index_offset);
    if (s.mesh_vertex_offset.empty()) s.mesh_vertex_offset = {0};
    if (s.mesh_index_offset.empty()) s.mesh_index_offset = {0};
    get(m, "mesh_positions", s.mesh_positions); get(m, "mesh_indices", s.mesh_indices);
    get(m, "mesh_material", s.mesh_material);
    std::vector<uint32_t> mt; std::vector<float> mf;
    get(m, "material_type_present", mt); get(m, "material_params", mf);
    for (size_t i = 0; i * 2 < mt.size(); i++) {
        nrcu_material mm; std::memset(&mm, 0, sizeof(mm));
        mm.type = mt[2 * i]; mm.present = mt[2 * i + 1];
        std::memcpy(mm.diffuse_color, &mf[22 * i], 22 * sizeof(float));
        s.materials.push_back(mm);
    }
    get(m, "point_intensity", s.point_intensity); get(m, "point_position", s.point_position);
    get(m, "area_radiance", s.area_radiance); get(m, "area_position", s.area_position);
    get(m, "area_u", s.area_u); get(m, "area_v", s.area_v);
    get(m, "texture_width", s.texture_width); get(m, "texture_height", s.texture_height);
    get(m, "texture_offset", s.texture_offset); get(m, "texture_rgba", s.texture_rgba);
    return s;
}

}  // namespace nrb200
