// scene_bridge.hpp — NRenderer::Scene <-> nrb200::FlatScene.
//
// Compiled only by translation units that include the reference's own headers
// (code/include/scene/*.hpp): the plugin adapter and the headless harness.  `flatten`
// is the "scene upload" front half: it reads the Scene a RenderComponent receives
// (reference code/include/scene/Scene.hpp:40-67) WITHOUT mutating it — the reference
// components transform it in place (VertexTransformer.cpp:6-54), the library does that on the
// device instead.  `unflatten` rebuilds a Scene from a fixture so the reference CPU
// components can be run on machines that have no .scn/.obj files.
#pragma once
#include <string>

#include "scene/Scene.hpp"
#include "../host/flat_scene.hpp"

namespace nrb200 {

inline void put3(std::vector<float>& v, const NRenderer::Vec3& a) { v.push_back(a.x); v.push_back(a.y); v.push_back(a.z); }
inline NRenderer::Vec3 get3(const std::vector<float>& v, size_t i) { return {v[3 * i], v[3 * i + 1], v[3 * i + 2]}; }
inline int32_t mat_index(const NRenderer::Handle& h) { return h.valid() ? (int32_t)h.index() : -1; }
inline NRenderer::Handle mat_handle(int32_t i) { return i < 0 ? NRenderer::Handle{} : NRenderer::Handle{(unsigned int)i}; }

// Property resolution as the reference shaders do it: first property whose key matches
// (Material::getProperty, Material.hpp:113-127).  A key present with a different variant type would
// throw std::bad_variant_access in the reference; here it is reported through `warn` and ignored.
inline nrcu_material flatten_material(const NRenderer::Material& m, std::string* warn) {
    using P = NRenderer::Property;
    nrcu_material o;
    std::memset(&o, 0, sizeof(o));
    o.type = m.type;
    auto find = [&](const char* key, P::Type want) -> const P* {
        for (auto& p : m.properties) {
            if (p.key == key) {
                if (p.type == want) return &p;
                if (warn) *warn += std::string("property '") + key + "' has an unexpected type; ignored. ";
                return nullptr;
            }
        }
        return nullptr;
    };
    auto rgb = [&](const char* key, uint32_t bit, float* dst) {
        if (auto* p = find(key, P::Type::RGB)) {
            auto v = std::get<P::Wrapper::RGBType>(p->valueWrapper).value;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; o.present |= bit;
        }
    };
    auto vec3 = [&](const char* key, uint32_t bit, float* dst) {
        if (auto* p = find(key, P::Type::VEC3)) {
            auto v = std::get<P::Wrapper::Vec3Type>(p->valueWrapper).value;
            dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; o.present |= bit;
        }
    };
    auto flt = [&](const char* key, uint32_t bit, float* dst) {
        if (auto* p = find(key, P::Type::FLOAT)) {
            *dst = std::get<P::Wrapper::FloatType>(p->valueWrapper).value; o.present |= bit;
        }
    };
    rgb("diffuseColor", NRCU_MP_DIFFUSE_COLOR, o.diffuse_color);
    rgb("specularColor", NRCU_MP_SPECULAR_COLOR, o.specular_color);
    flt("specularEx", NRCU_MP_SPECULAR_EX, &o.specular_ex);
    rgb("albedo", NRCU_MP_ALBEDO, o.albedo);
    vec3("eta_r", NRCU_MP_ETA_R, o.eta_r);
    vec3("eta_i", NRCU_MP_ETA_I, o.eta_i);
    flt("ior", NRCU_MP_IOR, &o.ior);
    rgb("absorbed", NRCU_MP_ABSORBED, o.absorbed);
    flt("roughness", NRCU_MP_ROUGHNESS, &o.roughness);
    flt("F0", NRCU_MP_F0, &o.f0);
    return o;
}

inline FlatScene flatten(const NRenderer::Scene& sc, std::string* warn = nullptr) {
    FlatScene f;
    f.width = sc.renderOption.width; f.height = sc.renderOption.height;
    f.depth = sc.renderOption.depth; f.samples_per_pixel = sc.renderOption.samplesPerPixel;
    for (int i = 0; i < 3; i++) {
        f.cam_position[i] = sc.camera.position[i]; f.cam_up[i] = sc.camera.up[i];
        f.cam_look_at[i] = sc.camera.lookAt[i]; f.ambient_constant[i] = sc.ambient.constant[i];
    }
    f.cam_fov = sc.camera.fov; f.cam_aperture = sc.camera.aperture;
    f.cam_focus_distance = sc.camera.focusDistance; f.cam_aspect = sc.camera.aspect;
    f.ambient_type = sc.ambient.type == NRenderer::Ambient::Type::CONSTANT ? NRCU_AMBIENT_CONSTANT
                                                                          : NRCU_AMBIENT_ENVIRONMENT_MAP;
    f.ambient_environment_map = mat_index(sc.ambient.environmentMap);
    for (auto& m : sc.models) put3(f.model_translation, m.translation);
    for (auto& n : sc.nodes) {
        f.node_type.push_back((uint32_t)n.type); f.node_entity.push_back(n.entity); f.node_model.push_back(n.model);
    }
    for (auto& s : sc.sphereBuffer) {
        put3(f.sphere_position, s.position); f.sphere_radius.push_back(s.radius);
        f.sphere_material.push_back(mat_index(s.material));
    }
    for (auto& t : sc.triangleBuffer) {
        put3(f.triangle_vertices, t.v1); put3(f.triangle_vertices, t.v2); put3(f.triangle_vertices, t.v3);
        put3(f.triangle_normal, t.normal); f.triangle_material.push_back(mat_index(t.material));
    }
    for (auto& p : sc.planeBuffer) {
        put3(f.plane_normal, p.normal); put3(f.plane_position, p.position);
        put3(f.plane_u, p.u); put3(f.plane_v, p.v); f.plane_material.push_back(mat_index(p.material));
    }
    for (auto& m : sc.meshBuffer) {
        for (auto& p : m.positions) put3(f.mesh_positions, p);
        for (auto i : m.positionIndices) f.mesh_indices.push_back(i);
        f.mesh_vertex_offset.push_back((uint32_t)(f.mesh_positions.size() / 3));
        f.mesh_index_offset.push_back((uint32_t)f.mesh_indices.size());
        f.mesh_material.push_back(mat_index(m.material));
    }
    for (auto& m : sc.materials) f.materials.push_back(flatten_material(m, warn));
    for (auto& l : sc.pointLightBuffer) { put3(f.point_intensity, l.intensity); put3(f.point_position, l.position); }
    for (auto& l : sc.areaLightBuffer) {
        put3(f.area_radiance, l.radiance); put3(f.area_position, l.position); put3(f.area_u, l.u); put3(f.area_v, l.v);
    }
    for (auto& t : sc.textures) {
        f.texture_width.push_back(t.width); f.texture_height.push_back(t.height);
        f.texture_offset.push_back(f.texture_rgba.size());
        size_t n = (size_t)t.width * t.height;
        for (size_t i = 0; i < n; i++)
            for (int c = 0; c < 4; c++) f.texture_rgba.push_back(t.rgba ? t.rgba[i][c] : 0.f);
    }
    return f;
}

// Inverse of `flatten` (lights of the unused kinds and per-node names are not represented).
inline NRenderer::SharedScene unflatten(const FlatScene& f) {
    using namespace NRenderer;
    using PW = Property::Wrapper;
    auto sp = std::make_shared<Scene>();
    Scene& sc = *sp;
    sc.renderOption.width = f.width; sc.renderOption.height = f.height;
    sc.renderOption.depth = f.depth; sc.renderOption.samplesPerPixel = f.samples_per_pixel;
    sc.camera = Camera{Vec3{f.cam_position[0], f.cam_position[1], f.cam_position[2]},
                       Vec3{f.cam_up[0], f.cam_up[1], f.cam_up[2]},
                       Vec3{f.cam_look_at[0], f.cam_look_at[1], f.cam_look_at[2]},
                       f.cam_fov, f.cam_aperture, f.cam_focus_distance, f.cam_aspect};
    sc.ambient.type = f.ambient_type == NRCU_AMBIENT_CONSTANT ? Ambient::Type::CONSTANT : Ambient::Type::ENVIROMENT_MAP;
    sc.ambient.constant = {f.ambient_constant[0], f.ambient_constant[1], f.ambient_constant[2]};
    sc.ambient.environmentMap = mat_handle(f.ambient_environment_map);
    for (size_t i = 0; i < f.model_translation.size() / 3; i++) {
        Model m; m.translation = get3(f.model_translation, i); sc.models.push_back(m);
    }
    for (size_t i = 0; i < f.node_type.size(); i++) {
        Node n; n.type = (Node::Type)f.node_type[i]; n.entity = f.node_entity[i]; n.model = f.node_model[i];
        if (n.model < sc.models.size()) sc.models[n.model].nodes.push_back((Index)i);
        sc.nodes.push_back(n);
    }
    for (size_t i = 0; i < f.sphere_radius.size(); i++) {
        Sphere s; s.position = get3(f.sphere_position, i); s.radius = f.sphere_radius[i];
        s.material = mat_handle(f.sphere_material[i]); sc.sphereBuffer.push_back(s);
    }
    for (size_t i = 0; i < f.triangle_material.size(); i++) {
        Triangle t; t.v1 = get3(f.triangle_vertices, 3 * i); t.v2 = get3(f.triangle_vertices, 3 * i + 1);
        t.v3 = get3(f.triangle_vertices, 3 * i + 2); t.normal = get3(f.triangle_normal, i);
        t.material = mat_handle(f.triangle_material[i]); sc.triangleBuffer.push_back(t);
    }
    for (size_t i = 0; i < f.plane_material.size(); i++) {
        Plane p; p.normal = get3(f.plane_normal, i); p.position = get3(f.plane_position, i);
        p.u = get3(f.plane_u, i); p.v = get3(f.plane_v, i); p.material = mat_handle(f.plane_material[i]);
        sc.planeBuffer.push_back(p);
    }
    for (size_t i = 0; i < f.mesh_material.size(); i++) {
        Mesh m;
        for (uint32_t v = f.mesh_vertex_offset[i]; v < f.mesh_vertex_offset[i + 1]; v++) m.positions.push_back(get3(f.mesh_positions, v));
        for (uint32_t k = f.mesh_index_offset[i]; k < f.mesh_index_offset[i + 1]; k++) m.positionIndices.push_back(f.mesh_indices[k]);
        m.material = mat_handle(f.mesh_material[i]); sc.meshBuffer.push_back(m);
    }
    for (auto& fm : f.materials) {
        Material m; m.type = fm.type;
        auto v3 = [](const float* p) { return Vec3{p[0], p[1], p[2]}; };
        if (fm.present & NRCU_MP_DIFFUSE_COLOR) m.registerProperty("diffuseColor", PW::RGBType{v3(fm.diffuse_color)});
        if (fm.present & NRCU_MP_SPECULAR_COLOR) m.registerProperty("specularColor", PW::RGBType{v3(fm.specular_color)});
        if (fm.present & NRCU_MP_SPECULAR_EX) m.registerProperty("specularEx", PW::FloatType{fm.specular_ex});
        if (fm.present & NRCU_MP_ALBEDO) m.registerProperty("albedo", PW::RGBType{v3(fm.albedo)});
        if (fm.present & NRCU_MP_ETA_R) m.registerProperty("eta_r", PW::Vec3Type{v3(fm.eta_r)});
        if (fm.present & NRCU_MP_ETA_I) m.registerProperty("eta_i", PW::Vec3Type{v3(fm.eta_i)});
        if (fm.present & NRCU_MP_IOR) m.registerProperty("ior", PW::FloatType{fm.ior});
        if (fm.present & NRCU_MP_ABSORBED) m.registerProperty("absorbed", PW::RGBType{v3(fm.absorbed)});
        if (fm.present & NRCU_MP_ROUGHNESS) m.registerProperty("roughness", PW::FloatType{fm.roughness});
        if (fm.present & NRCU_MP_F0) m.registerProperty("F0", PW::FloatType{fm.f0});
        sc.materials.push_back(m);
    }
    for (size_t i = 0; i < f.point_position.size() / 3; i++) {
        PointLight l; l.intensity = get3(f.point_intensity, i); l.position = get3(f.point_position, i);
        Light li{Light::Type::POINT}; li.entity = (Index)sc.pointLightBuffer.size();
        sc.lights.push_back(li); sc.pointLightBuffer.push_back(l);
    }
    for (size_t i = 0; i < f.area_position.size() / 3; i++) {
        AreaLight l; l.radiance = get3(f.area_radiance, i); l.position = get3(f.area_position, i);
        l.u = get3(f.area_u, i); l.v = get3(f.area_v, i);
        Light li{Light::Type::AREA}; li.entity = (Index)sc.areaLightBuffer.size();
        sc.lights.push_back(li); sc.areaLightBuffer.push_back(l);
    }
    for (size_t i = 0; i < f.texture_width.size(); i++) {
        Texture t; t.width = f.texture_width[i]; t.height = f.texture_height[i];
        size_t n = (size_t)t.width * t.height;
        t.rgba = new RGBA[n];
        const float* src = f.texture_rgba.data() + f.texture_offset[i];
        for (size_t k = 0; k < n; k++) t.rgba[k] = RGBA{src[4 * k], src[4 * k + 1], src[4 * k + 2], src[4 * k + 3]};
        sc.textures.push_back(std::move(t));
    }
    return sp;
}

}  // namespace nrb200
