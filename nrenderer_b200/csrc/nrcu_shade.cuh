// nrcu_shade.cuh — camera rays, material shading and the per-bounce path update.
//
// Device restatement (also compiled for the host emulation in tests/host_emu) of
//   Camera::shoot            ray_cast/include/Camera.hpp:49-57, acc_path_tracing/include/Camera.hpp:51-63
//   RayCastRenderer::trace   ray_cast/src/RayCastRenderer.cpp:40-64  + shaders/{Lambertian,Phong}.cpp
//   *PathTracerRenderer::trace  simple_path_tracing/src/SimplePathTracer.cpp:144-177,
//                               acc_path_tracing/src/AccPathTracer.cpp:121-181
//   Lambertian / Conductor / Glass / Microfacet ::shade  acc_path_tracing/src/shaders/*.cpp
//   samplers                 acc_path_tracing/include/samplers/{UniformInSquare,UniformInCircle,Hemisphere}.hpp
// (paths relative to /root/reference/code/components).  The recursion of trace() is a product of
// per-bounce factors, so it is run forward: throughput *= factor, radiance added when the path ends.
#pragma once
#include <float.h>
#include "nrcu_intersect.cuh"

namespace nrcu {

#define NRCU_PT_PI 3.1415926535898f   // acc_path_tracing/include/shaders/Shader.hpp:17
#define NRCU_PDF_HEMISPHERE (1.0f / (2.0f * NRCU_PT_PI))   // HemiSphere sampler: uniform over the hemisphere

// RayCast: pixel corner, no jitter (RayCastRenderer.cpp:27-35).  p = output pixel index (row 0 = top).
NR_HD Ray raycast_camera_ray(const DScene& s, uint32_t p) {
    int w = (int)s.width, h = (int)s.height;
    int row = (int)(p / (uint32_t)w), j = (int)(p % (uint32_t)w), i = h - 1 - row;
    float sx = (float)j / (float)w, ty = (float)i / (float)h;
    Ray r; r.o = s.cam.position;
    r.d = normalize(s.cam.lower_left + s.cam.horizontal * sx + s.cam.vertical * ty - s.cam.position);
    return r;
}

// Path tracers: U(-1,1)^2 jitter around the pixel corner (AccPathTracer.cpp:23-29) and the
// UniformInCircle lens sample (accept iff x*2 + y*2 <= 1, sic).  The lens draws are skipped when
// the lens radius is 0: the reference multiplies them by zero.
NR_HD Ray pt_camera_ray(const DScene& s, uint64_t seed, uint32_t p, uint32_t sample, float* film_x = nullptr, float* film_y = nullptr) {
    u32x4 rn = rng_block(seed, p, sample, NRCU_STREAM_CAMERA, 0);
    float rx = 2.f * u01(rn.x) - 1.f, ry = 2.f * u01(rn.y) - 1.f;
    int w = (int)s.width, h = (int)s.height;
    int row = (int)(p / (uint32_t)w), j = (int)(p % (uint32_t)w), i = h - 1 - row;
    float x = ((float)j + rx) / (float)w;
    float y = ((float)i + ry) / (float)h;
    if (film_x) { *film_x = x; *film_y = y; }
    vec3 offset = mk3(0.f);
    if (s.cam.lens_radius != 0.f) {
        float lx = 0.f, ly = 0.f;
        for (uint32_t blk = 1;; blk++) {
            rn = rng_block(seed, p, sample, NRCU_STREAM_CAMERA, blk);
            lx = 2.f * u01(rn.x) - 1.f; ly = 2.f * u01(rn.y) - 1.f;
            if (!((lx * 2 + ly * 2) > 1)) break;
            lx = 2.f * u01(rn.z) - 1.f; ly = 2.f * u01(rn.w) - 1.f;
            if (!((lx * 2 + ly * 2) > 1)) break;
        }
        offset = s.cam.u * (lx * s.cam.lens_radius) + s.cam.v * (ly * s.cam.lens_radius);
    }
    Ray r;
    r.o = s.cam.position + offset;
    r.d = normalize(s.cam.lower_left + s.cam.horizontal * x + s.cam.vertical * y - s.cam.position - offset);
    return r;
}

// Surface normal of primitive `id` at the hit point.  Spheres: (hit - c)/r, outward even from the
// inside (intersections.cpp:44-45); everything else: the stored normal (normalised at upload for RayCast).
// The record's w word is prim_meta[id] again (kind | material << 2): one gather per hit instead of two.
NR_HD vec3 hit_normal(const DScene& s, int id, vec3 hit_point, int& material) {
    f4 sh = ldg4(s.prim_shade + id);
    const uint32_t kind = (uint32_t)f2i(sh.w) & 3u;
    material = (int)((uint32_t)f2i(sh.w) >> 2);
    if (kind == KIND_SPHERE) {
        f4 g0 = ldg4(s.prim_geom + 3 * (size_t)id);
        return (hit_point - mk3(g0.x, g0.y, g0.z)) / g0.w;
    }
    return mk3(sh.x, sh.y, sh.z);
}

NR_HD float clamp01(float x) { if (x > 1.f) return 1.f; if (x < 0.f) return 0.f; return x; }   // geometry/vec.hpp:86-91

// ---- RayCast ---------------------------------------------------------------------------------
NR_HD vec3 raycast_shade(const DScene& s, int material, vec3 in, vec3 out, vec3 normal) {
    const DMaterial& m = s.materials[material];
    vec3 diffuse_color = ld3(m.diffuse_color);
    if (m.type == 1) {   // Phong::shade, ray_cast/src/shaders/Phong.cpp:25-31
        vec3 r = out - (2 * dot(out, normal)) * normal;
        vec3 diffuse = diffuse_color * dot(out, normal);
        vec3 specular = ld3(m.specular_color) * fabsf(powf(dot(in, r), m.specular_ex));
        return diffuse + specular;
    }
    return diffuse_color * dot(out, normal);   // Lambertian::shade, ray_cast/src/shaders/Lambertian.cpp:12-14
}

// One pixel of RayCastRenderer::render: returns the gamma-corrected colour.
NR_HD vec3 raycast_pixel(const DScene& s, uint32_t p, uint32_t* ray_count) {
    vec3 c = mk3(0.f);
    if (s.n_point_lights >= 1) {
        Ray r = raycast_camera_ray(s, p);
        float t; int id;
        closest_hit_linear<true>(s, r, t, id);
        (*ray_count)++;
        if (id >= 0) {
            vec3 hp = ray_at(r, t);
            int material;
            vec3 n = hit_normal(s, id, hp, material);
            vec3 out = normalize(s.point_position - hp);
            if (!(dot(out, n) < 0)) {
                float distance = length(s.point_position - hp);
                Ray sr; sr.o = hp; sr.d = out;
                float st; int sid;
                closest_hit_linear<true>(s, sr, st, sid);
                (*ray_count)++;
                vec3 col = raycast_shade(s, material, -r.d, out, n);
                if (sid < 0 || st > distance) c = col * s.point_intensity;
            }
        }
    }
    c = mk3(clamp01(c.x), clamp01(c.y), clamp01(c.z));
    return mk3(sqrtf(c.x), sqrtf(c.y), sqrtf(c.z));
}

// ---- path tracing ----------------------------------------------------------------------------
// closestHitLight, AccPathTracer.cpp:101-112 (+ the index of the light that was hit)
NR_HD float closest_light(const DScene& s, const Ray& r, vec3& radiance, int* which = nullptr) {
    float closest = NRCU_INF;
    radiance = mk3(0.f);
    if (which) *which = -1;
    for (uint32_t i = 0; i < s.n_area_lights; i++) {
        const f4* L = s.area_lights + NRCU_LIGHT_F4 * (size_t)i;
        float t;
        if (x_quad<false>(r, ldg4(L), ldg4(L + 1), ldg4(L + 2), (float)0.000001, closest, closest, t) && closest > t) {
            closest = t;
            f4 rad = ldg4(L + 3);
            radiance = mk3(rad.x, rad.y, rad.z);
            if (which) *which = (int)i;
        }
    }
    return closest;
}

// sin/cos of an angle in [0, 2*pi] evaluated the same way on the device, in the host emulation and in
// the oracle.  The reference calls libm's cosf/sinf (Hemisphere.hpp:28-29); CUDA's, glibc's and MSVC's differ in
// the last ulp, which is enough to flip self-intersection decisions (tMin = 1e-6 with no ray offset) and send a
// path down another branch.  A fixed fp32 sequence - three-constant Cody-Waite reduction by pi/2, the Cephes
// polynomials, every product-sum an explicit fused multiply-add - gives the same bits on any IEEE machine, so whole
// diffuse paths are reproducible across CPU and GPU.  Against double sin / cos over ALL floats of [0, 2 pi]
// (tools/micro/sincos_exhaustive.c): max absolute error 9.3e-8, max relative error 1.09 ulp, 98.8 % correctly rounded.
// (Until the end of round 2 this was a double-precision Taylor evaluation, ~55 FP64 instructions per diffuse vertex, 9 % of
// the shading kernel; this form is ~25 fp32 instructions: +1.5 % on the frame, and no FP64 on the hot path.)
NR_HD void sincos_det(float a, float& s, float& c) {
    const float kf = rintf(a * 0.63661977236758134308f);        // 2/pi
    float r = fmaf(-kf, 1.5703125f, a);                          // pi/2 = 1.5703125 + 4.8375...e-4 + 7.5497...e-8
    r = fmaf(-kf, 4.837512969970703125e-4f, r);
    r = fmaf(-kf, 7.54978995489188216e-8f, r);
    const float z = r * r;
    const float sp = fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f);
    const float sr = fmaf(sp * z, r, r);
    const float cp = fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f);
    const float cr = fmaf(cp, z * z, fmaf(-0.5f, z, 1.0f));
    const int k = ((int)kf) & 3;
    s = (k == 0) ? sr : (k == 1) ? cr : (k == 2) ? -sr : -cr;
    c = (k == 0) ? cr : (k == 1) ? -sr : (k == 2) ? -cr : sr;
}

// Lambertian::shade (Lambertian.cpp:16-34) + HemiSphere::sample3d (Hemisphere.hpp:24-32) + Onb (Onb.hpp:17-27);
// factor = attenuation * dot(N, dir) / pdf as used at AccPathTracer.cpp:142.
// x / pdf for the hemisphere pdf 1 / (2 PI) (Lambertian.cpp:33, AccPathTracer.cpp:142) without the IEEE division sequence:
// with c = pdf and y = RN(1 / c), q = RN(x y), r = x - q c (exact in an FMA), q' = RN(q + r y) IS the correctly rounded
// quotient RN(x / c) for every float x with 1e-30 <= |x| <= 1e30 - checked exhaustively over all 2^32 bit patterns for this
// constant (tools/micro/div_by_pdf_exhaustive.c: 0 mismatches inside that range); outside it, and for zero, the division runs.
// Same bits as `x / pdf` in the oracle, three instructions instead of about ten per colour channel.
NR_HD float div_by_hemisphere_pdf(float x) {
    const float c = 1 / (2 * NRCU_PT_PI);
    const float y = 1.0f / c;
    if (x == 0.f) return x;
    const float ax = fabsf(x);
    if (ax >= 1e-30f && ax <= 1e30f) {
        const float q = x * y;
        const float r = fmaf(-q, c, x);
        return fmaf(r, y, q);
    }
    return x / c;
}
// `attenuation` = albedo / PI (Lambertian.cpp:30), precomputed per material at upload with the same fp32 division
// (DScene::mat_head): three IEEE divisions per vertex less - 5 % of the shading kernel's instructions.
NR_HD Ray shade_lambertian(vec3 attenuation, vec3 hit_point, vec3 normal, float e1, float e2, vec3& factor) {
    const float C_PI = 3.14159265358979323846264338327950288f;
    float r = sqrtf(1 - e1 * e1);
    float sn, cs;
    sincos_det(2 * C_PI * e2, sn, cs);
    float x = cs * r;
    float y = sn * r;
    float z = e1;
    vec3 w = normal;
    // Onb.hpp:20 compares in double: (double)|w.x| > 0.9.  0.9f = 0.89999997... < 0.9 < its fp32 successor, so for a float
    // operand that is the same predicate as |w.x| > 0.9f - without the conversion and the FP64 compare.
#ifndef NRCU_OPT_ONB_FLOAT
#define NRCU_OPT_ONB_FLOAT 1
#endif
#if NRCU_OPT_ONB_FLOAT
    vec3 a = (fabsf(w.x) > 0.9f) ? mk3(0, 1, 0) : mk3(1, 0, 0);
#else
    vec3 a = ((double)fabsf(w.x) > 0.9) ? mk3(0, 1, 0) : mk3(1, 0, 0);
#endif
    vec3 v = normalize(cross(w, a));
    vec3 u = cross(w, v);
    vec3 local = x * u + y * v + z * w;
    Ray out; out.o = hit_point; out.d = normalize(local);
    float n_dot_in = dot(normal, out.d);
    const vec3 an = attenuation * n_dot_in;   // (attenuation * n_dot_in) / pdf, evaluated left to right as in the reference
    factor = mk3(div_by_hemisphere_pdf(an.x), div_by_hemisphere_pdf(an.y), div_by_hemisphere_pdf(an.z));
    return out;
}

// Conductor::shade, Conductor.cpp:6-42
NR_HD Ray shade_conductor(const DMaterial& m, const Ray& ray, vec3 hit_point, vec3 normal, vec3& factor) {
    vec3 V = -ray.d;
    vec3 N = normalize(normal);
    vec3 L = normalize(-V + (2.f * dot(V, N)) * N);
    float cos_l = fabsf(dot(L, N));
    float cos2 = cos_l * cos_l, sin2 = 1 - cos2, sin4 = sin2 * sin2;
    vec3 eta_r = ld3(m.eta_r), eta_i = ld3(m.eta_i), albedo = ld3(m.albedo);
    vec3 temp1 = eta_r * eta_r - eta_i * eta_i - sin2;
    vec3 a2pb2 = temp1 * temp1 + ((4.0f * eta_i) * eta_i) * eta_r * eta_r;
    a2pb2 = mk3(sqrtf(fmaxf(0.0f, a2pb2.x)), sqrtf(fmaxf(0.0f, a2pb2.y)), sqrtf(fmaxf(0.0f, a2pb2.z)));
    vec3 a = 0.5f * (a2pb2 + temp1);
    a = mk3(sqrtf(fmaxf(0.f, a.x)), sqrtf(fmaxf(0.f, a.y)), sqrtf(fmaxf(0.f, a.z)));
    vec3 term1 = a2pb2 + cos2, term2 = (2.f * cos_l) * a;
    vec3 term3 = a2pb2 * cos2 + sin4, term4 = term2 * sin2;
    vec3 rs = (term1 - term2) / (term1 + term2);
    vec3 rp = rs * (term3 - term4) / (term3 + term4);
    vec3 F = 0.5f * (rs + rp);
    factor = F * fabsf(dot(L, N)) * albedo;
    Ray out; out.o = hit_point; out.d = L;
    return out;
}

struct GlassSplit { Ray reflex, refraction; vec3 reflex_rate, refraction_rate; };
// (float)pow(double(x), 5) of the reference (Glass.cpp:37, Microfacet.cpp:34): four double products
// round to the same float except on a ~1e-8 measure of inputs.
NR_HD float pow5(float x) { double d = (double)x; double d2 = d * d; return (float)((d2 * d2) * d); }
// Glass::shade, Glass.cpp:15-57 (including its non-Snell refraction direction and the x_ > 1 branch
// that uses the colour `absorbed` as the reflected direction).
NR_HD GlassSplit shade_glass(const DMaterial& m, const Ray& ray, vec3 hit_point, vec3 normal) {
    vec3 absorbed = ld3(m.absorbed);
    float ior = m.ior;
    vec3 N = normalize(normal), V = normalize(ray.d);
    float ior_inverse = ior;
    if (dot(V, N) > 0.f) { N = -N; ior_inverse = 1 / ior; }
    vec3 reflex = normalize(V + (N * 2.f) * dot(-V, N));
    float n12 = (ior_inverse - 1.f) / (ior_inverse + 1.f);
    n12 = n12 * n12;
    float vdn = fabsf(dot(V, N));
    float Fs = n12 + (1.f - n12) * pow5(1 - vdn);
    vec3 F = mk3(Fs);
    GlassSplit g;
    g.reflex_rate = F * absorbed;
    g.refraction_rate = (mk3(1.f) - F) * absorbed;
    vec3 x = normalize(reflex + V);
    vec3 y = normalize(-N);
    // sqrt(pow(1-|V.N|, 2)) / ior_inverse evaluated in double by the reference; 1-|V.N| >= 0 so the
    // sqrt(pow(.,2)) pair is the identity up to rounding.
    float x_ = (float)(sqrt((double)(1 - vdn) * (double)(1 - vdn)) / (double)ior_inverse);
    float y_ = (float)sqrt(1.0 - (double)x_ * (double)x_);
    vec3 refraction = normalize(x * x_ + y * y_);
    if (x_ > 1.f) { reflex = absorbed; g.refraction_rate = mk3(0.f); refraction = mk3(0.f); }
    g.reflex.o = hit_point; g.reflex.d = reflex;
    g.refraction.o = hit_point; g.refraction.d = refraction;
    return g;
}

// SmithG1, Microfacet.cpp:16-32
NR_HD float smith_g1(vec3 v, vec3 h, vec3 n, float roughness) {
    double cos_v_n = dot(v, n);
    if (cos_v_n * dot(v, h) <= 0.0f) return 0.f;
    if (fabs(cos_v_n - 1.0) < DBL_EPSILON) return 1.0f;
    float c2 = (float)(cos_v_n * cos_v_n), t2 = (1.0f - c2) / c2, a2 = roughness * roughness;
    return 2.0f / (1.0f + sqrtf(1.0f + a2 * t2));
}

// Microfacet::shade with Sample/ToWorld/CoordinateSystem, Microfacet.cpp:71-118, 172-222.
NR_HD bool shade_microfacet(const DScene& s, const DMaterial& m, const Ray& ray, vec3 hit_point, vec3 normal, Ray& out, vec3& factor) {
    const float metalness = 0.2f;
    float roughness = m.roughness, F0 = m.f0;
    vec3 albedo = ld3(m.albedo);
    vec3 N = normalize(normal);
    float alpha_2 = roughness * roughness;
    float tan_theta_2 = alpha_2 * s.mf_u1 / (1.0f - s.mf_u1);
    float cos_theta = (float)(1.0 / (double)sqrtf(1.0f + tan_theta_2));
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    vec3 dir = mk3(sin_theta * s.mf_cos_phi, sin_theta * s.mf_sin_phi, cos_theta);
    vec3 B, C;
    if (fabsf(N.x) > fabsf(N.y)) {
        double len_inv = 1.0f / sqrtf(N.x * N.x + N.z * N.z);
        C = mk3((float)(N.z * len_inv), 0.f, (float)(-N.x * len_inv));
    } else {
        double len_inv = 1.0f / sqrtf(N.y * N.y + N.z * N.z);
        C = mk3(0.f, (float)(N.z * len_inv), (float)(-N.y * len_inv));
    }
    B = cross(C, N);
    vec3 H = normalize(dir.x * B + dir.y * C + dir.z * N);
    double ct = (double)cos_theta, q = (double)(1.0f + tan_theta_2 / alpha_2);
    float D = 1.0f / (float)((double)(NRCU_PT_PI * alpha_2) * (ct * ct * ct) * (q * q));
    H = normalize(H);
    vec3 V = -ray.d;
    float pdf = D * fabsf(1.0f / (4.0f * dot(ray.d, H)));
    vec3 L = normalize(normalize(ray.d - (2.0f * dot(ray.d, H)) * H));
    float cos_theta_i = dot(L, N);
    if (pdf == 0.f || dot(ray.d, normal) >= 0.f || cos_theta_i <= 0.f) return false;
    vec3 specularF0 = (1.f - metalness) * mk3(F0) + metalness * albedo;
    vec3 F = specularF0 + (mk3(1.0f) - specularF0) * pow5(1.0f - fabsf(dot(L, H)));
    float cos_theta_o = fabsf(dot(N, V));
    float G = smith_g1(L, H, N, roughness) * smith_g1(V, H, N, roughness);
    vec3 att = (F * G * D) / fabsf(4.0f * cos_theta_o);
    att = att / pdf;
    att = att * albedo;
    out.o = hit_point; out.d = L;
    factor = att;
    return true;
}

// Extension: latitude-longitude lookup of the ambient environment map on a miss.
// A zero or NaN direction (the reference's glass shader produces both) looks up nothing: black.
NR_HD vec3 env_lookup(const DScene& s, vec3 d) {
    if (!(dot(d, d) > 0.f) || !(dot(d, d) < NRCU_INF)) return mk3(0.f);
    vec3 n = normalize(d);
    float u = 0.5f + atan2f(n.x, n.z) * (0.5f / NRCU_PT_PI);
    float cy = fminf(1.f, fmaxf(-1.f, n.y));
    float v = acosf(cy) * (1.0f / NRCU_PT_PI);
    int x = (int)(u * (float)s.env_w), y = (int)(v * (float)s.env_h);
    x = x < 0 ? 0 : (x > s.env_w - 1 ? s.env_w - 1 : x);
    y = y < 0 ? 0 : (y > s.env_h - 1 ? s.env_h - 1 : y);
    f4 px = ldg4(s.env_rgba + (size_t)y * s.env_w + x);
    return mk3(px.x, px.y, px.z);
}

// ---- environment-map importance sampling (extension, NRCU_FLAG_ENV_IS) ---------------------------------------------------
// The map is piecewise constant (nearest-texel lookup above), so a texel is drawn with probability proportional to
// luminance x sin(theta_row) and the direction uniformly inside the texel: marginal CDF over rows, conditional CDF per row.
NR_HD float env_luminance(f4 px) { return (px.x + px.y) + px.z; }
// table layout: DScene::env_tab.  Row pass: one work item per row (sequential fp32 prefix sums: the same on every machine).
NR_HD void env_table_row(const f4* rgba, int w, int h, int y, float* tab) {
    const float C_PI = 3.14159265358979323846264338327950288f;
    float sn, cs;
    sincos_det(C_PI * ((float)y + 0.5f) / (float)h, sn, cs);
    tab[y] = sn;
    float* cdf = tab + 2 * (size_t)h + (size_t)y * w;
    float sum = 0.f;
    for (int x = 0; x < w; x++) { sum += env_luminance(rgba[(size_t)y * w + x]); cdf[x] = sum; }
    for (int x = 0; x < w; x++) cdf[x] = sum > 0.f ? cdf[x] / sum : (float)(x + 1) / (float)w;
    cdf[w - 1] = 1.f;
    tab[h + y] = sn * sum;   // row weight; turned into the marginal CDF by env_table_marginal
}
NR_HD float env_table_marginal(int h, float* tab) {   // ONE work item, after every row; returns the total weight
    float total = 0.f;
    for (int y = 0; y < h; y++) { total += tab[h + y]; tab[h + y] = total; }
    for (int y = 0; y < h; y++) tab[h + y] = total > 0.f ? tab[h + y] / total : (float)(y + 1) / (float)h;
    tab[2 * h - 1] = 1.f;
    return total;
}
// smallest i in [0, n) with e < cdf[i] (cdf[n-1] = 1 > e), and e remapped to [0,1) inside that step
NR_HD int cdf_pick(const float* cdf, int n, float e, float& frac) {
    int lo = 0, hi = n - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (e < cdf[mid]) hi = mid; else lo = mid + 1; }
    float c0 = lo ? cdf[lo - 1] : 0.f, c1 = cdf[lo];
    frac = c1 > c0 ? (e - c0) / (c1 - c0) : 0.5f;
    if (!(frac >= 0.f)) frac = 0.f;
    if (frac > 0.99999994f) frac = 0.99999994f;
    return lo;
}
// solid-angle pdf of drawing direction n (unit) through texel (x, y): sin_row lum w h / (total 2 pi^2 sin(theta_n))
NR_HD float env_pdf(const DScene& s, int x, int y, vec3 n) {
    float st = sqrtf(fmaxf(0.f, 1.f - n.y * n.y));
    if (!(st > 0.f) || !(s.env_total > 0.f)) return 0.f;
    float lum = env_luminance(ldg4(s.env_rgba + (size_t)y * s.env_w + x));
    return s.env_tab[y] * lum * (float)s.env_w * (float)s.env_h / (s.env_total * (2.f * NRCU_PT_PI * NRCU_PT_PI) * st);
}
// texel that env_lookup reads for direction d (d non-zero, finite)
NR_HD void env_texel(const DScene& s, vec3 n, int& x, int& y) {
    float u = 0.5f + atan2f(n.x, n.z) * (0.5f / NRCU_PT_PI);
    float cy = fminf(1.f, fmaxf(-1.f, n.y));
    float v = acosf(cy) * (1.0f / NRCU_PT_PI);
    x = (int)(u * (float)s.env_w); y = (int)(v * (float)s.env_h);
    x = x < 0 ? 0 : (x > s.env_w - 1 ? s.env_w - 1 : x);
    y = y < 0 ? 0 : (y > s.env_h - 1 ? s.env_h - 1 : y);
}
// One direct sample of the map at a Lambertian vertex: shadow ray + what it adds if it leaves the scene unoccluded
// (f Le / (p_env + p_hemisphere), balance heuristic with the hemisphere sample that follows).
NR_HD bool env_sample(const DScene& s, vec3 albedo, vec3 hit_point, vec3 normal, vec3 thr, float e1, float e2, Ray& shadow, vec3& contrib) {
    const float C_PI = 3.14159265358979323846264338327950288f;
    const int w = s.env_w, h = s.env_h;
    float jy, jx;
    const int y = cdf_pick(s.env_tab + h, h, e1, jy);
    const int x = cdf_pick(s.env_tab + 2 * (size_t)h + (size_t)y * w, w, e2, jx);
    const float u = ((float)x + jx) / (float)w, v = ((float)y + jy) / (float)h;
    float sp, cp, st, ct;
    sincos_det(2.f * C_PI * u, sp, cp);          // phi = 2 pi u - pi: sin(phi) = -sin(2 pi u), cos(phi) = -cos(2 pi u)
    sincos_det(C_PI * v, st, ct);
    vec3 d = mk3(st * -sp, ct, st * -cp);
    float cos_s = dot(normal, d);
    if (!(cos_s > 0.f)) return false;
    float p_env = env_pdf(s, x, y, d);
    if (!(p_env > 0.f)) return false;
    f4 px = ldg4(s.env_rgba + (size_t)y * w + x);
    vec3 f = (albedo / NRCU_PT_PI) * cos_s;
    contrib = thr * f * mk3(px.x, px.y, px.z) * (1.0f / (p_env + NRCU_PDF_HEMISPHERE));
    if (is_zero(contrib)) return false;
    shadow.o = hit_point; shadow.d = d;
    return true;
}
// Balance-heuristic weight of a hemisphere sample that left the scene in direction d (the previous vertex also sampled the map).
NR_HD float mis_env_weight(const DScene& s, vec3 d) {
    if (!(dot(d, d) > 0.f) || !(dot(d, d) < NRCU_INF)) return 1.f;
    vec3 n = normalize(d);
    int x, y; env_texel(s, n, x, y);
    return NRCU_PDF_HEMISPHERE / (NRCU_PDF_HEMISPHERE + env_pdf(s, x, y, n));
}
#define NRCU_NEE_LIGHT_ENV (-2)   // PathStep::nee_light of a shadow ray aimed at the environment map

// Result of processing one path vertex.
enum { PATH_CONTINUE = 0, PATH_TERMINATE = 1, PATH_SPLIT = 2 };
struct PathStep {
    int action;
    vec3 radiance;      // thr * L to add to the pixel when the path ends here (may be zero)
    Ray next; vec3 thr; // continuation
    Ray next2; vec3 thr2; // second branch (PATH_SPLIT, glass branch mode)
    // next-event estimation (extension): a shadow ray towards a sampled light point and what it adds if unoccluded
    bool nee; Ray shadow; vec3 nee_contrib; int nee_light;
    bool next_skips_light;   // the continuation weighs a light it hits with the MIS weight (its vertex also sent a shadow ray)
};

// Next-event estimation at a Lambertian vertex — NOT in the reference (area lights are only hit by
// chance); an opt-in estimator with the SAME expectation.  The reference integrates (albedo/pi)(N.w) L_in(w) over the
// hemisphere about N with pdf p_b = 1/(2 pi); the part of L_in that is light seen first along w can also be sampled
// from the light's area (two-sided quad, xAreaLight accepts both faces) with solid-angle pdf
// p_l(w) = r^2 / (|n_L.w| n_lights), n_L = u x v (length = area).  Both strategies are combined with the balance
// heuristic (multiple importance sampling): the shadow ray carries f Le / (p_l + p_b), and a hemisphere sample that
// finds the light is weighted by p_b / (p_b + p_l) at the next vertex (mis_light_weight).  Without the weights the
// light of path_tracing_cornel.scn, 3 units under the ceiling, makes the 1/r^2 term explode.  V is evaluated by the
// caller with the same closest-hit and closest-light queries the reference applies to any ray (so self-intersection
// "acne" blocks the light exactly where it would have turned a hemisphere sample into a surface hit).
NR_HD bool nee_sample(const DScene& s, vec3 albedo, vec3 hit_point, vec3 normal, vec3 thr, float e1, float e2, Ray& shadow, vec3& contrib, int& light) {
    if (s.n_area_lights == 0) return false;
    float fl = e1 * (float)s.n_area_lights;
    int li = (int)fl; if (li > (int)s.n_area_lights - 1) li = (int)s.n_area_lights - 1;
    float a = fl - (float)li;
    const f4* L = s.area_lights + NRCU_LIGHT_F4 * (size_t)li;
    f4 l0 = ldg4(L), l1 = ldg4(L + 1), rad = ldg4(L + 3), lu = ldg4(L + 4), lv = ldg4(L + 5);
    vec3 nl = mk3(l0.x, l0.y, l0.z), p = mk3(l0.w, l1.x, l1.y);
    vec3 y = p + mk3(lu.x, lu.y, lu.z) * a + mk3(lv.x, lv.y, lv.z) * e2;
    vec3 wv = y - hit_point;
    float r2 = dot(wv, wv);
    if (!(r2 > 0.f)) return false;
    vec3 w = wv * (1.0f / sqrtf(r2));
    float cos_s = dot(normal, w);
    float cl = fabsf(dot(nl, w));
    if (!(cos_s > 0.f) || !(cl > 0.f)) return false;
    float p_l = r2 / (cl * (float)s.n_area_lights);
    vec3 f = (albedo / NRCU_PT_PI) * cos_s;
    contrib = thr * f * mk3(rad.x, rad.y, rad.z) * (1.0f / (p_l + NRCU_PDF_HEMISPHERE));
    if (is_zero(contrib)) return false;
    shadow.o = hit_point; shadow.d = w; light = li;
    return true;
}
// Balance-heuristic weight of a hemisphere sample that reached light `which` at distance tl along `ray`.
NR_HD float mis_light_weight(const DScene& s, const Ray& ray, int which, float tl) {
    f4 l0 = ldg4(s.area_lights + NRCU_LIGHT_F4 * (size_t)which);
    float cl = fabsf(dot(mk3(l0.x, l0.y, l0.z), ray.d));
    if (!(cl > 0.f)) return 1.f;
    float p_l = tl * tl * dot(ray.d, ray.d) / (cl * (float)s.n_area_lights);
    return NRCU_PDF_HEMISPHERE / (NRCU_PDF_HEMISPHERE + p_l);
}
// Visibility of the sampled light along the shadow ray, given the closest object hit (t_obj, id_obj): the light
// counts iff it is the nearest light along the ray and no object is closer (AccPathTracer.cpp:128-130, 174-176).
NR_HD bool nee_visible(const DScene& s, const Ray& shadow, int light, float t_obj, int id_obj) {
    vec3 rad; int which;
    float tl = closest_light(s, shadow, rad, &which);
    if (light == NRCU_NEE_LIGHT_ENV) return id_obj < 0 && which < 0;   // the map is seen iff the ray leaves the scene: no object, no light quad
    if (which != light) return false;
    return !(id_obj >= 0 && t_obj < tl);
}

// The surface-hit branch of trace() (AccPathTracer.cpp:131-172): the ray's closest object hit (t, id) is in front of
// every light.  `ps` comes in initialised (path_vertex below); shared by the per-vertex form and by the pooled shading
// kernel, which runs it on dense warps of surface hits only.
// hit point, normal, material of the hit - everything the shading needs that does not depend on random numbers.
struct HitSetup { vec3 hp, n; int material; f4 mh; uint32_t type; };
NR_HD HitSetup hit_setup(const DScene& s, const Ray& ray, float t, int id) {
    HitSetup hs;
    hs.hp = ray_at(ray, t);
    hs.n = hit_normal(s, id, hs.hp, hs.material);
    hs.mh = ldg4(s.mat_head + hs.material);
    hs.type = s.mode == MODE_ACC ? (uint32_t)f2i(hs.mh.w) : 0u;
    return hs;
}
// Does a vertex on a material of this type always continue the path (below the depth limit)?  Lambertian and conductor
// vertices do; glass and microfacet vertices can end it (Glass.cpp:15-57, Microfacet.cpp:71-118).
NR_HD bool type_always_continues(uint32_t type) { return type != 2u && type != 3u; }
template <bool NEE>
NR_HD void path_vertex_shade(PathStep& ps, const DScene& s, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t d, uint32_t branch,
                             const Ray& ray, vec3 thr, const HitSetup& hs, int glass_branch_mode) {
    {
        const vec3 hp = hs.hp, n = hs.n;
        const DMaterial& m = s.materials[hs.material];
        const f4 mh = hs.mh;
        const uint32_t type = hs.type;
        if (type == 2u) {
            GlassSplit g = shade_glass(m, ray, hp, n);
            bool refl_zero = is_zero(g.reflex_rate);
            if (glass_branch_mode) {
                ps.next = g.reflex; ps.thr = thr * g.reflex_rate; ps.action = PATH_CONTINUE;
                if (!refl_zero) { ps.next2 = g.refraction; ps.thr2 = thr * g.refraction_rate; ps.action = PATH_SPLIT; }
            } else {
                float q = g.reflex_rate.x + g.reflex_rate.y + g.reflex_rate.z;
                float q2 = g.refraction_rate.x + g.refraction_rate.y + g.refraction_rate.z;
                if (refl_zero || !(q + q2 > 0.f)) return;
                float pr = q / (q + q2);
                u32x4 rn = rng_block(seed, pixel, sample, d, branch);
                if (u01(rn.z) < pr) { ps.thr = thr * (g.reflex_rate / pr); ps.next = g.reflex; }
                else { ps.thr = thr * (g.refraction_rate / (1.f - pr)); ps.next = g.refraction; }
                ps.action = PATH_CONTINUE;
            }
        } else if (type == 1u) {
            vec3 f; ps.next = shade_conductor(m, ray, hp, n, f); ps.thr = thr * f; ps.action = PATH_CONTINUE;
        } else if (type == 3u) {
            vec3 f; Ray nr;
            if (shade_microfacet(s, m, ray, hp, n, nr, f)) { ps.next = nr; ps.thr = thr * f; ps.action = PATH_CONTINUE; }
        } else {
            // type 0 (any other type falls off the end of the reference's trace(): treated as Lambertian)
            u32x4 rn = rng_block(seed, pixel, sample, d, branch);
            vec3 f; ps.next = shade_lambertian(mk3(mh.x, mh.y, mh.z), hp, n, u01(rn.x), u01(rn.y), f);
            ps.thr = thr * f; ps.action = PATH_CONTINUE;
            // NEE only where the continuation will really be traced: at the depth limit the reference returns the
            // ambient colour without looking for the light (AccPathTracer.cpp:122)
            if (NEE && d + 1 < s.depth) {
                if (s.nee == 2) { ps.nee = env_sample(s, ld3(m.diffuse_color), hp, n, thr, u01(rn.z), u01(rn.w), ps.shadow, ps.nee_contrib); ps.nee_light = NRCU_NEE_LIGHT_ENV; }
                else ps.nee = nee_sample(s, ld3(m.diffuse_color), hp, n, thr, u01(rn.z), u01(rn.w), ps.shadow, ps.nee_contrib, ps.nee_light);
                ps.next_skips_light = true;   // "the vertex before this ray also sampled its light source (area lights / the map) directly"
            }
        }
        // the continuation would return ambient at the depth limit (AccPathTracer.cpp:122)
        if (ps.action != PATH_TERMINATE && d + 1 == s.depth) {
            vec3 add = ps.thr * s.ambient;
            if (ps.action == PATH_SPLIT) add = add + ps.thr2 * s.ambient;
            ps.radiance = add; ps.action = PATH_TERMINATE;
        }
    }
}
template <bool NEE>
NR_HD void path_vertex_hit(PathStep& ps, const DScene& s, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t d, uint32_t branch,
                           const Ray& ray, vec3 thr, float t, int id, int glass_branch_mode) {
    const HitSetup hs = hit_setup(s, ray, t, id);
    path_vertex_shade<NEE>(ps, s, seed, pixel, sample, d, branch, ray, thr, hs, glass_branch_mode);
}
NR_HD void path_step_init(PathStep& ps, const Ray& ray, vec3 thr) {
    ps.action = PATH_TERMINATE; ps.radiance = mk3(0.f);
    ps.next = ray; ps.thr = thr; ps.next2 = ray; ps.thr2 = mk3(0.f);
    ps.nee = false; ps.shadow = ray; ps.nee_contrib = mk3(0.f); ps.nee_light = -1; ps.next_skips_light = false;
}

// One iteration of trace() for a ray at bounce `d` (d < depth) whose closest object hit is
// (t, id) — AccPathTracer.cpp:121-181.  `branch` = glass branch bits (RNG block).
template <bool NEE = false>
NR_HD PathStep path_vertex(const DScene& s, uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t d, uint32_t branch,
                           const Ray& ray, vec3 thr, float t, int id, int glass_branch_mode, bool skip_light = false) {
    PathStep ps;
    path_step_init(ps, ray, thr);
    vec3 radiance;
    int which_light = -1;
    float tl = closest_light(s, ray, radiance, NEE ? &which_light : nullptr);
    if (id >= 0 && t < tl) {
        path_vertex_hit<NEE>(ps, s, seed, pixel, sample, d, branch, ray, thr, t, id, glass_branch_mode);
    } else if (tl != NRCU_INF) {
        ps.radiance = thr * radiance;
        if (NEE && skip_light && s.nee == 1) ps.radiance = ps.radiance * mis_light_weight(s, ray, which_light, tl);
    } else if (s.env_rgba && s.mode == MODE_ACC) {
        ps.radiance = thr * env_lookup(s, ray.d);
        if (NEE && skip_light && s.nee == 2) ps.radiance = ps.radiance * mis_env_weight(s, ray.d);
    }
    return ps;
}

}  // namespace nrcu
