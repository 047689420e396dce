// nrcu_mlt.cuh — Metropolis light transport in primary sample space.
//
// Counterpart of the reference's MetropolisLightTransport component
//   code/components/metropolis_light_transport/src/Metropolis.cpp:25-135     renderTask / render (the Metropolis loop, b, tone map)
//   code/components/metropolis_light_transport/include/Metropolis.hpp:89-157 rnd / perturb / large_step / mutate
//   code/components/metropolis_light_transport/include/PathContribution.hpp:12-21  LargeStepProb 0.3, N_Init, 2 numbers per event
// which is a port of Kelemen et al.'s PSSMLT: a Markov chain over the vector of random numbers that drives the path
// sampler, large steps with probability 0.3, small steps with the exponential perturbation (s1 = 1/1024, s2 = 1/64; the
// two film coordinates with s1 = 2/(w+h), s2 = 0.1), expected-value accumulation
//      proposal += c' (a + large) / (I'/b + pLarge)        current += c (1 - a) / (I/b + pLarge)
// and b = the mean scalar contribution of N_Init independent samples.
//
// What is NOT taken over: the reference drives a bidirectional sampler with hard-coded material colours and emission and
// shares one random-number array between its 8 unsynchronised threads, so its output is not reproducible.  Here the chain
// drives THIS backend's path sampler - the same closest-hit and material code as the path tracers (closest_hit_bvh,
// shade_lambertian / conductor / glass / microfacet) with the scene's own materials and lights - so the Metropolis frame has
// the expectation of the path-traced frame, which is what tests/test_gpu_parity.py checks.  One chain per thread; every
// chain owns its numbers (Philox keyed by chain and mutation), so a frame is reproducible up to the order of the float
// atomics that splat it.
#pragma once
#include "nrcu_shade.cuh"

namespace nrcu {

#define NRCU_MLT_MAX_DEPTH 32                       // numbers per chain: 2 film coordinates + 2 per bounce
#define NRCU_MLT_STATES(depth) (2u + 2u * (depth))
#define NRCU_STREAM_MLT 0xFFFFFFFEu                 // Philox stream of the chains (the path tracers use the bounce index / 0xFFFFFFFF)

// Metropolis.hpp:100-121
NR_HD float mlt_perturb(float value, float s1, float s2, float r) {
    float result;
    if (r < 0.5f) {
        r = r * 2.0f;
        result = value + s2 * expf(-logf(s2 / s1) * r);
        if (result > 1.0f) result -= 1.0f;
    } else {
        r = (r - 0.5f) * 2.0f;
        result = value - s2 * expf(-logf(s2 / s1) * r);
        if (result < 0.0f) result += 1.0f;
    }
    if (!(result >= 0.f)) result = 0.f;
    if (result > 0.99999994f) result = 0.99999994f;
    return result;
}

// The path sampler as a function of the chain's numbers u[0 .. 2 + 2 depth): film position (in pixel units, over the film
// extended by one pixel on every side - the path tracers' U(-1,1) jitter gives every pixel a two-pixel-wide footprint,
// AccPathTracer.cpp:23-29) and the radiance the path carries.  trace() of AccPathTracer.cpp:121-181 with the random
// numbers supplied by the caller; glass picks a branch with the vertex's first number.
template <bool GATE>
NR_HD vec3 mlt_eval(const DScene& s, const float* u, float& film_x, float& film_y, uint32_t& rays) {
    const float w = (float)s.width, h = (float)s.height;
    film_x = u[0] * (w + 2.f) - 1.f;
    film_y = u[1] * (h + 2.f) - 1.f;
    if (s.depth == 0) return s.ambient;
    Ray ray;
    ray.o = s.cam.position;
    ray.d = normalize(s.cam.lower_left + s.cam.horizontal * (film_x / w) + s.cam.vertical * (film_y / h) - s.cam.position);
    vec3 thr = mk3(1.f), L = mk3(0.f);
    for (uint32_t d = 0; d < s.depth; d++) {
        float t; int id;
        LocalStack st;
        closest_hit_bvh<GATE>(s, ray, st, t, id);
        rays++;
        vec3 radiance;
        float tl = closest_light(s, ray, radiance);
        if (id >= 0 && t < tl) {
            vec3 hp = ray_at(ray, t);
            int material;
            vec3 n = hit_normal(s, id, hp, material);
            const DMaterial& m = s.materials[material];
            const uint32_t type = s.mode == MODE_ACC ? m.type : 0u;
            const float e1 = u[2 + 2 * d], e2 = u[3 + 2 * d];
            vec3 f;
            if (type == 2u) {
                GlassSplit g = shade_glass(m, ray, hp, n);
                float q = g.reflex_rate.x + g.reflex_rate.y + g.reflex_rate.z, q2 = g.refraction_rate.x + g.refraction_rate.y + g.refraction_rate.z;
                if (is_zero(g.reflex_rate) || !(q + q2 > 0.f)) return L;
                float pr = q / (q + q2);
                if (e1 < pr) { thr = thr * (g.reflex_rate / pr); ray = g.reflex; }
                else { thr = thr * (g.refraction_rate / (1.f - pr)); ray = g.refraction; }
            } else if (type == 1u) {
                ray = shade_conductor(m, ray, hp, n, f); thr = thr * f;
            } else if (type == 3u) {
                Ray nr;
                if (!shade_microfacet(s, m, ray, hp, n, nr, f)) return L;
                ray = nr; thr = thr * f;
            } else {
                { const f4 mh = ldg4(s.mat_head + material); ray = shade_lambertian(mk3(mh.x, mh.y, mh.z), hp, n, e1, e2, f); } thr = thr * f;
            }
            if (d + 1 == s.depth) return L + thr * s.ambient;   // trace() at the depth limit (AccPathTracer.cpp:122)
        } else if (tl != NRCU_INF) {
            return L + thr * radiance;
        } else {
            if (s.env_rgba && s.mode == MODE_ACC) L = L + thr * env_lookup(s, ray.d);
            return L;
        }
    }
    return L;
}

NR_HD float mlt_scalar(vec3 c) {   // the reference's scalar contribution function: Max(c) (Metropolis.hpp:603)
    float m = fmaxf(c.x, fmaxf(c.y, c.z));
    return (m > 0.f && m < NRCU_INF) ? m : 0.f;
}

}  // namespace nrcu
