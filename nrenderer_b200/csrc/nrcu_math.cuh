// nrcu_math.cuh — fp32 vector arithmetic in the reference's (glm 0.9.9.9) operation order.
//
// Everything here is __host__ __device__ so that tests/host_emu can run the very same device
// functions on the CPU against the oracle.  The translation units are compiled with
// --fmad=false (nvcc) / -ffp-contract=off (g++): a*b+c is never contracted, which is what makes the
// primitive tests bit-identical to the reference build (SURVEY.md §7 "RayCast 1e-4").  Where a
// fused multiply-add is wanted (conservative BVH slab tests) it is written explicitly as fmaf().
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NR_HD __host__ __device__ __forceinline__
#define NR_D __device__ __forceinline__
#else
#define NR_HD inline
#define NR_D inline
#ifndef NRCU_HOST_EMU
#define NRCU_HOST_EMU 1
#endif
#endif

namespace nrcu {

struct vec3 { float x, y, z; };

NR_HD vec3 mk3(float x, float y, float z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
NR_HD vec3 mk3(float s) { return mk3(s, s, s); }
NR_HD vec3 operator+(vec3 a, vec3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
NR_HD vec3 operator-(vec3 a, vec3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
NR_HD vec3 operator*(vec3 a, vec3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
NR_HD vec3 operator/(vec3 a, vec3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
NR_HD vec3 operator*(vec3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
NR_HD vec3 operator*(float s, vec3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
NR_HD vec3 operator/(vec3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
NR_HD vec3 operator+(vec3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
NR_HD vec3 operator-(vec3 a, float s) { return mk3(a.x - s, a.y - s, a.z - s); }
NR_HD vec3 operator-(vec3 a) { return mk3(-a.x, -a.y, -a.z); }
// glm compute_dot<vec3>: tmp = a*b; (tmp.x + tmp.y) + tmp.z   (glm/detail/func_geometric.inl)
NR_HD float dot(vec3 a, vec3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
// glm compute_cross
NR_HD vec3 cross(vec3 x, vec3 y) { return mk3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
// glm compute_normalize: v * inversesqrt(dot(v,v)); inversesqrt(x) = 1/sqrt(x)  (func_exponential.inl:136-139)
NR_HD vec3 normalize(vec3 a) { float s = 1.0f / sqrtf(dot(a, a)); return a * s; }
NR_HD float length(vec3 a) { return sqrtf(dot(a, a)); }
NR_HD float comp(vec3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
NR_HD vec3 vmin(vec3 a, vec3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
NR_HD vec3 vmax(vec3 a, vec3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
NR_HD vec3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
NR_HD void st3(float* p, vec3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
NR_HD bool is_zero(vec3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }

struct Ray { vec3 o, d; };
NR_HD vec3 ray_at(const Ray& r, float t) { return r.o + r.d * t; }   // Ray.hpp:30-33: origin + t*direction

#define NRCU_INF (__builtin_huge_valf())

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11), keyed as DESIGN.md describes:
//   key = (seed lo, seed hi), counter = (pixel, sample, stream, block)
//   stream = bounce index for path vertices (block = glass branch bits), 0xFFFFFFFF for the camera.
// ---------------------------------------------------------------------------------------------
NR_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
struct u32x4 { uint32_t x, y, z, w; };
NR_HD u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    u32x4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3; return o;
}
#define NRCU_STREAM_CAMERA 0xFFFFFFFFu
NR_HD u32x4 rng_block(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t stream, uint32_t block) {
    return philox4x32_10(pixel, sample, stream, block, (uint32_t)seed, (uint32_t)(seed >> 32));
}
// 24-bit uniform in [0,1): the value range of libstdc++'s generate_canonical<float,24>
NR_HD float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

}  // namespace nrcu
