// rng.cuh - counter-based sampler of the path: Squares (Widynski 2020) keyed Owen-scrambled Sobol pairs
// (Burley 2020, Ahmed 2024) with a per-pixel blue-noise toroidal shift. Integer-exact restatement of the
// reference's cuda/random.cuh:144-368; targets follow enum RandomTarget (random.cuh:24-66).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace lbrng {

// enum RandomTarget: every allocation consumes size * sets + 1 enumerators (the END marker takes one).
enum Target : uint32_t {
  T_LENS                    = 33,
  T_LENS_BLADE              = 35,
  T_BSDF_REFLECTION         = 39,  // + set
  T_BSDF_DIFFUSE            = 43,
  T_BSDF_REFRACTION         = 47,
  T_BSDF_RESAMPLING         = 51,
  T_BSDF_OPACITY            = 55,
  T_RUSSIAN_ROULETTE        = 61,
  T_CAMERA_JITTER           = 63,
  T_CAMERA_TIME             = 65,
  T_SKY_STEP_OFFSET         = 77,
  T_SKY_INSCATTERING_STEP   = 79,  // RANDOM_ALLOCATE leaves one unused value after every allocation
  T_LIGHT_SUN_BSDF          = 346,  // + set (the surface uses set 0, material.cuh:61)
  T_LIGHT_SUN_BSDF_METHOD   = 349,
  T_LIGHT_SUN_RAY           = 352,
  T_LIGHT_SUN_RESAMPLING    = 355,
  T_LIGHT_GEO_RAY           = 367,  // + lane (+ 8 * set)
  T_LIGHT_GEO_RESAMPLING    = 384,
  T_LIGHT_GEO_TREE_PREPASS  = 387,  // + lane
  T_LIGHT_GEO_TREE_POSTPASS = 404,  // + lane
  T_LIGHT_BSDF_CHOICE       = 569,
  T_LIGHT_BSDF_DIRECTION    = 571,
  T_LIGHT_BSDF_TRACE        = 573,
  T_LIGHT_BSDF_RR           = 575,
  T_COUNT                   = 577
};

__device__ __forceinline__ uint32_t swap16(uint32_t a) { return __byte_perm(a, a, 0x1032); }

__device__ __forceinline__ uint32_t squares32(uint32_t key, uint32_t counter) {
  uint32_t x = counter * key;
  const uint32_t y = x;
  uint32_t z = y + key;
  x = x * x + y;
  x = swap16(x);
  x = x * x + z;
  x = swap16(x);
  x = x * x + y;
  x = swap16(x);
  x = x * x + z;
  z = x;
  x = swap16(x);
  return z ^ (x * x + y);
}

__device__ __forceinline__ uint32_t lk_perm(uint32_t x, uint32_t seed) {
  x += seed;
  x ^= x * 0x6c50b47cu;
  x ^= x * 0xb82f1e52u;
  x ^= x * 0xc7afe638u;
  x ^= x * 0x8d22f6e6u;
  return x;
}

__device__ __forceinline__ uint32_t owen(uint32_t x, uint32_t seed) { return __brev(lk_perm(__brev(x), seed)); }
__device__ __forceinline__ uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }

__device__ __forceinline__ uint32_t sobol_P(uint32_t v) {
  v ^= v << 16;
  v ^= (v & 0x00FF00FFu) << 8;
  v ^= (v & 0x0F0F0F0Fu) << 4;
  v ^= (v & 0x33333333u) << 2;
  v ^= (v & 0x55555555u) << 1;
  return v;
}

__device__ __forceinline__ uint2 sobol(uint32_t offset, uint32_t dimension) {
  const uint32_t seed = squares32(0xfcbd6e15u, dimension);
  const uint32_t J    = lk_perm(__brev(offset), seed);
  return make_uint2(owen(J, hash_combine(seed, 0)), owen(sobol_P(J), hash_combine(seed, 1)));
}

__device__ __forceinline__ float u32_to_float(uint32_t v) { return __uint_as_float(0x3F800000u | (v >> 9)) - 1.0f; }

// 2 x 32 random bits for (target, pixel, sample, depth); the low bit of each word is always 0 in the
// reference as well (random.cuh:317).
__device__ __forceinline__ uint2 random_2d_bits(const uint32_t* __restrict__ bluenoise, uint32_t target, uint32_t px, uint32_t py,
                                                uint32_t sample_id, uint32_t depth) {
  const uint32_t dim = target + depth * T_COUNT;
  uint2 q            = sobol(sample_id, dim);
  const uint32_t ox  = ((1u + dim) * 3242174889u) >> 24;
  const uint32_t oy  = ((1u + dim) * 2447445413u) >> 24;
  const uint32_t n   = __ldg(bluenoise + ((px + ox) & 0xFFu) + ((py + oy) & 0xFFu) * 256u);
  q.x += n & 0xFFFF0000u;
  q.y += n << 16;
  return q;
}

struct Sampler {
  const uint32_t* bluenoise;
  uint32_t px, py, sample_id, depth;

  __device__ __forceinline__ float2 get2(uint32_t target) const {
    const uint2 q = random_2d_bits(bluenoise, target, px, py, sample_id, depth);
    return make_float2(u32_to_float(q.x), u32_to_float(q.y));
  }
  __device__ __forceinline__ float get1(uint32_t target) const {
    return u32_to_float(random_2d_bits(bluenoise, target, px, py, sample_id, depth).x);
  }
};

// Table-driven variant for the shading kernel. For a given launch, sample_id and depth are uniform, so the whole
// Owen-scrambled Sobol pair of a dimension is a launch constant: k_rng_table evaluates sobol(sample_id, dim) once per
// dimension and only the per-pixel blue-noise shift remains per thread. Bit-identical to Sampler.
// table[dim] = {q.x, q.y, blue-noise x offset, blue-noise y offset}
__device__ __forceinline__ uint4 table_entry(uint32_t sample_id, uint32_t dim) {
  const uint2 q = sobol(sample_id, dim);
  return make_uint4(q.x, q.y, ((1u + dim) * 3242174889u) >> 24, ((1u + dim) * 2447445413u) >> 24);
}

struct TabSampler {
  const uint32_t* bluenoise;
  const uint4* table;  // already offset by depth * T_COUNT
  uint32_t px, py;

  __device__ __forceinline__ uint2 bits(uint32_t target) const {
    const uint4 e    = __ldg(table + target);
    const uint32_t n = __ldg(bluenoise + ((px + e.z) & 0xFFu) + ((py + e.w) & 0xFFu) * 256u);
    return make_uint2(e.x + (n & 0xFFFF0000u), e.y + (n << 16));
  }
  __device__ __forceinline__ float2 get2(uint32_t target) const {
    const uint2 q = bits(target);
    return make_float2(u32_to_float(q.x), u32_to_float(q.y));
  }
  __device__ __forceinline__ float get1(uint32_t target) const { return u32_to_float(bits(target).x); }
};

__device__ __forceinline__ float saturate_random(float r) { return fminf(fmaxf(r, 0.0f), __uint_as_float(0x3F7FFFFFu)); }

}  // namespace lbrng
