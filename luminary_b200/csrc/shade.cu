// shade.cu - surface / miss shading, next-event estimation, BSDF LUT generation and sample accumulation.
//
// Replaces, on the reference's per-bounce path (device/device_renderer.c:53-134):
//   geometry_process_tasks          cuda/geometry.cuh:11-180   (context: geometry_utils.cuh:54-221)
//   sky_process_tasks               cuda/sky.cuh:609-633       (constant colour branch of sky_color_main)
//   NEE task creation               cuda/direct_lighting.cuh:318-443, light.cuh:49-159, light_tree.cuh:191-320,
//                                   light_triangle.cuh, light_bsdf.cuh:24-146, mis.cuh:19-57, ris.cuh:22-157
//   BSDF-sampled light enumeration  direct_lighting.cuh:601-669 + optix_anyhit.cuh:145-205 (emitter BVH8 instead of the light GAS)
//   bsdf_generate_*_lut             cuda/bsdf_lut.cuh:20-209
//   accumulation_collect_results    cuda/accumulation.cuh:36-84, accumulation_generate_result :86-190 (beauty mode)
// One thread shades one path; paths arrive sorted by material so that a warp mostly runs one BSDF
// configuration. The shadow rays it produces are traced afterwards by k_trace_shadow (trace.cu).
// Compiled with --use_fast_math like the reference (src/luminary/CMakeLists.txt:48).
#include <float.h>

#include "rng.cuh"
#include "shade_api.cuh"
#include "texture.cuh"
#include "traverse.cuh"

#define PI_F 3.141592653589f
#define EPS_F FLT_EPSILON
#define DELTA_PATH_CUTOFF 0.05f  // GEOMETRY_DELTA_PATH_CUTOFF, cuda/utils.cuh:45
#define ROUGHNESS_CLAMP 2e-2f    // BSDF_ROUGHNESS_CLAMP, cuda/utils.cuh:46
#define RR_CLAMP (1.0f / 8.0f)   // RUSSIAN_ROULETTE_CLAMP, cuda/directives.cuh:9
#define NUM_TREE_LANES 8         // LIGHT_TREE_NUM_OUTPUTS
#define LB_MAX_ROOT_CHILDREN 128  // LIGHT_TREE_ROOT_MAX_CHILD_COUNT, device_utils.h:45

// MaterialFlag, device_utils.h:252-259
#define MF_TRANSLUCENT 1u
#define MF_INSIDE 2u
#define MF_METALLIC 4u
#define MF_COLORED 8u
// DeviceMaterialFlags, device_structs.h:218-230
#define DMF_TRANSLUCENT 0x01u
#define DMF_EMISSION 0x02u
#define DMF_METALLIC 0x08u
#define DMF_COLORED 0x10u
#define DMF_SMOOTHNESS 0x20u
#define DMF_NORMAL_COMPRESSED 0x40u
#define DMF_BIDIRECTIONAL 0x80u

enum Hint { H_GENERAL = 0, H_MICROFACET = 1, H_DIFFUSE = 2, H_REFRACTION = 3 };

// Material class of a shading kernel instantiation (wavefront.cuh: LB_CLASS_*). What a class rules out is a compile-time constant, so
// the refraction lobe, the Fresnel-transmission terms, the medium stack and the opacity pass-through are not even compiled into
// the kernels that shade opaque dielectrics / metals; the arithmetic that remains is unchanged.
template <int kClass>
struct ShadeClass {
  static constexpr bool may_be_translucent = kClass == LB_CLASS_GENERIC;
  static constexpr bool may_be_transparent = kClass == LB_CLASS_GENERIC;  // opacity < 1
  static constexpr bool may_be_metallic    = kClass != LB_CLASS_DIELECTRIC;
  static constexpr bool may_be_dielectric  = kClass != LB_CLASS_METAL;
  __device__ static __forceinline__ bool translucent(uint32_t flags) { return may_be_translucent && (flags & 1u /* MF_TRANSLUCENT */) != 0; }
  __device__ static __forceinline__ bool metallic(uint32_t flags) {
    return may_be_metallic && (!may_be_dielectric || (flags & 4u /* MF_METALLIC */) != 0);
  }
};

// ---------------------------------------------------------------------------------------------
// small math
// ---------------------------------------------------------------------------------------------
struct C3 {
  float r, g, b;
};
__device__ __forceinline__ C3 c3(float r, float g, float b) {
  C3 c;
  c.r = r, c.g = g, c.b = b;
  return c;
}
__device__ __forceinline__ C3 operator+(C3 a, C3 b) { return c3(a.r + b.r, a.g + b.g, a.b + b.b); }
__device__ __forceinline__ C3 operator*(C3 a, C3 b) { return c3(a.r * b.r, a.g * b.g, a.b * b.b); }
__device__ __forceinline__ C3 operator*(C3 a, float s) { return c3(a.r * s, a.g * s, a.b * s); }
// color_any, math.cuh:944-946: any component > 0 - NaN fails the test, which is how the reference drops a non-finite NEE term
__device__ __forceinline__ bool c_any(C3 a) { return a.r > 0.0f || a.g > 0.0f || a.b > 0.0f; }
__device__ __forceinline__ float c_max(C3 a) { return fmaxf(a.r, fmaxf(a.g, a.b)); }
__device__ __forceinline__ float c_lum(C3 v) { return 0.212655f * v.r + 0.715158f * v.g + 0.072187f * v.b; }

__device__ __forceinline__ float len3(V3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ V3 norm3(V3 a) { return a * rsqrtf(dot3(a, a)); }
__device__ __forceinline__ V3 neg3(V3 a) { return v3(-a.x, -a.y, -a.z); }

__device__ __forceinline__ V3 reflect3(V3 V, V3 n) { return norm3(n * (2.0f * dot3(V, n)) - V); }  // math.cuh:192-197

__device__ __forceinline__ V3 refract3(V3 V, V3 n, float ratio, bool& total_reflection) {  // math.cuh:799-819
  if (ratio < EPS_F) {
    total_reflection = false;
    return neg3(V);
  }
  const float d    = fabsf(dot3(n, V));
  const float b    = 1.0f - ratio * ratio * (1.0f - d * d);
  total_reflection = b < 0.0f;
  if (total_reflection)
    return reflect3(V, n);
  return norm3(n * (ratio * d - sqrtf(b)) - V * ratio);
}

struct Q4 {
  float x, y, z, w;
};
__device__ __forceinline__ Q4 rotation_to_z(V3 v) {  // quaternion_rotation_to_z_canonical, math.cuh:385-409
  Q4 r;
  if (v.z < -1.0f + EPS_F) {
    r.x = 1.0f, r.y = 0.0f, r.z = 0.0f, r.w = 0.0f;
    return r;
  }
  r.x = v.y, r.y = -v.x, r.z = 0.0f, r.w = 1.0f + v.z;
  const float n = rsqrtf(r.x * r.x + r.y * r.y + r.w * r.w);
  r.x *= n, r.y *= n, r.w *= n;
  return r;
}
__device__ __forceinline__ V3 q_apply(Q4 q, V3 v) { return quat_apply(q.x, q.y, q.z, q.w, v); }
__device__ __forceinline__ V3 q_apply_inv(Q4 q, V3 v) { return quat_apply(-q.x, -q.y, -q.z, q.w, v); }

__device__ __forceinline__ V3 unpack_normal(uint32_t data) {  // math.cuh:1716-1730
  float x = (data & 0xFFFFu) * (1.0f / 0xFFFF);
  float y = (data >> 16) * (1.0f / 0xFFFF);
  x       = (x * 2.0f) - 1.0f;
  y       = (y * 2.0f) - 1.0f;
  V3 n    = v3(x, y, 1.0f - fabsf(x) - fabsf(y));
  const float t = __saturatef(-n.z);
  n.x += (n.x >= 0.0f) ? -t : t;
  n.y += (n.y >= 0.0f) ? -t : t;
  return norm3(n);
}

__device__ __forceinline__ uint32_t pack_normal(V3 n) {  // math.cuh:1732-1758
  float x = n.x, y = n.y, z = n.z;
  const float rn = 1.0f / (fabsf(x) + fabsf(y) + fabsf(z));
  x *= rn, y *= rn, z *= rn;
  const float t = fmaxf(fminf(-z, 1.0f), 0.0f);
  x += (x >= 0.0f) ? t : -t;
  y += (y >= 0.0f) ? t : -t;
  x = fmaxf(fminf(x, 1.0f), -1.0f);
  y = fmaxf(fminf(y, 1.0f), -1.0f);
  x = (x + 1.0f) * 0.5f;
  y = (y + 1.0f) * 0.5f;
  return (((uint32_t) (y * 0xFFFF + 0.5f)) << 16) | ((uint32_t) (x * 0xFFFF + 0.5f));
}

__device__ __forceinline__ uint2 record_pack(C3 c) {  // math.cuh:1609-1619
  const uint32_t r = __float_as_uint(c.r) >> 11, g = __float_as_uint(c.g) >> 11, b = __float_as_uint(c.b) >> 11;
  return make_uint2(r | (g << 21), (g >> 11) | (b << 10));
}
__device__ __forceinline__ C3 record_unpack(uint2 p) {  // math.cuh:1595-1607
  const uint32_t r = p.x & 0x1FFFFFu;
  const uint32_t g = (p.x >> 21) | ((p.y & 0x3FFu) << 11);
  const uint32_t b = p.y >> 10;
  return c3(__uint_as_float(r << 11), __uint_as_float(g << 11), __uint_as_float(b << 11));
}

__device__ __forceinline__ uint2 ray_pack(V3 ray) {  // math.cuh:1637-1664
  float x = ray.x, y = ray.y, z = ray.z;
  const float rn = 1.0f / (fabsf(x) + fabsf(y) + fabsf(z));
  x *= rn, y *= rn, z *= rn;
  const float t = __saturatef(-z);
  x += (x >= 0.0f) ? t : -t;
  y += (y >= 0.0f) ? t : -t;
  x = fminf(1.0f, fmaxf(-1.0f, x));
  y = fminf(1.0f, fmaxf(-1.0f, y));
  x = (x + 1.0f) * 0.5f;
  y = (y + 1.0f) * 0.5f;
  return make_uint2((uint32_t) (x * 4294967296.0f + 0.5f), (uint32_t) (y * 4294967296.0f + 0.5f));
}
__device__ __forceinline__ V3 ray_unpack(uint2 p) {  // math.cuh:1621-1635
  float x = p.x * (1.0f / 4294967296.0f);
  float y = p.y * (1.0f / 4294967296.0f);
  x       = (x * 2.0f) - 1.0f;
  y       = (y * 2.0f) - 1.0f;
  V3 r    = v3(x, y, 1.0f - fabsf(x) - fabsf(y));
  const float t = __saturatef(-r.z);
  r.x += (r.x >= 0.0f) ? -t : t;
  r.y += (r.y >= 0.0f) ? -t : t;
  return norm3(r);
}

__device__ __forceinline__ uint32_t ior_compress(float ior) { return (__float_as_uint((0.5f * (ior - 1.0f)) + 1.0f) >> 15) & 0xFFu; }
__device__ __forceinline__ float ior_decompress(uint32_t c) { return ((__uint_as_float(0x3F800000u | (c << 15)) - 1.0f) * 2.0f) + 1.0f; }

// ---------------------------------------------------------------------------------------------
// material record + quantised shading parameters (cuda/material.cuh)
// ---------------------------------------------------------------------------------------------
struct Mat {
  uint32_t flags;
  float roughness_clamp, roughness, ior;
  float ar, ag, ab, aa;
  C3 emission;
  // texture ids (LB_TEXTURE_NONE = 0xFFFF), only read by the kTex variants
  uint32_t albedo_tex, luminance_tex, roughness_tex, normal_tex, metallic_tex;
  float emission_scale;
};

__device__ __forceinline__ Mat load_material(const uint4* __restrict__ materials, uint32_t id) {  // memory.cuh:442-469
  const uint4 a = __ldg(materials + 2 * id + 0);
  const uint4 b = __ldg(materials + 2 * id + 1);
  const float s = 1.0f / 0xFFFF;
  Mat m;
  m.flags           = a.x & 0xFFu;
  m.roughness_clamp = (a.x & 0xFF00u) * s;  // the reference does not shift the byte down
  m.roughness       = (a.y & 0xFFFFu) * s;
  m.ior             = (a.y >> 16) * s * 2.0f + 1.0f;
  m.ar              = (a.z & 0xFFFFu) * s;
  m.ag              = (a.z >> 16) * s;
  m.ab              = (a.w & 0xFFFFu) * s;
  m.aa              = (a.w >> 16) * s;
  const float scale = __uint_as_float((b.y >> 16) << 15);
  m.emission        = c3((b.x & 0xFFFFu) * s, (b.x >> 16) * s, (b.y & 0xFFFFu) * s) * scale;
  m.emission_scale  = scale;
  m.metallic_tex    = a.x >> 16;
  m.albedo_tex      = b.z & 0xFFFFu;
  m.luminance_tex   = b.z >> 16;
  m.roughness_tex   = b.w & 0xFFFFu;
  m.normal_tex      = b.w >> 16;
  return m;
}

__device__ __forceinline__ float quant_norm(float v, uint32_t maxv) {
  const uint32_t q = (uint32_t) (__saturatef(v) * maxv + 0.5f);
  return q * (1.0f / maxv);
}

__device__ __forceinline__ C3 quant_emission(C3 value) {  // MATERIAL_PARAM_TYPE_COLOR set/get round trip, material.cuh:213-238,292-330
  uint32_t comp;
  float mx, lo, hi;
  if (value.r > value.g && value.r > value.b) {
    comp = 0, mx = value.r, lo = value.g, hi = value.b;
  }
  else if (value.g > value.b) {
    comp = 1, mx = value.g, lo = value.r, hi = value.b;
  }
  else {
    comp = 2, mx = value.b, lo = value.r, hi = value.g;
  }
  mx = __saturatef(mx * (1.0f / 1023.0f)) * 2.0f;
  lo = __saturatef(lo * (2.0f / 1023.0f) * (1.0f / mx));
  hi = __saturatef(hi * (2.0f / 1023.0f) * (1.0f / mx));
  const uint32_t dmax = (__float_as_uint(mx) >= 0x30000000u) ? (__float_as_uint(mx) >> 14) & 0x3FFFu : 0u;
  const uint32_t dlo  = (uint32_t) (lo * 0xFF + 0.5f);
  const uint32_t dhi  = (uint32_t) (hi * 0xFF + 0.5f);
  const float mv      = (dmax > 0) ? __uint_as_float((dmax << 14) | 0x30000000u) * (1023.0f / 2.0f) : 0.0f;
  const float l       = dlo * (1.0f / 0xFF) * mv;
  const float h       = dhi * (1.0f / 0xFF) * mv;
  if (comp == 0)
    return c3(mv, l, h);
  if (comp == 1)
    return c3(l, mv, h);
  return c3(l, h, mv);
}

struct Params {
  uint32_t flags;
  C3 albedo;
  float opacity, roughness, ior;
  C3 emission;
};

struct Ctx {
  uint32_t instance_id, tri_id, prim;
  V3 position, V, normal;
  uint32_t face_normal;
  uint32_t state;
  Params p;
};

// ---------------------------------------------------------------------------------------------
// BSDF (cuda/bsdf_utils.cuh, cuda/bsdf.cuh)
// ---------------------------------------------------------------------------------------------
struct RayCtx {
  V3 V;
  float fresnel, NdotH, NdotL, NdotV, HdotL, HdotV;
  bool is_refraction;
};

__device__ __forceinline__ float bsdf_fresnel(V3 n, V3 V, V3 refr, float ior) {  // bsdf_utils.cuh:79-96
  const float NdotV = dot3(V, n);
  const float NdotT = -dot3(refr, n);
  const float s1 = ior * NdotV, s2 = NdotT;
  const float p1 = ior * NdotT, p2 = NdotV;
  float rs = (s1 - s2) / (s1 + s2);
  float rp = (p1 - p2) / (p1 + p2);
  return __saturatef(0.5f * (rs * rs + rp * rp));
}

__device__ __forceinline__ C3 fresnel_schlick(C3 f0, float f90, float HdotV) {  // :105-118
  const float o  = 1.0f - fabsf(HdotV);
  const float p2 = o * o;
  const float t  = p2 * p2 * o;
  return c3(fmaf(f90 - f0.r, t, f0.r), fmaf(f90 - f0.g, t, f0.g), fmaf(f90 - f0.b, t, f0.b));
}
__device__ __forceinline__ float shadowed_f90(C3 f0) { return fminf(1.0f, (1.0f / 0.04f) * c_lum(f0)); }

__device__ __forceinline__ V3 normal_from_pair(V3 L, V3 V, float ior) {  // :137-143
  const V3 n      = L + V * ior;
  const float len = len3(n);
  return (len > 0.0f) ? n * (1.0f / len) : V;
}

__device__ __forceinline__ float smith_g1(float r4, float NdotS) {
  const float n2 = fmaxf(0.0001f, NdotS * NdotS);
  return 2.0f / (sqrtf(((r4 * (1.0f - n2)) + n2) / n2) + 1.0f);
}
__device__ __forceinline__ float smith_g2(float r4, float NdotL, float NdotV) {
  const float a = NdotV * sqrtf(r4 + NdotL * (NdotL - r4 * NdotL));
  const float b = NdotL * sqrtf(r4 + NdotV * (NdotV - r4 * NdotV));
  return 0.5f / (a + b);
}
__device__ __forceinline__ float smith_g2_over_g1(float r4, float NdotL, float NdotV) {
  const float gv = smith_g1(r4, NdotV), gl = smith_g1(r4, NdotL);
  return gl / (gv + gl - gv * gl);
}
__device__ __forceinline__ float ggx_d(float NdotH, float r4) {
  const float n2 = fminf(NdotH * NdotH, 1.0f);
  const float a  = 1.0f - n2 + r4 * n2;
  return r4 / (PI_F * a * a);
}
__device__ __forceinline__ float pow4(float r) {
  const float r2 = r * r;
  return r2 * r2;
}

// bounded VNDF sampling (Eto & Tokuyoshi 2023), bsdf_utils.cuh:185-203
__device__ __forceinline__ V3 microfacet_sample_normal(V3 V, float roughness, float2 rnd) {
  const float r2 = roughness * roughness, r4 = r2 * r2;
  const V3 v     = norm3(v3(r2 * V.x, r2 * V.y, V.z));
  const float phi = 2.0f * PI_F * rnd.x;
  const float s   = 1.0f + sqrtf(V.x * V.x + V.y * V.y);
  const float s2  = s * s;
  const float k   = (1.0f - r4) * s2 / (s2 + r4 * V.z * V.z);
  const float b   = k * v.z;
  const float z   = (1.0f - rnd.y) * (1.0f + b) - b;
  const float st  = sqrtf(__saturatef(1.0f - z * z));
  const V3 smp    = v3(st * cosf(phi), st * sinf(phi), z) + v;
  return norm3(v3(smp.x * r2, smp.y * r2, smp.z));
}

__device__ __forceinline__ float vndf_norm(V3 V, float r4, float NdotV) {
  const float len2 = r4 * (V.x * V.x + V.y * V.y);
  const float t    = sqrtf(len2 + V.z * V.z);
  const float s    = 1.0f + sqrtf(V.x * V.x + V.y * V.y);
  const float s2   = s * s;
  const float k    = (1.0f - r4) * s2 / (s2 + r4 * V.z * V.z);
  return 2.0f * (k * NdotV + t);
}
__device__ __forceinline__ float microfacet_pdf(V3 V, float roughness, float NdotH, float NdotV) {
  const float r4 = pow4(roughness);
  return ggx_d(NdotH, r4) / vndf_norm(V, r4, NdotV);
}
__device__ __forceinline__ float microfacet_eval(float roughness, float NdotH, float NdotL, float NdotV) {
  const float r4 = pow4(roughness);
  return ggx_d(NdotH, r4) * smith_g2(r4, NdotL, NdotV) * NdotL;
}
__device__ __forceinline__ float microfacet_eval_sampled_microfacet(V3 V, float roughness, float NdotL, float NdotV) {
  const float r4 = pow4(roughness);
  return vndf_norm(V, r4, NdotV) * smith_g2(r4, NdotL, NdotV) * NdotL;
}
__device__ __forceinline__ float microfacet_eval_sampled_diffuse(float roughness, float NdotH, float NdotL, float NdotV) {
  const float r4 = pow4(roughness);
  return ggx_d(NdotH, r4) * smith_g2(r4, NdotL, NdotV) * PI_F;
}
// spherical-cap VNDF sampling (Dupuy & Benyoub 2023), bsdf_utils.cuh:276-288
__device__ __forceinline__ V3 refraction_sample_normal(V3 V, float roughness, float2 rnd) {
  const float r2  = roughness * roughness;
  const V3 v      = norm3(v3(r2 * V.x, r2 * V.y, V.z));
  const float phi = 2.0f * PI_F * rnd.x;
  const float z   = (1.0f - rnd.y) * (1.0f + v.z) - v.z;
  const float st  = sqrtf(__saturatef(1.0f - z * z));
  const V3 smp    = v3(st * cosf(phi), st * sinf(phi), z) + v;
  return norm3(v3(smp.x * r2, smp.y * r2, smp.z));
}
__device__ __forceinline__ float refraction_pdf(float roughness, float NdotH, float NdotV, float HdotV, float HdotL, float ior) {
  const float r4 = pow4(roughness);
  float den      = ior * HdotV + HdotL;
  den            = den * den;
  return ggx_d(NdotH, r4) * smith_g1(r4, NdotV) * (HdotV / NdotV) * (HdotL / den);
}
__device__ __forceinline__ float refraction_eval(float roughness, float HdotL, float HdotV, float NdotH, float NdotL, float NdotV, float ior) {
  const float r4 = pow4(roughness);
  float den      = ior * HdotV + HdotL;
  den            = den * den;
  return 4.0f * NdotL * HdotV * HdotL * ggx_d(NdotH, r4) * smith_g2(r4, NdotL, NdotV) / den;
}
__device__ __forceinline__ float diffuse_pdf(float NdotL) { return __saturatef(NdotL) * (1.0f / PI_F); }
__device__ __forceinline__ float diffuse_eval_sampled_microfacet(V3 V, float roughness, float NdotL, float NdotH, float NdotV) {
  const float r4 = pow4(roughness);
  return NdotL * vndf_norm(V, r4, NdotV) / (PI_F * ggx_d(NdotH, r4));
}

__device__ __forceinline__ float ss_term(Hint hint, float r, const RayCtx& c, float one_over_pdf, float ior_quirk) {
  switch (hint) {
    case H_GENERAL:
      return microfacet_eval(r, c.NdotH, c.NdotL, c.NdotV) * one_over_pdf;
    case H_MICROFACET:
      return microfacet_eval_sampled_microfacet(c.V, r, c.NdotL, c.NdotV);
    case H_DIFFUSE:
      return microfacet_eval_sampled_diffuse(r, c.NdotH, c.NdotL, c.NdotV);
    default:
      return microfacet_eval(r, c.NdotH, c.NdotL, c.NdotV) / refraction_pdf(r, c.NdotH, c.NdotV, c.HdotV, c.HdotL, ior_quirk);
  }
}

// bsdf_multiscattering_evaluate with its three lobes (bsdf_utils.cuh:383-587). The reference's quirks are
// kept on purpose: `ior` of the conductor / glossy refraction-hint terms and of the dielectric lobe is read
// from the ROUGHNESS parameter, and the dielectric DIFFUSE-hint case falls through to the refraction case.
template <int kClass>
__device__ C3 bsdf_multiscattering(const LbLutTexObjects& luts, const Params& p, const RayCtx& c, Hint hint, float one_over_pdf) {
  if (c.NdotL <= 0.0f || c.NdotV <= 0.0f)
    return c3(0.0f, 0.0f, 0.0f);
  const float r           = p.roughness;
  const bool translucent  = ShadeClass<kClass>::translucent(p.flags);
  C3 total                = c3(0.0f, 0.0f, 0.0f);

  if (translucent) {
    const float ior = p.roughness;
    float term      = 0.0f;
    if (c.is_refraction) {
      if (hint == H_GENERAL)
        term = refraction_eval(r, c.HdotL, c.HdotV, c.NdotH, c.NdotL, c.NdotV, ior) * one_over_pdf;
      else if (hint == H_REFRACTION)
        term = smith_g2_over_g1(pow4(r), c.NdotL, c.NdotV);
      term *= (1.0f - c.fresnel);
    }
    else {
      if (hint == H_GENERAL)
        term = microfacet_eval(r, c.NdotH, c.NdotL, c.NdotV) * one_over_pdf;
      else if (hint == H_MICROFACET)
        term = microfacet_eval_sampled_microfacet(c.V, r, c.NdotL, c.NdotV);
      else
        term = microfacet_eval(r, c.NdotH, c.NdotL, c.NdotV) / refraction_pdf(r, c.NdotH, c.NdotV, c.HdotV, c.HdotL, ior);
      term *= c.fresnel;
    }
    const bool use_inv = ior > 1.0f;  // bsdf_dielectric_directional_albedo, :495-505
    const float coord  = use_inv ? (ior - 1.0f) * 0.5f : (1.0f / ior - 1.0f) * 0.5f;
    term /= tex3D<float>(use_inv ? luts.dielectric_inv : luts.dielectric, c.NdotV, r, coord);
    if (ior == 1.0f && c.is_refraction)
      term = (hint == H_REFRACTION) ? 1.0f : 0.0f;
    total = p.albedo * term;
  }
  else if (!c.is_refraction) {
    const float ior = (hint == H_REFRACTION) ? p.roughness : 1.0f;
    const float ss  = ss_term(hint, r, c, one_over_pdf, ior);
    const float cda = tex2D<float>(luts.conductor, c.NdotV, r);
    if (ShadeClass<kClass>::metallic(p.flags)) {
      const C3 f0 = p.albedo;
      const C3 fr = fresnel_schlick(f0, shadowed_f90(f0), c.HdotV);
      total       = fr * ss + f0 * (fr * (((1.0f / cda) - 1.0f) * ss));
    }
    else {
      float diff;
      switch (hint) {
        case H_GENERAL:
          diff = diffuse_pdf(c.NdotL) * one_over_pdf;
          break;
        case H_DIFFUSE:
          diff = 1.0f;
          break;
        case H_MICROFACET:
          diff = diffuse_eval_sampled_microfacet(c.V, r, c.NdotL, c.NdotH, c.NdotV);
          break;
        default:
          diff = diffuse_pdf(c.NdotL) / refraction_pdf(r, c.NdotH, c.NdotV, c.HdotV, c.HdotL, ior);
          break;
      }
      const float gda = tex2D<float>(luts.glossy, c.NdotV, r);
      const C3 f0     = c3(0.04f, 0.04f, 0.04f);
      const C3 fr     = fresnel_schlick(f0, shadowed_f90(f0), c.HdotV);
      total           = fr * (ss / cda) + p.albedo * (diff * (1.0f - gda));
    }
  }
  return ShadeClass<kClass>::may_be_transparent ? total * p.opacity : total;  // opacity == 1 in the opaque classes
}

// kFresnel = false: the caller's class has no translucent material, c.fresnel (only read by the dielectric lobe) is left at 0
template <int kClass>
__device__ __forceinline__ RayCtx evaluate_analyze(const Params& p, V3 normal, V3 V, V3 L) {  // bsdf.cuh:11-50
  RayCtx c;
  c.NdotL         = dot3(normal, L);
  c.NdotV         = __saturatef(dot3(normal, V));
  c.is_refraction = c.NdotL < 0.0f;
  c.NdotL         = fabsf(c.NdotL);
  V3 refr, H;
  bool total_reflection = false;
  if (c.is_refraction) {
    H    = normal_from_pair(L, V, p.ior);
    refr = L;
  }
  else {
    H = normal_from_pair(L, V, 1.0f);
    if (ShadeClass<kClass>::may_be_translucent)
      refr = refract3(V, H, p.ior, total_reflection);
  }
  c.HdotV = fabsf(dot3(H, V));
  c.HdotL = fabsf(dot3(H, L));
  c.NdotH = dot3(normal, H);
  if (c.NdotH < 0.0f) {
    H       = neg3(H);
    c.NdotH = -c.NdotH;
  }
  c.fresnel = 0.0f;
  if (ShadeClass<kClass>::may_be_translucent)
    c.fresnel = total_reflection ? 1.0f : bsdf_fresnel(H, V, refr, p.ior);
  c.V = V;
  return c;
}

template <int kClass>
__device__ __forceinline__ RayCtx sample_context(const Params& p, V3 normal, V3 V, V3 H, V3 L, bool is_refraction) {  // bsdf.cuh:103-133
  RayCtx c;
  c.NdotL         = dot3(normal, L);
  c.NdotV         = __saturatef(dot3(normal, V));
  c.is_refraction = is_refraction;
  c.NdotL         = is_refraction ? -c.NdotL : c.NdotL;
  c.HdotV         = fabsf(dot3(H, V));
  c.HdotL         = fabsf(dot3(H, L));
  c.NdotH         = dot3(normal, H);
  float flip      = 1.0f;
  if (c.NdotH < 0.0f) {
    flip    = -1.0f;
    c.NdotH = -c.NdotH;
  }
  c.fresnel = 0.0f;
  if (ShadeClass<kClass>::may_be_translucent) {
    bool total_reflection = false;
    const V3 refr         = is_refraction ? L : refract3(V, H, p.ior, total_reflection);
    c.fresnel             = total_reflection ? 1.0f : bsdf_fresnel(H * flip, V, refr, p.ior);
  }
  c.V = V;
  return c;
}

template <int kClass>
__device__ __forceinline__ C3 evaluate_core(const LbLutTexObjects& luts, const Params& p, const RayCtx& c, Hint hint, V3 L, V3 face_normal,
                                            float one_over_pdf) {  // bsdf.cuh:52-65
  const float fndl = dot3(face_normal, L);
  const float flip = c.is_refraction ? -1.0f : 1.0f;
  if (fndl * flip < EPS_F)
    return c3(0.0f, 0.0f, 0.0f);
  return bsdf_multiscattering<kClass>(luts, p, c, hint, one_over_pdf);
}

struct Bounce {
  V3 ray;
  C3 weight;
  bool transparent_pass, microfacet_based;
};

// bsdf_sample<MATERIAL_GEOMETRY> with RandomSet::BSDF<0>, bsdf.cuh:135-301
template <int kClass, typename SamplerT>
__device__ Bounce bsdf_sample(const LbLutTexObjects& luts, const Ctx& ctx, const SamplerT& smp) {
  const Params& p = ctx.p;
  Bounce info;

  if (ShadeClass<kClass>::may_be_transparent && p.opacity < 1.0f) {
    if (smp.get1(lbrng::T_BSDF_OPACITY) > p.opacity) {
      info.ray              = neg3(ctx.V);
      info.weight           = (p.flags & MF_COLORED) ? p.albedo : c3(1.0f, 1.0f, 1.0f);
      info.microfacet_based = false;
      info.transparent_pass = true;
      return info;
    }
  }

  const Q4 rot      = rotation_to_z(ctx.normal);
  const V3 V_local  = q_apply(rot, ctx.V);
  const V3 fn_local = q_apply(rot, unpack_normal(ctx.face_normal));
  const V3 up       = v3(0.0f, 0.0f, 1.0f);

  const bool translucent        = ShadeClass<kClass>::translucent(p.flags);
  const bool include_diffuse    = !translucent && !ShadeClass<kClass>::metallic(p.flags);
  const bool include_refraction = translucent;
  const float rough             = p.roughness;
  const float ior               = p.ior;

  float resampling = smp.get1(lbrng::T_BSDF_RESAMPLING);
  float sum_weights;
  C3 selected_eval;
  V3 ray_local;
  info.transparent_pass = false;
  info.microfacet_based = true;

  {
    const V3 m       = microfacet_sample_normal(V_local, rough, smp.get2(lbrng::T_BSDF_REFLECTION));
    const V3 r       = reflect3(V_local, m);
    const RayCtx c   = sample_context<kClass>(p, up, V_local, m, r, false);
    const C3 ev      = evaluate_core<kClass>(luts, p, c, H_MICROFACET, r, fn_local, 1.0f);
    const float pdf  = microfacet_pdf(V_local, rough, c.NdotH, c.NdotV);
    const float dpdf = include_diffuse ? diffuse_pdf(c.NdotL) : 0.0f;
    const float rpdf = include_refraction ? refraction_pdf(rough, c.NdotH, c.NdotV, c.HdotV, c.HdotL, ior) : 0.0f;
    const float sum  = pdf + dpdf + rpdf;
    const float mis  = (sum > 0.0f) ? pdf / sum : 0.0f;
    ray_local        = r;
    sum_weights      = c_max(ev) * mis;
    selected_eval    = ev;
  }

  if (include_diffuse) {
    const float2 rnd = smp.get2(lbrng::T_BSDF_DIFFUSE);
    V3 r;  // sample_ray_sphere, math.cuh:339-358
    if (fabsf(rnd.x) > 1.0f - EPS_F)
      r = v3(0.0f, 0.0f, copysignf(1.0f, rnd.x));
    else {
      const float a = sqrtf(1.0f - rnd.x * rnd.x);
      const float b = 2.0f * PI_F * rnd.y;
      r             = v3(a * cosf(b), a * sinf(b), rnd.x);
    }
    const V3 m       = norm3(V_local + r);
    const RayCtx c   = sample_context<kClass>(p, up, V_local, m, r, false);
    const C3 ev      = evaluate_core<kClass>(luts, p, c, H_DIFFUSE, r, fn_local, 1.0f);
    const float pdf  = diffuse_pdf(c.NdotL);
    const float mpdf = microfacet_pdf(V_local, rough, c.NdotH, c.NdotV);
    const float rpdf = include_refraction ? refraction_pdf(rough, c.NdotH, c.NdotV, c.HdotV, c.HdotL, ior) : 0.0f;
    const float sum  = pdf + mpdf + rpdf;
    const float mis  = (sum > 0.0f) ? pdf / sum : 0.0f;
    const float w    = c_max(ev) * mis;
    sum_weights += w;
    const float prob = w / sum_weights;
    if (resampling < prob) {
      ray_local             = r;
      selected_eval         = ev;
      info.transparent_pass = false;
      info.microfacet_based = false;
      resampling            = lbrng::saturate_random(resampling / prob);
    }
    else
      resampling = lbrng::saturate_random((resampling - prob) / (1.0f - prob));
  }

  if (include_refraction) {
    bool total_reflection;
    const V3 m     = refraction_sample_normal(V_local, rough, smp.get2(lbrng::T_BSDF_REFRACTION));
    const V3 r     = refract3(V_local, m, ior, total_reflection);
    const RayCtx c = sample_context<kClass>(p, up, V_local, m, r, !total_reflection);
    const C3 ev    = evaluate_core<kClass>(luts, p, c, H_REFRACTION, r, fn_local, 1.0f);
    float mis      = 1.0f;
    if (total_reflection) {
      const float pdf  = refraction_pdf(rough, c.NdotH, c.NdotV, c.HdotV, c.HdotL, ior);
      const float rpdf = microfacet_pdf(V_local, rough, c.NdotH, c.NdotV);
      const float dpdf = include_diffuse ? diffuse_pdf(c.NdotL) : 0.0f;
      const float sum  = pdf + rpdf + dpdf;
      mis              = (sum > 0.0f) ? pdf / sum : 0.0f;
    }
    const float w = c_max(ev) * mis;
    sum_weights += w;
    const float prob = w / sum_weights;
    if (resampling < prob) {
      ray_local             = r;
      selected_eval         = ev;
      info.transparent_pass = !total_reflection;
      info.microfacet_based = true;
    }
  }

  info.weight = (sum_weights > 0.0f) ? selected_eval * (sum_weights / c_max(selected_eval)) : c3(0.0f, 0.0f, 0.0f);
  info.ray    = norm3(q_apply_inv(rot, ray_local));
  return info;
}

// ---------------------------------------------------------------------------------------------
// RIS reservoir (cuda/ris.cuh:22-84)
// ---------------------------------------------------------------------------------------------
struct Reservoir {
  float sum_weight, selected_target, random;
};
__device__ __forceinline__ bool reservoir_add(Reservoir& r, float target, float sampling_weight) {
  const float weight = target * sampling_weight;
  r.sum_weight += weight;
  if (weight == 0.0f)
    return false;
  const float prob    = weight / r.sum_weight;
  const bool accepted = r.random < prob;
  r.selected_target   = accepted ? target : r.selected_target;
  const float shift   = accepted ? 0.0f : prob;
  const float scale   = accepted ? prob : 1.0f - prob;
  r.random            = lbrng::saturate_random((r.random - shift) / scale);
  return accepted;
}
__device__ __forceinline__ float reservoir_weight(const Reservoir& r) { return (r.selected_target > 0.0f) ? r.sum_weight / r.selected_target : 0.0f; }

// ---------------------------------------------------------------------------------------------
// light tree (cuda/light_tree.cuh). Records are read as 16-byte words:
//   root header uint4: {x | y<<16, z | num_root_lights<<16, power_norm | num_sections<<16, exps}
//   root section 3 x uint4: rel_mean_x[8], rel_mean_y[8] | rel_mean_z[8], rel_std_dev[8] | rel_power[8] (u16)
//   node 4 x uint4: {x|y, z|pad, exps, num_lights} {child_ptr, light_ptr, mean_x[8]} {mean_y[8], mean_z[8]} {std_dev[8], power[8]}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf16(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float exp_i8(uint32_t byte) { return exp2f((float) (int8_t) byte); }
__device__ __forceinline__ float byte_of(uint2 v, uint32_t i) { return (float) (((i < 4) ? (v.x >> (8 * i)) : (v.y >> (8 * (i - 4)))) & 0xFFu); }

template <int kClass>
__device__ __forceinline__ float tree_importance(const Ctx& ctx, float power, V3 mean, float std_dev) {  // :68-89
  const V3 PO     = mean - ctx.position;
  const float d2  = dot3(PO, PO);
  const float var = std_dev * std_dev;
  const float inv = 1.0f / (d2 + var);
  float result    = power * inv;
  if (ShadeClass<kClass>::translucent(ctx.p.flags))
    return result;
  const float t     = var * inv;
  const float NdotL = __saturatef(dot3(PO, ctx.normal) * sqrtf(inv));
  return result * (NdotL * (1.0f - t) + t);
}

template <int kClass>
__device__ __forceinline__ float child_importance(const Ctx& ctx, float power, float rel_std, float mx, float my, float mz, V3 base, V3 ex,
                                                  float exp_v) {
  if (power == 0.0f)
    return 0.0f;
  const V3 mean = v3(mx, my, mz) * ex + base;
  return fmaxf(tree_importance<kClass>(ctx, power, mean, rel_std * exp_v), 0.0f);
}

struct TreeWork {
  uint32_t cont[NUM_TREE_LANES];  // LightTreeContinuation: is_light (1) | child_index (8) << 1 | probability (20) << 9
  float root_sum;
};

// Root children decoded once per light-tree upload (k_unpack_light_root): every path evaluates all of them, and the
// byte / u16 extraction, int -> float conversion and the mean reconstruction are identical for all 32 lanes of a
// warp and for every path. Two float4 per child: {mean.xyz, std_dev}, {power, 0, 0, 0}; same arithmetic as
// child_importance's inline decode, so the values are unchanged.
__global__ void __launch_bounds__(128) k_unpack_light_root(const uint4* __restrict__ root, float4* __restrict__ out) {
  const uint4 h               = root[0];
  const uint32_t num_sections = (h.z >> 16) & 0xFFu;
  const uint32_t c            = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= num_sections * 8u)
    return;
  const V3 base     = v3(bf16(h.x & 0xFFFFu), bf16(h.x >> 16), bf16(h.y & 0xFFFFu));
  const V3 ex       = v3(exp_i8(h.w & 0xFFu), exp_i8((h.w >> 8) & 0xFFu), exp_i8((h.w >> 16) & 0xFFu));
  const float exp_v = exp_i8(h.w >> 24);
  const uint32_t s = c >> 3, k = c & 7u;
  const uint4 a = root[1 + 3 * s + 0], b = root[1 + 3 * s + 1], w = root[1 + 3 * s + 2];
  const uint32_t pw_word = (k < 2) ? w.x : (k < 4) ? w.y : (k < 6) ? w.z : w.w;
  const float power      = (float) ((pw_word >> (16 * (k & 1))) & 0xFFFFu);
  const V3 mean          = v3(byte_of(make_uint2(a.x, a.y), k), byte_of(make_uint2(a.z, a.w), k), byte_of(make_uint2(b.x, b.y), k)) * ex + base;
  const float std_dev    = byte_of(make_uint2(b.z, b.w), k) * exp_v;
  out[2 * c + 0]         = make_float4(mean.x, mean.y, mean.z, std_dev);
  out[2 * c + 1]         = make_float4(power, 0.0f, 0.0f, 0.0f);
}

void lb_launch_unpack_light_root(const void* root, float4* out, uint32_t num_sections, cudaStream_t s) {
  if (num_sections)
    k_unpack_light_root<<<(num_sections * 8u + 127u) / 128u, 128, 0, s>>>((const uint4*) root, out);
}

// light_tree_traverse_prepass (light_tree.cuh:191-262): 8 independent reservoir lanes stream over the root children.
// Restated for throughput with unchanged semantics:
//   * the child loop is NOT unrolled (8 children x 8 lanes of straight-line code made k_shade stall on instruction
//     fetch) and reads the decoded children;
//   * lane update u' = accepted ? u / p : (u - p) / (1 - p) is evaluated as accepted ? u * (1/p) : fma(u, 1/(1-p), -p/(1-p))
//     with the three per-child constants hoisted: 2 FMA-pipe + 3 ALU-pipe instructions per lane instead of 8. The
//     clamp of the reference's saturate_random is dropped: u' < 1 up to one rounding and u' only feeds `u' < p` tests;
//   * a lane keeps only the INDEX of its selected child (one byte each, two registers for 8 lanes); the target value
//     of the final selection is re-evaluated once after the loop instead of being carried through every update.
//   * the decoded children are STAGED IN SHARED MEMORY once per block (k_shade: s_child_mean / s_child_power, at most 128 x 20 bytes):
//     every thread of the block streams over the same records, so the per-child loads become conflict-free broadcasts
//     (LDS, ~25 cycles) instead of two dependent global loads per child and warp whose latency sat on the critical path of
//     the loop (ncu source page, round 2: 2.8 % of the kernel's stall samples on the first use of the child record).
struct RootChildrenShared {  // staged by the block (default)
  const float4* mean_std;
  const float* power_of;
  __device__ __forceinline__ float4 mean(uint32_t c) const { return mean_std[c]; }
  __device__ __forceinline__ float power(uint32_t c) const { return power_of[c]; }
};
struct RootChildrenGlobal {  // round-1 path, kept for the A/B measurement (LB_STAGE_ROOT_CHILDREN=0)
  const float4* records;
  __device__ __forceinline__ float4 mean(uint32_t c) const { return __ldg(records + 2 * c + 0); }
  __device__ __forceinline__ float power(uint32_t c) const { return __ldg(&records[2 * c + 1].x); }
};

#ifndef LB_ROOT_PIPELINE
#define LB_ROOT_PIPELINE 1
#endif
#ifndef LB_ROOT_FFMA2
#define LB_ROOT_FFMA2 0
#endif
#ifndef LB_SHADE_PREFETCH
#define LB_SHADE_PREFETCH 0
#endif
template <int kClass, typename SamplerT, typename ChildrenT>
__device__ void tree_prepass(const uint4* __restrict__ root, const ChildrenT children, const Ctx& ctx, const SamplerT& smp, TreeWork& work) {
  const uint4 h               = __ldg(root);
  const uint32_t num_lights   = h.y >> 16;
  const uint32_t num_children = ((h.z >> 16) & 0xFFu) * 8u;

  float lane_random[NUM_TREE_LANES];
  uint32_t selected[NUM_TREE_LANES];
#pragma unroll
  for (int l = 0; l < NUM_TREE_LANES; l++) {
    lane_random[l] = smp.get1(lbrng::T_LIGHT_GEO_TREE_PREPASS + l);
    selected[l]    = 0xFFFFFFFFu;
  }
  float agg = 0.0f, sum = 0.0f;

#if LB_ROOT_PIPELINE
  // Software-pipelined, branch-free form: the importance of child c + 1 (an FMA chain with two MUFU results, independent of the lanes)
  // is evaluated while the eight lanes consume child c, so that its latency hides behind the 32 update instructions instead of
  // heading every iteration (ncu source page, round 2: stall_wait + stall_short_scoreboard = 38 % of k_shade's samples at 5 warps
  // per scheduler). Children without power or importance need no branch: target = 0 gives prob = 0, 1 / (1 - prob) = 1 and
  // -prob / (1 - prob) = -0, the accept test u < 0 fails and fma(u, 1, -0) returns u unchanged; agg and sum receive + 0.
  float target_next = 0.0f;
  if (num_children) {
    const float4 m = children.mean(0);
    target_next    = fmaxf(tree_importance<kClass>(ctx, children.power(0), v3(m.x, m.y, m.z), m.w), 0.0f);
  }
#ifndef LB_ROOT_UNROLL
#define LB_ROOT_UNROLL 1
#endif
  constexpr int kRootUnroll = LB_ROOT_UNROLL;
#pragma unroll kRootUnroll
  for (uint32_t c = 0; c < num_children; c++) {
    const float target = target_next;
    {
      const uint32_t cn = min(c + 1u, num_children - 1u);
      const float4 m    = children.mean(cn);
      target_next       = fmaxf(tree_importance<kClass>(ctx, children.power(cn), v3(m.x, m.y, m.z), m.w), 0.0f);
    }
    agg += target;
    const float prob = (target > 0.0f) ? target / agg : 0.0f;
    sum += (prob != 0.0f) ? target : 0.0f;
    const float inv_p = __fdividef(1.0f, prob);
    const float inv_q = __fdividef(1.0f, fmaxf(1.0f - prob, 1e-30f));
    const float off_q = -prob * inv_q;
#if LB_ROOT_FFMA2
    // rejection update of two lanes with one packed FMA (fma.rn.ftz.f32x2), acceptance overrides per lane: 3.5 instead of 4 instructions per lane
#pragma unroll
    for (int l = 0; l < NUM_TREE_LANES; l += 2) {
      asm("{\n\t"
          ".reg .pred a0, a1;\n\t"
          ".reg .b64 u2, q2, o2, r2;\n\t"
          ".reg .f32 r0, r1;\n\t"
          "mov.b64 u2, {%0, %1};\n\t"
          "mov.b64 q2, {%6, %6};\n\t"
          "mov.b64 o2, {%7, %7};\n\t"
          "fma.rn.ftz.f32x2 r2, u2, q2, o2;\n\t"
          "mov.b64 {r0, r1}, r2;\n\t"
          "setp.lt.ftz.f32 a0, %0, %4;\n\t"
          "setp.lt.ftz.f32 a1, %1, %4;\n\t"
          "@a0 mul.ftz.f32 r0, %0, %5;\n\t"
          "@a1 mul.ftz.f32 r1, %1, %5;\n\t"
          "@a0 mov.u32 %2, %8;\n\t"
          "@a1 mov.u32 %3, %8;\n\t"
          "mov.f32 %0, r0;\n\t"
          "mov.f32 %1, r1;\n\t"
          "}"
          : "+f"(lane_random[l]), "+f"(lane_random[l + 1]), "+r"(selected[l]), "+r"(selected[l + 1])
          : "f"(prob), "f"(inv_p), "f"(inv_q), "f"(off_q), "r"(c));
    }
#else
#pragma unroll
    for (int l = 0; l < NUM_TREE_LANES; l++) {
      asm("{\n\t"
          ".reg .pred acc;\n\t"
          "setp.lt.ftz.f32 acc, %0, %2;\n\t"
          "@acc mul.ftz.f32 %0, %0, %3;\n\t"
          "@!acc fma.rn.ftz.f32 %0, %0, %4, %5;\n\t"
          "@acc mov.u32 %1, %6;\n\t"
          "}"
          : "+f"(lane_random[l]), "+r"(selected[l])
          : "f"(prob), "f"(inv_p), "f"(inv_q), "f"(off_q), "r"(c));
    }
#endif
  }
#else
#pragma unroll 1
  for (uint32_t c = 0; c < num_children; c++) {
    const float pw  = children.power(c);
    if (pw == 0.0f)
      continue;
    const float4 m     = children.mean(c);
    const float target = fmaxf(tree_importance<kClass>(ctx, pw, v3(m.x, m.y, m.z), m.w), 0.0f);
    // ris_aggregator_add_sample + ris_lane_add_sample, ris.cuh:114-151
    agg += target;
    const float prob = (target > 0.0f) ? target / agg : 0.0f;
    if (prob == 0.0f)
      continue;
    sum += target;
    const float inv_p = __fdividef(1.0f, prob);
    const float inv_q = __fdividef(1.0f, fmaxf(1.0f - prob, 1e-30f));
    const float off_q = -prob * inv_q;
    // Predicated in-place update, 4 instructions per lane. The C++ select form compiled to FSETP + FMUL + @P FFMA + SEL plus
    // two MOVs per lane that copy the temporaries back into the loop-carried registers (cuobjdump -sass, 48 instead of 32
    // instructions per child); same arithmetic (--use_fast_math: .ftz), same results.
#pragma unroll
    for (int l = 0; l < NUM_TREE_LANES; l++) {
      asm("{\n\t"
          ".reg .pred acc;\n\t"
          "setp.lt.ftz.f32 acc, %0, %2;\n\t"
          "@acc mul.ftz.f32 %0, %0, %3;\n\t"
          "@!acc fma.rn.ftz.f32 %0, %0, %4, %5;\n\t"
          "@acc mov.u32 %1, %6;\n\t"
          "}"
          : "+f"(lane_random[l]), "+r"(selected[l])
          : "f"(prob), "f"(inv_p), "f"(inv_q), "f"(off_q), "r"(c));
    }
  }

#endif

  work.root_sum = sum * (bf16(h.z & 0xFFFFu) / 0xFFFF);
  // unrolled: a rolled loop indexes selected[] dynamically, which forces the array into consecutive registers + local memory and
  // made the child loop above copy every selected[l] once per child (8 MOVs per child, cuobjdump -sass)
#pragma unroll
  for (int l = 0; l < NUM_TREE_LANES; l++) {
    float lane_target = 0.0f;
    uint32_t sel      = selected[l];
    if (sel != 0xFFFFFFFFu) {
      const float4 m = children.mean(sel);
      lane_target    = fmaxf(tree_importance<kClass>(ctx, children.power(sel), v3(m.x, m.y, m.z), m.w), 0.0f);
    }
    else
      sel = 0;
    const bool is_light  = sel < num_lights;
    const uint32_t index = (is_light ? sel : sel - num_lights) & 0xFFu;
    const float prob     = (agg > 0.0f) ? lane_target / agg : 0.0f;
    uint32_t q           = 0;
    if (prob > 0.0f)
      q = max((uint32_t) ((0xFFFFF * prob) + 0.5f), 1u);
    work.cont[l] = (is_light ? 1u : 0u) | (index << 1) | ((q & 0xFFFFFu) << 9);
  }
}

template <int kClass, bool kCount, typename SamplerT>
__device__ void tree_postpass(const uint4* __restrict__ nodes, const Ctx& ctx, const SamplerT& smp, uint32_t lane, uint32_t cont,
                              uint32_t& light_id, float& weight, uint32_t& nodes_visited) {  // :264-320
  const float prob = (cont >> 9) * (1.0f / 0xFFFFF) * NUM_TREE_LANES;
  light_id         = LB_LIGHT_ID_INVALID;
  weight           = (prob > 0.0f) ? 1.0f / prob : 0.0f;
  if (prob == 0.0f)
    return;
  const uint32_t index = (cont >> 1) & 0xFFu;
  if (cont & 1u) {
    light_id = index;
    return;
  }
  uint32_t node_index = index;
  Reservoir res;
  res.sum_weight = 0.0f, res.selected_target = 0.0f;
  res.random = smp.get1(lbrng::T_LIGHT_GEO_TREE_POSTPASS + lane);

#pragma unroll 1
  for (int guard = 0; guard < 64; guard++) {
    if (kCount)
      nodes_visited++;
    const uint4 n0 = __ldg(nodes + 4 * (size_t) node_index + 0);
    const uint4 n1 = __ldg(nodes + 4 * (size_t) node_index + 1);
    const uint4 n2 = __ldg(nodes + 4 * (size_t) node_index + 2);
    const uint4 n3 = __ldg(nodes + 4 * (size_t) node_index + 3);
    const V3 base     = v3(bf16(n0.x & 0xFFFFu), bf16(n0.x >> 16), bf16(n0.y & 0xFFFFu));
    const V3 ex       = v3(exp_i8(n0.z & 0xFFu), exp_i8((n0.z >> 8) & 0xFFu), exp_i8((n0.z >> 16) & 0xFFu));
    const float exp_v = exp_i8(n0.z >> 24);
    const uint32_t num_lights = n0.w & 0xFFu;
    uint32_t sel = 0xFFu;
    unsigned long long w_pw = ((unsigned long long) n3.w << 32) | n3.z, w_sd = ((unsigned long long) n3.y << 32) | n3.x;
    unsigned long long w_mx = ((unsigned long long) n1.w << 32) | n1.z, w_my = ((unsigned long long) n2.y << 32) | n2.x;
    unsigned long long w_mz = ((unsigned long long) n2.w << 32) | n2.z;
#pragma unroll 1
    for (uint32_t k = 0; k < 8; k++) {
      const float target =
        child_importance<kClass>(ctx, (float) (uint32_t) (w_pw & 0xFFull), (float) (uint32_t) (w_sd & 0xFFull), (float) (uint32_t) (w_mx & 0xFFull),
                         (float) (uint32_t) (w_my & 0xFFull), (float) (uint32_t) (w_mz & 0xFFull), base, ex, exp_v);
      w_pw >>= 8, w_sd >>= 8, w_mx >>= 8, w_my >>= 8, w_mz >>= 8;
      if (reservoir_add(res, target, 1.0f))
        sel = k;
    }
    if (sel == 0xFFu)
      return;
    weight *= reservoir_weight(res);
    if (sel < num_lights) {
      light_id = n1.y + sel;
      return;
    }
    node_index          = n1.x + (sel - num_lights);
    res.sum_weight      = 0.0f;
    res.selected_target = 0.0f;
  }
}

// ---------------------------------------------------------------------------------------------
// triangle lights (cuda/light_triangle.cuh)
// ---------------------------------------------------------------------------------------------
// Per-light record, built once per scene upload (k_build_light_records) instead of being re-derived for each of the up
// to 9 lights a path touches: the reference's light_load (light_triangle.cuh:33-72) walks
// handle -> instance -> mesh pointers -> vertices / textri -> material, five dependent loads plus a quaternion
// transform. The record holds the results of exactly that arithmetic (same translation unit, same flags):
//   r0 = {vertex.xyz, material id | bidirectional << 16}   r1 = {edge1.xyz, flattened prim of the light}
//   r2 = {edge2.xyz, 1 if the colour is textured}          r3 = {light colour rgb (light_get_color, untextured part), 0}
// Emitters whose material has a luminance or albedo texture (r2.w != 0) take the reference's route at shading time:
// texture coordinates of the sampled point (light_triangle_sample_finalize_dist_and_uvs, :74-90) + light_get_color.
#define LB_LIGHT_RECORD_FLOAT4S 4

__global__ void __launch_bounds__(128) k_build_light_records(LbShadeParams P, float4* __restrict__ records) {
  const uint32_t light_id = blockIdx.x * blockDim.x + threadIdx.x;
  if (light_id >= P.num_lights)
    return;
  const uint2 handle   = P.light_handles[light_id];
  const uint32_t mesh  = P.instance_mesh[handle.x];
  const LbTransform tr = P.instance_xform[handle.x];
  const float4* vb     = P.mesh_vertices[mesh];
  const float4 a = vb[3 * (size_t) handle.y + 0], b = vb[3 * (size_t) handle.y + 1], c = vb[3 * (size_t) handle.y + 2];
  const V3 v0            = v3(a.x, a.y, a.z);
  const V3 vertex        = transform_point(tr, v0);
  const V3 edge1         = transform_relative(tr, v3(b.x, b.y, b.z) - v0);
  const V3 edge2         = transform_relative(tr, v3(c.x, c.y, c.z) - v0);
  const uint32_t mid     = P.mesh_textris[mesh][handle.y].w & 0xFFFFu;
  const Mat m            = load_material(P.materials, mid);
  const uint32_t bidir   = (m.flags & DMF_BIDIRECTIONAL) ? 1u : 0u;
  C3 col                 = m.emission;  // light_get_color, light_triangle.cuh:244-280 (untextured)
  if (c_any(col))
    col = col * m.aa;
  float4* r = records + LB_LIGHT_RECORD_FLOAT4S * (size_t) light_id;
  r[0]      = make_float4(vertex.x, vertex.y, vertex.z, __uint_as_float(mid | (bidir << 16)));
  r[1]      = make_float4(edge1.x, edge1.y, edge1.z, __uint_as_float(P.light_prims[light_id]));
  const bool textured = (m.luminance_tex != LB_TEXTURE_NONE) || (m.albedo_tex != LB_TEXTURE_NONE);
  r[2]      = make_float4(edge2.x, edge2.y, edge2.z, textured ? 1.0f : 0.0f);
  r[3]      = make_float4(col.r, col.g, col.b, 0.0f);
}

void lb_launch_build_light_records(const LbShadeParams& sp, float4* records, cudaStream_t s) {
  if (sp.num_lights)
    k_build_light_records<<<(sp.num_lights + 127u) / 128u, 128, 0, s>>>(sp, records);
}

struct TriLight {
  V3 vertex, edge1, edge2;
  C3 color;
  uint32_t material_id;
  uint32_t prim;  // flattened primitive index of the emitter
  bool bidirectional;
  bool textured;  // colour depends on the texture coordinates of the sampled point
};

__device__ __forceinline__ TriLight light_init(const LbShadeParams& P, uint32_t light_id) {
  const float4* r = P.light_records + LB_LIGHT_RECORD_FLOAT4S * (size_t) light_id;
  const float4 r0 = __ldg(r + 0), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
  TriLight L;
  L.vertex        = v3(r0.x, r0.y, r0.z);
  L.edge1         = v3(r1.x, r1.y, r1.z);
  L.edge2         = v3(r2.x, r2.y, r2.z);
  L.color         = c3(r3.x, r3.y, r3.z);
  L.material_id   = __float_as_uint(r0.w) & 0xFFFFu;
  L.bidirectional = (__float_as_uint(r0.w) >> 16) != 0;
  L.prim          = __float_as_uint(r1.w);
  L.textured      = r2.w != 0.0f;
  return L;
}

__device__ __forceinline__ float light_intersect(const TriLight& L, V3 origin, V3 ray, float2& coords) {  // light_triangle_intersection_uv, :10-31
  const V3 h    = cross3(ray, L.edge2);
  const float a = dot3(L.edge1, h);
  const float f = 1.0f / a;
  const V3 s    = origin - L.vertex;
  const float u = f * dot3(s, h);
  const V3 q    = cross3(s, L.edge1);
  const float v = f * dot3(ray, q);
  coords        = make_float2(u, v);
  if (v < 0.0f || u < 0.0f || !(u + v <= 1.0f))
    return FLT_MAX;
  const float t = f * dot3(L.edge2, q);
  return (t >= 0.0f) ? t : FLT_MAX;
}

__device__ __forceinline__ float light_solid_angle(const TriLight& L, V3 origin) {  // :92-106
  const V3 v0    = norm3(L.vertex - origin);
  const V3 v1    = norm3((L.vertex + L.edge1) - origin);
  const V3 v2    = norm3((L.vertex + L.edge2) - origin);
  const float G0 = fabsf(dot3(cross3(v0, v1), v2));
  const float G1 = dot3(v0, v2) + dot3(v1, v2);
  const float G2 = 1.0f + dot3(v0, v1);
  return 2.0f * atan2f(G0, G1 + G2);
}

__device__ __forceinline__ float light_area(const TriLight& L) { return len3(cross3(L.edge1, L.edge2)) * 0.5f; }
__device__ __forceinline__ bool non_finite(float a) { return isnan(a) || isinf(a); }

// solid angle sampling (Peters 2021), light_triangle.cuh:112-158
__device__ bool light_sample_solid_angle(const TriLight& L, V3 origin, float2 rnd, V3& ray, float& solid_angle) {
  const V3 v0     = norm3(L.vertex - origin);
  const V3 v1     = norm3((L.vertex + L.edge1) - origin);
  const V3 v2     = norm3((L.vertex + L.edge2) - origin);
  const float G0s = dot3(cross3(v0, v1), v2);
  if (!L.bidirectional && G0s >= 0.0f)
    return false;
  const float G0 = fabsf(G0s);
  const float G1 = dot3(v0, v2) + dot3(v1, v2);
  const float G2 = 1.0f + dot3(v0, v1);
  solid_angle    = 2.0f * atan2f(G0, G1 + G2);
  if (non_finite(solid_angle) || solid_angle < 1e-7f)
    return false;
  const float ssa = rnd.x * solid_angle;
  const float sn = sinf(0.5f * ssa), cs = cosf(0.5f * ssa);
  const V3 r     = v0 * (G0 * cs - G1 * sn) + v2 * (G2 * sn);
  const V3 v2t   = r * (2.0f * dot3(v0, r) / dot3(r, r)) - v0;
  const float s2 = dot3(v1, v2t);
  const float s  = (1.0f - rnd.y) + rnd.y * s2;
  const float t  = sqrtf(fmaxf((1.0f - s * s) / (1.0f - s2 * s2), 0.0f));
  ray            = norm3(v1 * (s - t * s2) + v2t * t);
  return !(non_finite(ray.x) || non_finite(ray.y) || non_finite(ray.z));
}

__device__ __forceinline__ LbTexScene tex_scene(const LbShadeParams& P) {
  LbTexScene T;
  T.textures      = P.textures;
  T.num_textures  = P.num_textures;
  T.materials     = P.materials;
  T.prim_handle   = P.prim_handle;
  T.instance_mesh = P.instance_mesh;
  T.mesh_textris  = P.mesh_textris;
  T.prim_material = P.prim_material;
  T.shadow_tab    = nullptr;  // only read by lb_alpha_cutout (traversal kernels)
  return T;
}

// light_get_color, light_triangle.cuh:244-280; coords = barycentrics of the sampled point on the emitter
template <bool kTex>
__device__ __forceinline__ C3 light_color_of(const LbShadeParams& P, const TriLight& L, float2 coords) {
  if (!kTex || !L.textured)
    return L.color;
  const Mat m     = load_material(P.materials, L.material_id);
  const float2 uv = lb_lerp_uv(lb_prim_textri(tex_scene(P), L.prim), coords.x, coords.y);
  C3 col          = m.emission;
  if (m.luminance_tex != LB_TEXTURE_NONE) {
    const float4 e = lb_texture_load(P.textures, P.num_textures, m.luminance_tex, uv, true, true, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    col            = c3(e.x, e.y, e.z) * m.emission_scale;
  }
  if (c_any(col)) {
    float alpha = m.aa;
    if (m.albedo_tex != LB_TEXTURE_NONE)
      alpha = lb_texture_load(P.textures, P.num_textures, m.albedo_tex, uv, true, true, make_float4(0.0f, 0.0f, 0.0f, 1.0f)).w;
    col = col * alpha;
  }
  return col;
}

// ---------------------------------------------------------------------------------------------
// BSDF-sampled light direction + MIS (cuda/light_bsdf.cuh, mis.cuh)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float lbsdf_sampling_roughness(float r) { return r + 0.04f * (1.0f - r); }
__device__ __forceinline__ float lbsdf_rr_probability(float r) { return __saturatef((r - 0.5f) / (0.1f - 0.5f)); }

template <int kClass>
__device__ float light_bsdf_probability(const Ctx& ctx, V3 L) {  // light_bsdf.cuh:106-146
  const Params& p  = ctx.p;
  const Q4 rot     = rotation_to_z(ctx.normal);
  const V3 V_local = norm3(q_apply(rot, ctx.V));
  const V3 L_local = norm3(q_apply(rot, L));
  const bool include_refraction = ShadeClass<kClass>::translucent(p.flags);
  const float refraction_prob   = include_refraction ? 0.5f : 0.0f;
  const RayCtx c = evaluate_analyze<kClass>(p, v3(0.0f, 0.0f, 1.0f), V_local, L_local);
  const float sr = lbsdf_sampling_roughness(p.roughness);
  float prob;
  // the refraction branch is kept for every class: 0 x refraction_pdf() is NaN when the pdf overflows, and the reference's NEE
  // sample of that vertex is then NaN as well (QUIRK)
  if (c.is_refraction)
    prob = refraction_prob * refraction_pdf(sr, c.NdotH, c.NdotV, c.HdotV, c.HdotL, p.ior);
  else
    prob = (1.0f - refraction_prob) * microfacet_pdf(V_local, sr, c.NdotH, c.NdotV);
  return prob * lbsdf_rr_probability(p.roughness);
}

__device__ __forceinline__ float mis_weight_base(float gi_pdf, float solid_angle, float power, float dist_sq, float root_sum) {  // mis.cuh:19-24
  const float dl_pdf = NUM_TREE_LANES * (1.0f / solid_angle) * (power / dist_sq) * (1.0f / root_sum);
  return (dl_pdf > 0.0f) ? gi_pdf / (gi_pdf + dl_pdf) : 1.0f;
}

// ---------------------------------------------------------------------------------------------
// geometry_get_context, cuda/geometry_utils.cuh:54-221. kTex = false: the scene has no textured material, all texture
// branches compile away.
// ---------------------------------------------------------------------------------------------
template <bool kTex>
__device__ Ctx get_context(const LbShadeParams& P, uint32_t prim, V3 hit_point, V3 ray_world, uint32_t state, uint32_t medium_ior) {
  const uint2 handle   = __ldg(P.prim_handle + prim);
  const uint32_t mesh  = __ldg(P.instance_mesh + handle.x);
  const LbTransform tr = P.instance_xform[handle.x];
  const float4* vb     = P.mesh_vertices[mesh];
  const float4 a = __ldg(vb + 3 * (size_t) handle.y + 0), b = __ldg(vb + 3 * (size_t) handle.y + 1), c = __ldg(vb + 3 * (size_t) handle.y + 2);
  const uint4 tt = __ldg(P.mesh_textris[mesh] + handle.y);

  const V3 vertex = v3(a.x, a.y, a.z);
  const V3 edge1  = v3(b.x, b.y, b.z) - vertex;
  const V3 edge2  = v3(c.x, c.y, c.z) - vertex;

  V3 position  = transform_point_inv(tr, hit_point);
  const V3 ray = transform_rotate_inv(tr, ray_world);

  V3 face_normal = norm3(cross3(edge1, edge2));

  // get_coordinates_in_triangle, math.cuh:203-213
  const V3 diff     = position - vertex;
  const float d00   = dot3(edge1, edge1), d01 = dot3(edge1, edge2), d11 = dot3(edge2, edge2);
  const float d20   = dot3(diff, edge1), d21 = dot3(diff, edge2);
  const float denom = 1.0f / (d00 * d11 - d01 * d01);
  const float cu    = (d11 * d20 - d01 * d21) * denom;
  const float cv    = (d00 * d21 - d01 * d20) * denom;

  position = vertex + (edge1 * cu + edge2 * cv);
  position = transform_point(tr, position);

  const Mat mat = load_material(P.materials, tt.w & 0xFFFFu);

  const V3 n0  = unpack_normal(__float_as_uint(a.w));
  const V3 en1 = unpack_normal(__float_as_uint(b.w)) - n0;
  const V3 en2 = unpack_normal(__float_as_uint(c.w)) - n0;

  // geometry_compute_normal, geometry_utils.cuh:13-52
  const bool is_inside = dot3(face_normal, ray) > 0.0f;
  if (is_inside)
    face_normal = neg3(face_normal);
  V3 normal = v3(n0.x + cu * en1.x + cv * en2.x, n0.y + cu * en1.y + cv * en2.y, n0.z + cu * en1.z + cv * en2.z);
  {
    const float len = len3(normal);
    normal          = (len < EPS_F) ? face_normal : normal * (1.0f / len);
  }
  float2 tex_coords = make_float2(0.0f, 0.0f);
  if (kTex) {
    tex_coords = lb_lerp_uv(tt, cu, cv);
    if (mat.normal_tex != LB_TEXTURE_NONE) {  // normal map, geometry_utils.cuh:26-49
      LbTexture t;
      const bool valid = lb_texture_valid(P.textures, P.num_textures, mat.normal_tex, t);
      V3 map_normal    = v3(0.0f, 0.0f, 1.0f);
      if (valid) {
        const float4 nf = lb_texture_fetch(t, tex_coords, true, false);
        map_normal      = v3(nf.x, nf.y, nf.z);
        if (mat.flags & DMF_NORMAL_COMPRESSED)
          map_normal = map_normal * 2.0f - v3(1.0f, 1.0f, 1.0f);
      }
      map_normal = norm3(map_normal);
      normal     = q_apply_inv(rotation_to_z(normal), map_normal);
    }
  }
  {  // normal_adaptation_apply, math.cuh:1547-1569
    const V3 Vl = neg3(ray);
    if (dot3(normal, face_normal) < 0.0f)
      normal = neg3(normal);
    if (dot3(Vl, normal) < 0.0f)
      normal = norm3(normal - (Vl * dot3(normal, Vl)) * 1.1f);
  }

  float ar = mat.ar, ag = mat.ag, ab = mat.ab, aa = mat.aa;
  if (kTex && mat.albedo_tex != LB_TEXTURE_NONE) {
    const float4 a4 = lb_texture_load(P.textures, P.num_textures, mat.albedo_tex, tex_coords, true, true, make_float4(0.9f, 0.9f, 0.9f, 1.0f));
    ar = a4.x, ag = a4.y, ab = a4.z, aa = a4.w;
  }

  const bool emissive_side    = (!is_inside) || (mat.flags & DMF_BIDIRECTIONAL);
  const bool include_emission = (mat.flags & DMF_EMISSION) && emissive_side && (state & LB_STATE_ALLOW_EMISSION);
  C3 emission                 = include_emission ? mat.emission : c3(0.0f, 0.0f, 0.0f);
  if (kTex && include_emission && mat.luminance_tex != LB_TEXTURE_NONE) {
    const float4 l4 = lb_texture_load(P.textures, P.num_textures, mat.luminance_tex, tex_coords, true, true, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    emission        = c3(l4.x, l4.y, l4.z) * (aa * mat.emission_scale);
  }

  float roughness = mat.roughness;
  if (kTex && mat.roughness_tex != LB_TEXTURE_NONE)
    roughness = lb_texture_load(P.textures, P.num_textures, mat.roughness_tex, tex_coords, true, true, make_float4(0.5f, 0.0f, 0.0f, 0.0f)).x;
  if (mat.flags & DMF_SMOOTHNESS)
    roughness = 1.0f - roughness;
  roughness = fmaxf(roughness, ROUGHNESS_CLAMP);
  if ((state & LB_STATE_DELTA_PATH) == 0)
    roughness = fmaxf(roughness, mat.roughness_clamp);

  uint32_t flags = mat.flags & DMF_TRANSLUCENT;
  // a material WITH a metallic map is not metallic: the reference leaves the map unimplemented (geometry_utils.cuh:160-162)
  if ((mat.flags & DMF_METALLIC) && !(kTex && mat.metallic_tex != LB_TEXTURE_NONE))
    flags |= MF_METALLIC;
  if (mat.flags & DMF_COLORED)
    flags |= MF_COLORED;
  if (is_inside)
    flags |= MF_INSIDE;

  const float other_ior = ior_decompress((is_inside ? (medium_ior >> 8) : medium_ior) & 0xFFu);  // medium_stack_ior_peek
  const float ior_in    = is_inside ? mat.ior : other_ior;
  const float ior_out   = is_inside ? other_ior : mat.ior;

  if ((flags & MF_TRANSLUCENT) && (fabsf(1.0f - ior_in / ior_out) < 1e-4f)) {
    if ((flags & MF_COLORED) == 0) {
      ar = 1.0f + aa * (ar - 1.0f);
      ag = 1.0f + aa * (ag - 1.0f);
      ab = 1.0f + aa * (ab - 1.0f);
    }
    aa = 0.0f;
    flags |= MF_COLORED;
  }

  Ctx ctx;
  ctx.instance_id = handle.x;
  ctx.tri_id      = handle.y;
  ctx.prim        = prim;
  ctx.normal      = transform_rotate(tr, normal);
  ctx.face_normal = pack_normal(face_normal);  // mesh space, as in the reference (geometry_utils.cuh:206)
  ctx.position    = position;
  ctx.V           = neg3(ray_world);
  ctx.state       = state;
  ctx.p.flags     = flags;
  ctx.p.albedo    = c3(quant_norm(ar, 1023), quant_norm(ag, 1023), quant_norm(ab, 1023));
  ctx.p.opacity   = quant_norm(aa, 255);
  ctx.p.roughness = quant_norm(roughness, 1023);
  ctx.p.emission  = quant_emission(emission);
  ctx.p.ior       = quant_norm((ior_in / ior_out) * (1.0f / 3.0f), 255) * 3.0f;
  return ctx;
}

// ---------------------------------------------------------------------------------------------
// the shading kernels: one instantiation per material class over that class's range of the sorted queue, one for the misses,
// and the evaluation of the BSDF-sampled light after its emitter enumeration
// ---------------------------------------------------------------------------------------------
// kAdaptive: paths of one launch carry different sample ids (adaptive sampling executions, cuda/kernels.cuh:195-356), so the
// per-launch Sobol table of lbrng::TabSampler does not apply: the sampler evaluates the Owen-scrambled Sobol pair per call.
template <bool kAdaptive>
struct ShadeSampler {
  using type = lbrng::TabSampler;
};
template <>
struct ShadeSampler<true> {
  using type = lbrng::Sampler;
};

// warp-aggregated append of one NEE segment per participating lane to shadow-queue region `slot` (one atomic per warp and slot)
__device__ __forceinline__ void push_shadow(const LbShadeParams& P, uint32_t slot, bool has, uint32_t path, V3 origin, V3 ray, float dist, C3 color,
                                            uint32_t target_prim) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t mask = __ballot_sync(0xFFFFFFFFu, has);
  if (mask == 0)
    return;
  const uint32_t leader = __ffs(mask) - 1u;
  uint32_t pos          = 0;
  if (lane == leader)
    pos = atomicAdd(&P.counters->n_shadow[slot], (uint32_t) __popc(mask));
  pos = __shfl_sync(0xFFFFFFFFu, pos, leader);
  if (has) {
    const uint32_t q  = slot * P.paths.capacity + pos + __popc(mask & ((1u << lane) - 1u));
    P.paths.sq_org[q] = make_float4(origin.x, origin.y, origin.z, __uint_as_float(path | (slot << 30)));
    P.paths.sq_dir[q] = make_float4(ray.x, ray.y, ray.z, dist);
    P.paths.sq_col[q] = make_float4(color.r, color.g, color.b, __uint_as_float(target_prim));
  }
}

#ifndef LB_SHADE_THREADS
#define LB_SHADE_THREADS 128  // block size of k_shade (64 / 128); the grid keeps the same number of threads
#endif
#ifndef LB_SHADE_MIN_BLOCKS_GENERIC
#define LB_SHADE_MIN_BLOCKS_GENERIC 4
#endif
#ifndef LB_SHADE_MIN_BLOCKS_OPAQUE
#define LB_SHADE_MIN_BLOCKS_OPAQUE 5
#endif
#define LB_SHADE_MIN_BLOCKS(kClass) ((kClass) == LB_CLASS_GENERIC ? LB_SHADE_MIN_BLOCKS_GENERIC : LB_SHADE_MIN_BLOCKS_OPAQUE)

// kCount: the instrumented variant of lumb200_device_measure_traversal (light-tree nodes descended, vertices shaded)
// sun NEE of the procedural sky: direct_lighting_sun_create_task -> direct_lighting_sun_direct (direct_lighting.cuh:21-120, 353-383),
// bsdf_sample_for_sun / bsdf_sample_for_sun_pdf (bsdf.cuh:379-458). Returns false when the task is PACKED_RECORD_BLACK.
template <int kClass>
__device__ __forceinline__ float sun_sampling_pdf(const Ctx& ctx, V3 L, float reflection_prob) {
  const Params& p = ctx.p;
  const RayCtx c  = evaluate_analyze<kClass>(p, ctx.normal, ctx.V, L);
  // QUIRK kept: the reference passes the WORLD-space V to bsdf_microfacet_pdf here (bsdf.cuh:454)
  if (c.is_refraction)
    return (1.0f - reflection_prob) * refraction_pdf(p.roughness, c.NdotH, c.NdotV, c.HdotV, c.HdotL, p.ior);
  return reflection_prob * microfacet_pdf(ctx.V, p.roughness, c.NdotH, c.NdotV);
}

template <int kClass, typename SamplerT>
__device__ bool sun_create_task(const LbShadeParams& P, const Ctx& ctx, const SamplerT& smp, V3& out_ray, C3& out_color) {
  const LbSkyDev& S = P.sky;
  const V3 sky_pos  = lbsky::world_to_sky(S, ctx.position);
  const V3 sun      = lbsky::sun_pos(S);
  const bool sun_below_horizon = lbsky::sph_ray_hit_p0(lbsky::normalize3(sun - sky_pos), sky_pos, LB_SKY_EARTH_RADIUS);
  const bool inside_earth      = lbsky::length3(sky_pos) < LB_SKY_EARTH_RADIUS;
  if (sun_below_horizon || inside_earth)
    return false;
  const Params& p = ctx.p;
  const V3 face_n = unpack_normal(ctx.face_normal);

  // direction by BSDF importance sampling
  const bool translucent      = ShadeClass<kClass>::translucent(p.flags);
  const float reflection_prob = translucent ? 0.5f : 1.0f;  // bsdf_sample_for_light_probabilities, bsdf.cuh:355-374
  V3 dir_bsdf;
  {
    const Q4 rot      = rotation_to_z(ctx.normal);
    const V3 V_local  = q_apply(rot, ctx.V);
    const float method = smp.get1(lbrng::T_LIGHT_SUN_BSDF_METHOD);
    const float2 rnd   = smp.get2(lbrng::T_LIGHT_SUN_BSDF);
    V3 ray_local;
    if (method < reflection_prob)
      ray_local = reflect3(V_local, microfacet_sample_normal(V_local, p.roughness, rnd));
    else {
      bool total_reflection;
      ray_local = refract3(V_local, refraction_sample_normal(V_local, p.roughness, rnd), p.ior, total_reflection);
    }
    dir_bsdf = norm3(q_apply_inv(rot, ray_local));
  }
  C3 light_bsdf = c3(0.0f, 0.0f, 0.0f);
  if (lbsky::sphere_ray_hit(dir_bsdf, sky_pos, sun, LB_SKY_SUN_RADIUS)) {
    const float3 sc = lbsky::sun_color(S, sky_pos, dir_bsdf);
    const RayCtx rc = evaluate_analyze<kClass>(p, ctx.normal, ctx.V, dir_bsdf);
    light_bsdf      = c3(sc.x, sc.y, sc.z) * evaluate_core<kClass>(P.luts, p, rc, H_GENERAL, dir_bsdf, face_n, 1.0f);
  }

  // direction inside the sun's solid angle
  float solid_angle;
  const V3 dir_sa = lbsky::sample_sphere(sun, LB_SKY_SUN_RADIUS, sky_pos, smp.get2(lbrng::T_LIGHT_SUN_RAY), solid_angle);
  C3 light_sa;
  {
    const float3 sc = lbsky::sun_color(S, sky_pos, dir_sa);
    const RayCtx rc = evaluate_analyze<kClass>(p, ctx.normal, ctx.V, dir_sa);
    light_sa        = c3(sc.x, sc.y, sc.z) * evaluate_core<kClass>(P.luts, p, rc, H_GENERAL, dir_sa, face_n, 1.0f);
  }

  // resampled importance sampling between the two
  const float target_bsdf = c_max(light_bsdf);
  const float target_sa   = c_max(light_sa);
  const float mis_bsdf    = solid_angle / (sun_sampling_pdf<kClass>(ctx, dir_bsdf, reflection_prob) * solid_angle + 1.0f);
  const float mis_sa      = solid_angle / (sun_sampling_pdf<kClass>(ctx, dir_sa, reflection_prob) * solid_angle + 1.0f);
  const float w_bsdf      = target_bsdf * mis_bsdf;
  const float w_sa        = target_sa * mis_sa;
  const float sum_weights = w_bsdf + w_sa;
  if (sum_weights == 0.0f)
    return false;
  float target;
  C3 color;
  if (smp.get1(lbrng::T_LIGHT_SUN_RESAMPLING) * sum_weights < w_bsdf)
    out_ray = dir_bsdf, target = target_bsdf, color = light_bsdf;
  else
    out_ray = dir_sa, target = target_sa, color = light_sa;
  color = color * (sum_weights / target);
  if (target == 0.0f || c_max(color) == 0.0f)
    return false;
  out_color = color;  // volume_integrate_transmittance: VOLUME_TYPE_NONE on this path
  return true;
}

template <int kClass, bool kTex, bool kAdaptive, bool kCount, bool kSun>
#ifdef LB_SHADE_MAXNREG
__global__ void __maxnreg__(LB_SHADE_MAXNREG) k_shade(LbShadeParams P) {  // tuning experiments: explicit register cap instead of launch bounds
#else
__global__ void __launch_bounds__(LB_SHADE_THREADS, LB_SHADE_MIN_BLOCKS(kClass)) k_shade(LbShadeParams P) {
#endif
  const uint32_t k_begin  = P.counters->class_begin[kClass];
  const uint32_t k_end    = P.counters->class_begin[kClass + 1];
  const bool sky_on       = P.frame.sky_mode != 0;  // ambient NEE, direct_lighting_ambient_is_allowed: sky.mode != DEFAULT
  const C3 sky            = (P.frame.sky_mode == 2) ? c3(P.frame.sky_r, P.frame.sky_g, P.frame.sky_b) : c3(0.0f, 0.0f, 0.0f);
  const bool has_lights   = P.num_lights > 0;
  const uint32_t lane     = threadIdx.x & 31u;
  uint32_t tree_nodes     = 0;
  // the grid is sized for the whole frame, not for this class at this depth: blocks past the end of the range leave before the staging
  if (k_begin + blockIdx.x * blockDim.x >= k_end && !(kCount && blockIdx.x == 0))
    return;

  // stage the decoded light-tree root children (<= 16 sections x 8) in shared memory: tree_prepass streams over all of them per path
#ifndef LB_STAGE_ROOT_CHILDREN
#define LB_STAGE_ROOT_CHILDREN 1  // 0: read the decoded children from global memory (A/B measurement, profiles/r2_variants.md)
#endif
#if LB_STAGE_ROOT_CHILDREN
  __shared__ float4 s_child_mean[LB_MAX_ROOT_CHILDREN];
  __shared__ float s_child_power[LB_MAX_ROOT_CHILDREN];
  if (has_lights) {
    const uint32_t num_children = min(((__ldg(P.light_root).z >> 16) & 0xFFu) * 8u, (uint32_t) LB_MAX_ROOT_CHILDREN);
    for (uint32_t c = threadIdx.x; c < num_children; c += blockDim.x) {
      s_child_mean[c]  = __ldg(P.light_root_children + 2 * c + 0);
      s_child_power[c] = __ldg(&P.light_root_children[2 * c + 1].x);
    }
    __syncthreads();
  }
  const RootChildrenShared root_children = {s_child_mean, s_child_power};
#else
  const RootChildrenGlobal root_children = {P.light_root_children};
#endif

  // warps take chunks of 32 consecutive queue entries of the class range
  for (uint32_t base = k_begin + ((blockIdx.x * blockDim.x + threadIdx.x) & ~31u); base < k_end; base += gridDim.x * blockDim.x) {
    const uint32_t k  = base + lane;
    const bool valid  = k < k_end;
    bool survives     = false;
    uint32_t i        = 0;

    // everything below is executed by the whole warp (the queue appends vote); lanes past the end of the range carry `valid = false`
    uint32_t state = 0, prim = 0, medium = 0;
    C3 rec_in      = c3(0.0f, 0.0f, 0.0f);
    V3 ray         = v3(0.0f, 0.0f, 1.0f);
    V3 hit_point   = v3(0.0f, 0.0f, 0.0f);
    typename ShadeSampler<kAdaptive>::type smp;
    smp.bluenoise = P.bluenoise;
    smp.px = smp.py = 0;
    if constexpr (kAdaptive) {
      smp.sample_id = 0;
      smp.depth     = P.rng_depth;
    }
    else
      smp.table = P.rng_table + P.rng_depth * lbrng::T_COUNT;

    Ctx ctx;
    if (valid) {
      i                    = P.queue_in[k];
      state                = P.paths.state[i];
      rec_in               = record_unpack(P.paths.record[i]);
      const float4 o4      = P.paths.org[i];
      const float4 d4      = P.paths.dir[i];
      prim                 = P.paths.prim[i];
      const uint32_t pixel = P.paths.pixel[i];
      medium               = P.paths.medium[i];
      ray                  = v3(d4.x, d4.y, d4.z);
      hit_point            = v3(o4.x, o4.y, o4.z) + ray * d4.w;
      smp.py               = pixel / P.frame.width;
      smp.px               = pixel - smp.py * P.frame.width;
      if constexpr (kAdaptive)
        smp.sample_id = P.paths.sample_id[i];
      ctx = get_context<kTex>(P, prim, hit_point, ray, state, medium);
    }

#if LB_SHADE_PREFETCH
    // software prefetch of the next chunk's path state (the queue is a gather): issued here, consumed one loop iteration later
    {
      const uint32_t k_next = k + gridDim.x * blockDim.x;
      if (k_next < k_end) {
        const uint32_t i_next = P.queue_in[k_next];
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.paths.org + i_next));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.paths.dir + i_next));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.paths.record + i_next));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.paths.state + i_next));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.paths.pixel + i_next));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.paths.medium + i_next));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.paths.prim + i_next));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.paths.result + i_next));
#if LB_SHADE_PREFETCH > 1
        const uint32_t prim_next = P.paths.prim[i_next];
        const uint2 h_next       = __ldg(P.prim_handle + prim_next);
        const uint32_t mesh_next = __ldg(P.instance_mesh + h_next.x);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.mesh_vertices[mesh_next] + 3 * (size_t) h_next.y));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.mesh_vertices[mesh_next] + 3 * (size_t) h_next.y + 2));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(P.mesh_textris[mesh_next] + h_next.y));
#endif
      }
    }
#endif

    float root_sum = 0.0f;
    if (has_lights) {
      // ---- light tree NEE: light_sample, light.cuh:85-159 ----
      uint32_t sel_prim = LB_PRIM_NONE;
      V3 sel_ray        = v3(0.0f, 0.0f, 1.0f);
      C3 cfin           = c3(0.0f, 0.0f, 0.0f);
      float sel_dist    = 0.0f;
      if (valid) {
        TreeWork work;
        tree_prepass<kClass>(P.light_root, root_children, ctx, smp, work);
        root_sum = work.root_sum;
        Reservoir res;
        res.sum_weight = 0.0f, res.selected_target = 0.0f;
        res.random         = smp.get1(lbrng::T_LIGHT_GEO_RESAMPLING);
        uint32_t sel_light = LB_LIGHT_ID_INVALID;
        C3 sel_color       = c3(0.0f, 0.0f, 0.0f);
#pragma unroll 1
        for (uint32_t out = 0; out < NUM_TREE_LANES; out++) {
          uint32_t light_id;
          float tree_weight;
          tree_postpass<kClass, kCount>(P.light_nodes, ctx, smp, out, work.cont[out], light_id, tree_weight, tree_nodes);
          if (light_id == LB_LIGHT_ID_INVALID)
            continue;
          const TriLight L = light_init(P, light_id);
          if (L.prim == prim)
            continue;  // a triangle never samples itself
          const float2 rr  = smp.get2(lbrng::T_LIGHT_GEO_RAY + out);
          V3 lray;
          float solid_angle;
          if (!light_sample_solid_angle(L, ctx.position, rr, lray, solid_angle))
            continue;
          float2 lcoords;
          const float dist = light_intersect(L, ctx.position, lray, lcoords);
          if (dist == FLT_MAX)
            continue;
          C3 lcol          = light_color_of<kTex>(P, L, lcoords);
          const RayCtx rc  = evaluate_analyze<kClass>(ctx.p, ctx.normal, ctx.V, lray);
          const C3 bw      = evaluate_core<kClass>(P.luts, ctx.p, rc, H_GENERAL, lray, unpack_normal(ctx.face_normal), 1.0f);
          const float power = c_max(lcol) * light_area(L);
          const float gi    = light_bsdf_probability<kClass>(ctx, lray);
          const float mis   = 1.0f - mis_weight_base(gi, solid_angle, power, dist * dist, root_sum);
          lcol              = (lcol * bw) * mis;
          if (reservoir_add(res, c_max(lcol), tree_weight * solid_angle)) {
            sel_light = light_id;
            sel_prim  = L.prim;
            sel_ray   = lray;
            sel_color = lcol;
            sel_dist  = dist;
          }
        }
        if (sel_light != LB_LIGHT_ID_INVALID)
          cfin = (sel_color * reservoir_weight(res)) * rec_in;
      }
      // shadow rays start at the raw hit point, the bounce at the snapped position (optix_kernel_shadow.cu:32)
      push_shadow(P, 0, c_any(cfin), i, hit_point, sel_ray, sel_dist, cfin, sel_prim);

      // ---- BSDF-sampled light: light_bsdf_get_sample (light_bsdf.cuh:24-104); the emitters along the direction are enumerated by
      //      k_trace_enum, the sample is evaluated by k_enum_finish (direct_lighting.cuh:601-669) ----
      bool has_enum = false;
      V3 bray       = v3(0.0f, 0.0f, 1.0f);
      C3 weight     = c3(0.0f, 0.0f, 0.0f);
      float prob    = 0.0f;
      if (valid) {
        const Params& p        = ctx.p;
        const float choice     = smp.get1(lbrng::T_LIGHT_BSDF_CHOICE);
        const bool translucent = ShadeClass<kClass>::translucent(p.flags);
        const uint32_t ntech   = translucent ? 2u : 1u;
        const bool use_refr    = translucent && (((uint32_t) (choice * ntech)) == 1u);
        const float refr_prob  = translucent ? 0.5f : 0.0f;
        const float rr_random  = smp.get1(lbrng::T_LIGHT_BSDF_RR);
        const float rr_prob    = lbsdf_rr_probability(p.roughness);
        if (rr_random < rr_prob) {
          const Q4 rot      = rotation_to_z(ctx.normal);
          const V3 V_local  = q_apply(rot, ctx.V);
          const V3 fn_local = q_apply(rot, unpack_normal(ctx.face_normal));
          const V3 up       = v3(0.0f, 0.0f, 1.0f);
          const float sr    = lbsdf_sampling_roughness(p.roughness);
          const float2 rnd  = smp.get2(lbrng::T_LIGHT_BSDF_DIRECTION);
          V3 r;
          if (!use_refr) {
            const V3 m      = microfacet_sample_normal(V_local, sr, rnd);
            r               = reflect3(V_local, m);
            const RayCtx rc = sample_context<kClass>(p, up, V_local, m, r, false);
            const float pdf = microfacet_pdf(V_local, sr, rc.NdotH, rc.NdotV);
            weight          = evaluate_core<kClass>(P.luts, p, rc, H_GENERAL, r, fn_local, 1.0f / pdf);
            prob            = (1.0f - refr_prob) * pdf;
          }
          else {
            bool total_reflection;
            const V3 m      = refraction_sample_normal(V_local, sr, rnd);
            r               = refract3(V_local, m, p.ior, total_reflection);
            const RayCtx rc = sample_context<kClass>(p, up, V_local, m, r, !total_reflection);
            const float pdf = refraction_pdf(sr, rc.NdotH, rc.NdotV, rc.HdotV, rc.HdotL, p.ior);
            weight          = evaluate_core<kClass>(P.luts, p, rc, H_GENERAL, r, fn_local, 1.0f / pdf);
            prob            = refr_prob * pdf;
          }
          weight = weight * (1.0f / rr_prob);
          prob *= rr_prob;
          bray     = norm3(q_apply_inv(rot, r));
          has_enum = prob != 0.0f;
        }
      }
      {
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, has_enum);
        if (mask) {
          const uint32_t leader = __ffs(mask) - 1u;
          uint32_t pos          = 0;
          if (lane == leader)
            pos = atomicAdd(&P.counters->n_enum, (uint32_t) __popc(mask));
          pos = __shfl_sync(0xFFFFFFFFu, pos, leader);
          if (has_enum) {
            const uint32_t q     = pos + __popc(mask & ((1u << lane) - 1u));
            const float trnd     = smp.get1(lbrng::T_LIGHT_BSDF_TRACE);
            P.paths.eq_org[q]    = make_float4(hit_point.x, hit_point.y, hit_point.z, __uint_as_float(i));
            P.paths.eq_dir[q]    = make_float4(bray.x, bray.y, bray.z, trnd);
            P.paths.eq_weight[q] = make_float4(weight.r, weight.g, weight.b, prob);
            P.paths.eq_rec[q]    = make_float4(rec_in.r, rec_in.g, rec_in.b, root_sum);
          }
        }
      }
    }

    // ---- sun NEE (procedural sky): one unbounded shadow ray towards the sun disc, evaluated like the reference's
    //      direct_lighting_sun_evaluate_task (direct_lighting.cuh:465-529, no ocean): task colour x throughput x transmittance ----
    if constexpr (kSun) {
      bool has_sun = false;
      V3 sray      = v3(0.0f, 0.0f, 1.0f);
      C3 scol      = c3(0.0f, 0.0f, 0.0f);
      if (valid) {
        V3 dir;
        C3 color;
        if (sun_create_task<kClass>(P, ctx, smp, dir, color)) {
          const uint2 pc = record_pack(color);  // the task travels as record_pack / ray_pack (DeviceTaskDirectLightSun)
          if (pc.x != 0 || pc.y != 0) {
            sray    = ray_unpack(ray_pack(dir));
            scol    = record_unpack(pc) * rec_in;
            has_sun = c_any(scol);
          }
        }
      }
      push_shadow(P, LB_NEE_SLOT_SUN, has_sun, i, hit_point, sray, FLT_MAX, scol, LB_PRIM_NONE);
    }

    // ---- bounce ----
    Bounce bounce;
    bounce.ray = v3(0.0f, 0.0f, 1.0f), bounce.weight = c3(0.0f, 0.0f, 0.0f), bounce.transparent_pass = false, bounce.microfacet_based = false;
    if (valid)
      bounce = bsdf_sample<kClass>(P.luts, ctx, smp);

    // ---- ambient NEE along the bounce direction (direct_lighting.cuh:382-401, 531-599) ----
    if (sky_on) {
      bool has_amb = false;
      V3 aray      = v3(0.0f, 0.0f, 1.0f);
      C3 col       = c3(0.0f, 0.0f, 0.0f);
      if (valid) {
        // sky_color_no_compute(position, ray, state = 0), sky.cuh:534-565: the constant colour, or the HDRI table without the sun's disc
        C3 amb_sky = sky;
        if constexpr (kSun) {
          if (P.sky.mode == 1) {
            const float3 h = lbsky::sky_color_hdri(P.sky, ctx.position, bounce.ray, false);
            amb_sky        = c3(h.x, h.y, h.z);
          }
        }
        const uint2 pc = record_pack(amb_sky * bounce.weight);
        if (pc.x != 0 || pc.y != 0) {
          aray    = ray_unpack(ray_pack(bounce.ray));
          col     = record_unpack(pc) * rec_in;
          has_amb = c_any(col);
        }
      }
      push_shadow(P, 2, has_amb, i, hit_point, aray, FLT_MAX, col, LB_PRIM_NONE);
    }

    if (valid) {
      // ---- delta / pass-through bookkeeping (geometry.cuh:80-97) ----
      bool is_delta;
      if (bounce.transparent_pass) {
        const float scale = (ctx.p.ior >= 1.0f) ? ctx.p.ior : 1.0f / ctx.p.ior;
        is_delta          = ctx.p.roughness * fminf(scale - 1.0f, 1.0f) <= DELTA_PATH_CUTOFF;
      }
      else
        is_delta = bounce.microfacet_based && (ctx.p.roughness <= DELTA_PATH_CUTOFF);
      const bool pass_through = bounce.transparent_pass && ((ctx.p.ior == 1.0f) || !bounce.microfacet_based);

      // ---- emission ----
      if (c_any(ctx.p.emission)) {
        const C3 e = ctx.p.emission * rec_in;
        float4 res = P.paths.result[i];
        res.x += e.r, res.y += e.g, res.z += e.b;
        P.paths.result[i] = res;
      }

      C3 rec = rec_in * bounce.weight;

      uint32_t new_state = state | LB_STATE_USE_IGNORE_HANDLE;
      if (sky_on && !pass_through)
        new_state &= ~LB_STATE_ALLOW_AMBIENT;
      else
        new_state |= LB_STATE_ALLOW_AMBIENT;
      if (!is_delta)
        new_state &= ~LB_STATE_DELTA_PATH;
      if (!pass_through)
        new_state &= ~(LB_STATE_CAMERA_DIRECTION | LB_STATE_ALLOW_EMISSION);

      // ---- Russian roulette on the incoming state (directives.cuh:11-32) ----
      survives = true;
      if (!(state & LB_STATE_DELTA_PATH)) {
        const float value = c_max(rec);
        if (value < P.camera.rr_threshold) {
          const float pr = (value > 0.0f) ? fmaxf(value / P.camera.rr_threshold, RR_CLAMP) : 0.0f;
          if (smp.get1(lbrng::T_RUSSIAN_ROULETTE) > pr)
            survives = false;
          else
            rec = rec * (1.0f / pr);
        }
      }
      if (P.is_last)
        survives = false;  // the reference still writes the bounce task of the last iteration, nothing consumes it

      if (survives && bounce.transparent_pass) {  // medium transition, geometry.cuh:160-175
        if (!(ctx.p.flags & MF_INSIDE))
          medium = (medium << 8) | ior_compress(ior_decompress(medium & 0xFFu) / ctx.p.ior);
        else
          medium >>= 8;
      }

      if (survives) {
        P.paths.org[i]    = make_float4(ctx.position.x, ctx.position.y, ctx.position.z, 0.0f);
        P.paths.dir[i]    = make_float4(bounce.ray.x, bounce.ray.y, bounce.ray.z, FLT_MAX);
        P.paths.record[i] = record_pack(rec);
        P.paths.state[i]  = new_state;
        P.paths.medium[i] = medium;
      }
    }

    // warp-aggregated append of the survivors to the next queue
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, survives);
    if (mask) {
      uint32_t pos = 0;
      if (lane == (uint32_t) (__ffs(mask) - 1))
        pos = atomicAdd(&P.counters->n_next, (uint32_t) __popc(mask));
      pos = __shfl_sync(0xFFFFFFFFu, pos, __ffs(mask) - 1);
      if (survives)
        P.queue_out[pos + __popc(mask & ((1u << lane) - 1u))] = i;
    }
  }
  if (kCount) {
    for (int o = 16; o > 0; o >>= 1)
      tree_nodes += __shfl_xor_sync(0xFFFFFFFFu, tree_nodes, o);
    if (lane == 0 && tree_nodes)
      atomicAdd(&P.counters->light_tree_nodes, (unsigned long long) tree_nodes);
    if (blockIdx.x == 0 && threadIdx.x == 0)
      atomicAdd(&P.counters->shaded_vertices, (unsigned long long) (k_end - k_begin));
  }
}

// sky_process_tasks, sky.cuh:609-633: the misses are the tail [n_hits, n_active) of the sorted queue
__global__ void __launch_bounds__(256) k_shade_miss(LbShadeParams P) {
  const uint32_t n_active = P.counters->n_active;
  const uint32_t n_hits   = P.counters->n_hits;
  if (P.frame.sky_mode != 2)
    return;
  const C3 sky = c3(P.frame.sky_r, P.frame.sky_g, P.frame.sky_b);
  for (uint32_t k = n_hits + blockIdx.x * blockDim.x + threadIdx.x; k < n_active; k += gridDim.x * blockDim.x) {
    const uint32_t i = P.queue_in[k];
    if (P.paths.state[i] & LB_STATE_ALLOW_AMBIENT) {
      const C3 s = sky * record_unpack(P.paths.record[i]);
      if (c_any(s)) {
        float4 res = P.paths.result[i];
        res.x += s.r, res.y += s.g, res.z += s.b;
        P.paths.result[i] = res;
      }
    }
  }
}

// Debug shading modes (LuminaryShadingMode != DEFAULT): geometry_process_tasks_debug (cuda/geometry.cuh:182-246) over the hits
// [0, n_hits) of the sorted queue and the IDENTIFICATION colour of sky_process_tasks_debug (cuda/sky.cuh:635-668) over the misses.
// ALBEDO misses are the sky itself, sky_color_main with the camera state: the regular miss kernels shade them (record = 1).
template <bool kTex>
__global__ void __launch_bounds__(128) k_shade_debug(LbShadeParams P, uint32_t mode) {
  const uint32_t n_active = P.counters->n_active;
  const uint32_t n_hits   = P.counters->n_hits;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_active; k += gridDim.x * blockDim.x) {
    const uint32_t i = P.queue_in[k];
    C3 result        = c3(0.0f, 0.0f, 0.0f);
    if (k < n_hits) {
      const float4 o4     = P.paths.org[i];
      const float4 d4     = P.paths.dir[i];
      const uint32_t prim = P.paths.prim[i];
      const V3 ray        = v3(d4.x, d4.y, d4.z);
      const V3 hit_point  = v3(o4.x, o4.y, o4.z) + ray * d4.w;  // task.origin + task.ray * trace.depth
      switch (mode) {
        case 1: {  // LUMINARY_SHADING_MODE_ALBEDO
          const Ctx ctx = get_context<kTex>(P, prim, hit_point, ray, P.paths.state[i], P.paths.medium[i]);
          result        = ctx.p.albedo + ctx.p.emission;
        } break;
        case 2: {  // DEPTH
          const float v = __saturatef((1.0f / d4.w) * 2.0f);
          result        = c3(v, v, v);
        } break;
        case 3: {  // NORMAL
          const Ctx ctx = get_context<kTex>(P, prim, hit_point, ray, P.paths.state[i], P.paths.medium[i]);
          result        = c3(__saturatef(ctx.normal.x), __saturatef(ctx.normal.y), __saturatef(ctx.normal.z));
        } break;
        case 4: {  // IDENTIFICATION
          const uint2 handle = __ldg(P.prim_handle + prim);
          const uint32_t v   = lbrng::squares32(0x55555555u, (handle.x << 16) | handle.y);
          result = c3((float) (v & 0x7ffu) / 0x7ff, (float) ((v >> 10) & 0x7ffu) / 0x7ff, (float) ((v >> 20) & 0x7ffu) / 0x7ff);
        } break;
        case 5: {  // LIGHTS
          const Ctx ctx = get_context<kTex>(P, prim, hit_point, ray, P.paths.state[i], P.paths.medium[i]);
          result        = ctx.p.albedo * 0.025f + ctx.p.emission;
        } break;
        default:
          break;
      }
    }
    else if (mode == 4)
      result = c3(0.0f, 0.63f, 1.0f);
    if (c_any(result)) {  // write_beauty_buffer, memory.cuh:359-368
      float4 res = P.paths.result[i];
      res.x += result.r, res.y += result.g, res.z += result.b;
      P.paths.result[i] = res;
    }
  }
}

int lb_launch_shade_debug(const LbShadeParams& sp, uint32_t mode, int grid, cudaStream_t s) {
  int launches = 1;
  if (sp.textured)
    k_shade_debug<true><<<grid, 128, 0, s>>>(sp, mode);
  else
    k_shade_debug<false><<<grid, 128, 0, s>>>(sp, mode);
  if (mode == 1) {
    if (sp.frame.sky_mode == 2)
      k_shade_miss<<<max(grid / 8, 1), 256, 0, s>>>(sp);
    else
      lb_launch_shade_miss_sky(sp, grid, s);
    launches++;
  }
  return launches;
}

// direct_lighting_bsdf_evaluate_task after the any-hit enumeration (direct_lighting.cuh:601-669): re-intersect the selected emitter
// (light_triangle_intersection_uv), MIS against the light tree's pdf (mis_compute_weight_gi, mis.cuh:26-39), scale by the number of
// emitters the reservoir saw, and queue the transmittance test as a slot-1 shadow segment.
template <bool kTex>
__global__ void __launch_bounds__(128) k_enum_finish(LbShadeParams P) {
  const uint32_t n = P.counters->n_enum;
  for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
    const uint32_t q = base + (threadIdx.x & 31u);
    bool has         = false;
    uint32_t path = 0, target = LB_PRIM_NONE;
    V3 origin = v3(0.0f, 0.0f, 0.0f), bray = v3(0.0f, 0.0f, 1.0f);
    C3 lcol   = c3(0.0f, 0.0f, 0.0f);
    float dist = 0.0f;
    if (q < n) {
      const float4 o       = P.paths.eq_org[q];
      const float4 d       = P.paths.eq_dir[q];
      const uint32_t light = __float_as_uint(d.w);
      path                 = __float_as_uint(o.w);
      origin               = v3(o.x, o.y, o.z);
      bray                 = v3(d.x, d.y, d.z);
      if (light != LB_LIGHT_ID_INVALID) {
        const float4 w4          = P.paths.eq_weight[q];
        const float4 r4          = P.paths.eq_rec[q];
        const uint32_t num_hits  = P.paths.eq_hits[q];
        const TriLight L         = light_init(P, light);
        float2 lcoords;
        dist = light_intersect(L, origin, bray, lcoords);
        if (dist != FLT_MAX) {
          lcol      = light_color_of<kTex>(P, L, lcoords);
          float mis = 1.0f;
          if (r4.w != 0.0f)
            mis = mis_weight_base(w4.w, light_solid_angle(L, origin), c_max(lcol) * light_area(L), dist * dist, r4.w);
          lcol   = ((lcol * (mis * num_hits)) * c3(w4.x, w4.y, w4.z)) * c3(r4.x, r4.y, r4.z);
          target = L.prim;
          has    = c_any(lcol);
        }
      }
    }
    push_shadow(P, 1, has, path, origin, bray, dist, lcol, target);
  }
}

// accumulation_collect_results (accumulation.cuh:36-84): one path per pixel and pass, so no atomics are needed
// A sample whose radiance is not finite is dropped (and counted, Lumb200Stats.nonfinite_samples): the reference keeps a sanitising
// store for exactly this (store_RGBF, cuda/memory.cuh:343-350: non-finite -> black) - one such sample in 10^9 would otherwise turn
// its pixel NaN for the rest of the render.
__global__ void __launch_bounds__(256) k_accumulate(LbPaths paths, uint32_t n, float* __restrict__ planes, LbCounters* __restrict__ counters) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 r           = paths.result[i];
    const uint32_t pix = paths.pixel[i];
#pragma unroll
    for (int s = 0; s < LB_NEE_SLOTS; s++) {
      const float4 a = paths.nee[LB_NEE_SLOTS * (size_t) i + s];
      r.x += a.x, r.y += a.y, r.z += a.z;
    }
    if (non_finite(r.x + r.y + r.z)) {
      atomicAdd(&counters->nonfinite_samples, 1u);
      counters->nonfinite_pixel = pix;
      continue;
    }
    planes[0 * (size_t) n + pix] += r.x;
    planes[1 * (size_t) n + pix] += r.y;
    planes[2 * (size_t) n + pix] += r.z;
    planes[3 * (size_t) n + pix] += c_lum(c3(r.x * r.x, r.y * r.y, r.z * r.z));
  }
}

// accumulation_generate_result, beauty mode without local error minimisation (accumulation.cuh:86-190)
__global__ void __launch_bounds__(256) k_generate_result(const float* __restrict__ planes, float* __restrict__ result, uint32_t n,
                                                         float normalization) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    result[0 * (size_t) n + i] = planes[0 * (size_t) n + i] * normalization;
    result[1 * (size_t) n + i] = planes[1 * (size_t) n + i] * normalization;
    result[2 * (size_t) n + i] = planes[2 * (size_t) n + i] * normalization;
  }
}

// ---------------------------------------------------------------------------------------------
// output chain: generate_final_image + convert_RGBF_to_ARGB8 (cuda/kernels.cuh:503-644) with tonemap_apply
// (cuda/tonemap.cuh:7-246) for supersampling 0 / undersampling 0 / filter NONE: mean -> exposure -> tone map ->
// sRGB transfer -> blue-noise dither -> LuminaryARGB8 {b, g, r, a}. Purkinje shift, colour correction and film
// grain are not implemented (the host layer rejects them).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float srgb_to_linear(float v) { return (v <= 0.04045f) ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f); }
__device__ __forceinline__ float linear_to_srgb(float v) { return (v <= 0.0031308f) ? 12.92f * v : 1.055f * powf(v, 0.416666666667f) - 0.055f; }

__device__ C3 tm_aces(C3 p) {
  C3 c = c3(0.59719f * p.r + 0.35458f * p.g + 0.04823f * p.b, 0.07600f * p.r + 0.90834f * p.g + 0.01566f * p.b,
            0.02840f * p.r + 0.13383f * p.g + 0.83777f * p.b);
  C3 a = c * (c + c3(0.0245786f, 0.0245786f, 0.0245786f)) + c3(-0.000090537f, -0.000090537f, -0.000090537f);
  C3 b = c * (c * 0.983729f + c3(0.432951f, 0.432951f, 0.432951f)) + c3(0.238081f, 0.238081f, 0.238081f);
  c    = c3(a.r / b.r, a.g / b.g, a.b / b.b);
  return c3(1.60475f * c.r - 0.53108f * c.g - 0.07367f * c.b, -0.10208f * c.r + 1.10813f * c.g - 0.00605f * c.b,
            -0.00327f * c.r - 0.07276f * c.g + 1.07602f * c.b);
}
__device__ __forceinline__ float tm_u2(float x) {
  const float a = 0.15f, b = 0.50f, c = 0.10f, d = 0.20f, e = 0.02f, f = 0.30f;
  return ((x * (a * x + c * b) + d * e) / (x * (a * x + b) + d * f)) - e / f;
}
__device__ __forceinline__ float agx_poly(float v) {
  const float v2 = v * v, v4 = v2 * v2;
  return 15.5f * v4 * v2 - 40.14f * v4 * v + 31.96f * v4 - 6.868f * v2 * v + 0.4298f * v2 + 0.1191f * v - 0.00232f;
}
__device__ C3 agx_forward(C3 p) {
  C3 a = c3(0.842479062253094f, 0.0423282422610123f, 0.0423756549057051f) * p.r + c3(0.0784335999999992f, 0.878468636469772f, 0.0784336f) * p.g +
         c3(0.0792237451477643f, 0.0791661274605434f, 0.879142973793104f) * p.b;
  const float lo = -12.47393f, hi = 4.026069f;
  a.r = (fminf(fmaxf(log2f(fmaxf(a.r, 0.00017578139f)), lo), hi) - lo) / (hi - lo);
  a.g = (fminf(fmaxf(log2f(fmaxf(a.g, 0.00017578139f)), lo), hi) - lo) / (hi - lo);
  a.b = (fminf(fmaxf(log2f(fmaxf(a.b, 0.00017578139f)), lo), hi) - lo) / (hi - lo);
  return c3(agx_poly(a.r), agx_poly(a.g), agx_poly(a.b));
}
__device__ C3 agx_inverse(C3 p) {
  C3 a = c3(1.19687900512017f, -0.0528968517574562f, -0.0529716355144438f) * p.r + c3(-0.0980208811401368f, 1.15190312990417f, -0.0980434501171241f) * p.g +
         c3(-0.0990297440797205f, -0.0989611768448433f, 1.15107367264116f) * p.b;
  return c3(srgb_to_linear(fmaxf(a.r, 0.0f)), srgb_to_linear(fmaxf(a.g, 0.0f)), srgb_to_linear(fmaxf(a.b, 0.0f)));
}
__device__ C3 agx_look(C3 p, float slope, float power, float saturation) {
  const float lum = c_lum(p);
  p               = p * slope;
  p               = c3(powf(p.r, power), powf(p.g, power), powf(p.b, power));
  return c3(lum + saturation * (p.r - lum), lum + saturation * (p.g - lum), lum + saturation * (p.b - lum));
}

// purkinje_shift, cuda/purkinje.cuh:19-90 (Kirk & O'Brien 2011; constants of the reference)
__device__ C3 purkinje_shift(C3 pixel, float kappa1, float kappa2) {
  const float strength = 5000.0f;
  if (c_lum(pixel) >= (1.0f / strength))
    return pixel;
  const float long_cone   = 0.096869562190332f * pixel.r + 0.318940374720484f * pixel.g - 0.188428411786113f * pixel.b;
  const float medium_cone = 0.020208210904239f * pixel.r + 0.291385283197581f * pixel.g - 0.090918262127325f * pixel.b;
  const float short_cone  = 0.002760510899553f * pixel.r - 0.008341563564118f * pixel.g + 0.067213551661950f * pixel.b;
  const float rod         = -0.007607045462440f * pixel.r + 0.122492925567539f * pixel.g + 0.022445835141881f * pixel.b;
  const float lm = 1.0f / 0.63721f, mm = 1.0f / 0.39242f, sm = 1.0f / 1.6064f;
  const float sr = rsqrtf(fmaxf(1.0f + (1.0f / 3.0f) * lm * (long_cone + kappa1 * rod), EPS_F));
  const float sg = rsqrtf(fmaxf(1.0f + (1.0f / 3.0f) * mm * (medium_cone + kappa1 * rod), EPS_F));
  const float sb = rsqrtf(fmaxf(1.0f + (1.0f / 3.0f) * sm * (short_cone + kappa2 * rod), EPS_F));
  const float K = 45.0f, S = 10.0f, k3 = 0.6f, rw = 0.139f, p = 0.6189f;
  C3 opp = c3(((-k3 - rw) * sr + (1.0f + k3 * rw) * sg) * kappa1 * lm, (p * k3 * sr + (1.0f - p) * k3 * sg + sb) * kappa1 * mm,
              (p * S * sr + (1.0f - p) * S * sg) * kappa2 * sm);
  opp    = opp * ((K / S) * rod);
  const C3 lms = c3(long_cone + 0.5f * (opp.b - opp.r), medium_cone + 0.5f * (opp.b + opp.r), short_cone + opp.g + opp.b);
  const C3 xyz = c3(1.9102f * lms.r - 1.1121f * lms.g + 0.2019f * lms.b, 0.3710f * lms.r + 0.6291f * lms.g + 0.0000f * lms.b,
                    0.0000f * lms.r + 0.0000f * lms.g + 1.0000f * lms.b);
  const C3 rgb = c3(3.2405f * xyz.r - 1.5371f * xyz.g - 0.4985f * xyz.b, -0.9693f * xyz.r + 1.876f * xyz.g + 0.0416f * xyz.b,
                    0.0556f * xyz.r - 0.2040f * xyz.g + 1.0572f * xyz.b);
  float blend  = __saturatef(1.0f - strength * c_lum(pixel));
  blend *= blend;
  return pixel * (1.0f - blend) + rgb * blend;
}

__device__ C3 rgb_to_hsv(C3 rgb) {  // math.cuh:1483-1511
  const float mx = fmaxf(rgb.r, fmaxf(rgb.g, rgb.b)), mn = fminf(rgb.r, fminf(rgb.g, rgb.b));
  const float s  = (mx - mn) / mx;
  float h        = 0.0f;
  if (s != 0.0f) {
    const float delta = mx - mn;
    if (mx == rgb.r)
      h = (rgb.g - rgb.b) / delta;
    else if (mx == rgb.g)
      h = 2.0f + (rgb.b - rgb.r) / delta;
    else
      h = 4.0f + (rgb.r - rgb.g) / delta;
    h *= 1.0f / 6.0f;
    if (h < 0.0f)
      h += 1.0f;
  }
  return c3(h, s, mx);
}
__device__ C3 hsv_to_rgb(C3 hsv) {  // math.cuh:1516-1541
  const float s = hsv.g, v = hsv.b;
  if (s == 0.0f)
    return c3(v, v, v);
  const float h = hsv.r * 6.0f;
  C3 hue        = c3(fmodf(h, 6.0f), fmodf(h + 4.0f, 6.0f), fmodf(h + 2.0f, 6.0f));
  hue           = c3(__saturatef(fabsf(hue.r - 3.0f) - 1.0f), __saturatef(fabsf(hue.g - 3.0f) - 1.0f), __saturatef(fabsf(hue.b - 3.0f) - 1.0f));
  return (c3(1.0f - s, 1.0f - s, 1.0f - s) + hue * s) * v;
}
// random_uint16_t (Squares, 16-bit output, random.cuh:197-212,297-299) as a float in [0, 1): the film-grain mask
__device__ float white_noise_offset(uint32_t offset) {
  const uint32_t key = 0xfcbd6e15u;
  uint32_t x = offset * key;
  const uint32_t y = x, z = y + key;
  x = x * x + y;
  x = lbrng::swap16(x);
  x = x * x + z;
  x = lbrng::swap16(x);
  const uint32_t v = (x * x + y) >> 16;
  return __uint_as_float(0x3F800000u | (v << 7)) - 1.0f;
}

__device__ C3 tonemap_transform(C3 p, const Lumb200OutputParams& op);

// (x, y): INTERNAL pixel, width = internal width (random_grain_mask, random.cuh:377-379)
__device__ C3 tonemap_pixel(C3 p, const Lumb200OutputParams& op, uint32_t x, uint32_t y, uint32_t width) {  // tonemap_apply, cuda/tonemap.cuh:205-246
  if (op.purkinje)
    p = purkinje_shift(p, op.purkinje_kappa1, op.purkinje_kappa2);
  if (op.use_color_correction) {
    C3 hsv = rgb_to_hsv(p) + c3(op.color_correction[0], op.color_correction[1], op.color_correction[2]);
    if (hsv.r < 0.0f)
      hsv.r += 1.0f;
    if (hsv.r > 1.0f)
      hsv.r -= 1.0f;
    hsv.g = __saturatef(hsv.g);
    if (hsv.b < 0.0f)
      hsv.b = 0.0f;
    p = hsv_to_rgb(hsv);
  }
  p = p * op.exposure;
  const float grain = op.film_grain * (white_noise_offset(x + y * width) - 0.5f);
  p = c3(fmaxf(p.r + grain, 0.0f), fmaxf(p.g + grain, 0.0f), fmaxf(p.b + grain, 0.0f));
  return tonemap_transform(p, op);
}

// exposure + tonemap_apply_transform only (the adaptive sampler's compression factor, adaptive_sampling.cuh:9-17)
__device__ C3 tonemap_pixel(C3 p, const Lumb200OutputParams& op) {
  if (op.purkinje)
    p = purkinje_shift(p, op.purkinje_kappa1, op.purkinje_kappa2);
  p = p * op.exposure;
  p = c3(fmaxf(p.r, 0.0f), fmaxf(p.g, 0.0f), fmaxf(p.b, 0.0f));
  return tonemap_transform(p, op);
}

__device__ C3 tonemap_transform(C3 p, const Lumb200OutputParams& op) {  // tonemap_apply_transform, cuda/tonemap.cuh:175-203
  switch (op.tonemap) {
    case 1: p = tm_aces(p); break;
    case 2: p = p * (1.0f / (1.0f + c_lum(p))); break;
    case 3: {
      const float s = 1.0f / tm_u2(11.2f);
      p             = c3(tm_u2(2.0f * p.r) * s, tm_u2(2.0f * p.g) * s, tm_u2(2.0f * p.b) * s);
    } break;
    case 4: p = agx_inverse(agx_forward(p)); break;
    case 5: p = agx_inverse(agx_look(agx_forward(p), 1.0f, 1.35f, 1.4f)); break;
    case 6: p = agx_inverse(agx_look(agx_forward(p), op.agx_slope, op.agx_power, op.agx_saturation)); break;
    default: break;
  }
  return p;
}

// One thread per OUTPUT pixel; `width` / `height` are the internal (rendered) resolution.
// raw: tonemap_apply returns the pixel untouched under a debug shading mode (tonemap.cuh:207-208); filter and dithering still apply.
__global__ void __launch_bounds__(256) k_output_argb8(const float* __restrict__ planes, uint32_t width, uint32_t height, float normalization,
                                                      Lumb200OutputParams op, const uint16_t* __restrict__ bluenoise_1d, uchar4* __restrict__ dst,
                                                      bool raw) {
  const uint32_t n     = width * height;
  const uint32_t scale = 1u << op.supersampling;
  const uint32_t ow = width >> op.supersampling, oh = height >> op.supersampling;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ow * oh; i += gridDim.x * blockDim.x) {
    const uint32_t y = i / ow, x = i - y * ow;
    C3 p = c3(0.0f, 0.0f, 0.0f);
    for (uint32_t yi = 0; yi < scale; yi++)
      for (uint32_t xi = 0; xi < scale; xi++) {
        const uint32_t px = min(x * scale + xi, width - 1), py = min(y * scale + yi, height - 1);
        const size_t k    = px + (size_t) py * width;
        const C3 mean     = c3(planes[k], planes[(size_t) n + k], planes[2 * (size_t) n + k]) * normalization;
        p                 = p + (raw ? mean : tonemap_pixel(mean, op, px, py, width));
      }
    p = p * (1.0f / (scale * scale));
    // random_dither_mask (random.cuh:370-375): the filters threshold against it whether or not the output is dithered
    const float mask = bluenoise_1d ? __uint_as_float(0x3F800000u | ((uint32_t) bluenoise_1d[(x & 0xFFu) + (y & 0xFFu) * 256u] << 7)) - 1.0f : 0.5f;
    switch (op.filter) {  // convert_RGBF_to_ARGB8, kernels.cuh:615-637; math.cuh:1081-1168
      case 1: {
        const float v = c_lum(p);
        p             = c3(v, v, v);
      } break;
      case 2:
        p = c3(p.r * 0.393f + p.g * 0.769f + p.b * 0.189f, p.r * 0.349f + p.g * 0.686f + p.b * 0.168f, p.r * 0.272f + p.g * 0.534f + p.b * 0.131f);
        break;
      case 3: {
        const int tone = (int) (4.0f * c_lum(p) + mask);
        p = (tone == 0)   ? c3(15.0f / 255.0f, 56.0f / 255.0f, 15.0f / 255.0f)
            : (tone == 1) ? c3(48.0f / 255.0f, 98.0f / 255.0f, 48.0f / 255.0f)
            : (tone == 2) ? c3(139.0f / 255.0f, 172.0f / 255.0f, 15.0f / 255.0f)
                          : c3(155.0f / 255.0f, 188.0f / 255.0f, 15.0f / 255.0f);
      } break;
      case 4: {
        const int tone = (int) (4.0f * c_lum(p) + mask);
        const float v  = (tone == 0) ? 0.0f : (tone == 1) ? 1.0f / 3.0f : (tone == 2) ? 2.0f / 3.0f : 1.0f;
        p              = c3(v, v, v);
      } break;
      case 5: {
        p = p * 1.5f;
        const uint32_t row = y % 3u;
        p = (row == 0) ? c3(0.0f, 0.0f, p.b) : (row == 1) ? c3(p.r, 0.0f, 0.0f) : c3(0.0f, p.g, 0.0f);
      } break;
      case 6: {
        const int tone = (int) (2.0f * c_lum(p) + mask);
        const float v  = (tone == 0) ? 0.0f : 1.0f;
        p              = c3(v, v, v);
      } break;
      default: break;
    }
    const float dither = op.dithering ? mask : 0.5f;
    const float r = fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * linear_to_srgb(p.r)));
    const float g = fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * linear_to_srgb(p.g)));
    const float b = fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * linear_to_srgb(p.b)));
    dst[i]        = make_uchar4((uint8_t) b, (uint8_t) g, (uint8_t) r, 0xFFu);
  }
}

// ---------------------------------------------------------------------------------------------
// bloom: post_image_downsample / post_image_upsample (cuda/post_common.cuh:8-143) driven by _device_post_bloom_apply
// (device/device_post.c:62-140) on the mean-radiance planes, before the tone map. One fp32 plane at a time.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float post_sample(const float* __restrict__ buffer, float x, float y, uint32_t width, uint32_t height) {
  const float source_x = fmaxf(0.0f, x * (width - 1));
  const float source_y = fmaxf(0.0f, y * (height - 1));
  const uint32_t x0 = (uint32_t) source_x, y0 = (uint32_t) source_y;
  const uint32_t x1 = min((uint32_t) (source_x + 1.0f), width - 1), y1 = min((uint32_t) (source_y + 1.0f), height - 1);
  const float p00 = buffer[x0 + y0 * width], p01 = buffer[x0 + y1 * width], p10 = buffer[x1 + y0 * width], p11 = buffer[x1 + y1 * width];
  const float fx = source_x - x0, ifx = 1.0f - fx, fy = source_y - y0, ify = 1.0f - fy;
  float result = p00 * (ifx * ify);
  result       = __fmaf_rn(p01, ifx * fy, result);
  result       = __fmaf_rn(p10, fx * ify, result);
  result       = __fmaf_rn(p11, fx * fy, result);
  return result;
}

__device__ __forceinline__ float post_sample_border(const float* __restrict__ image, float x, float y, uint32_t width, uint32_t height) {
  const float below_one = __uint_as_float(0x3F7FFFFFu);
  if (x > below_one || x < 0.0f || y > below_one || y < 0.0f)
    return 0.0f;
  return post_sample(image, x, y, width, height);
}

__global__ void __launch_bounds__(256) k_post_downsample(const float* __restrict__ src, uint32_t sw, uint32_t sh, float* __restrict__ dst,
                                                         uint32_t tw, uint32_t th) {
  const float scale_x = 1.0f / (tw - 1), scale_y = 1.0f / (th - 1);
  const float step_x = 1.0f / (sw - 1), step_y = 1.0f / (sh - 1);
  for (uint32_t index = blockIdx.x * blockDim.x + threadIdx.x; index < tw * th; index += gridDim.x * blockDim.x) {
    const uint32_t y = index / tw, x = index - y * tw;
    const float sx = scale_x * x, sy = scale_y * y;
    const float hx = 0.5f * step_x, hy = 0.5f * step_y;
    float pixel = 0.0f;
    pixel += post_sample_border(src, sx - hx, sy - hy, sw, sh);
    pixel += post_sample_border(src, sx + hx, sy - hy, sw, sh);
    pixel += post_sample_border(src, sx - hx, sy + hy, sw, sh);
    pixel += post_sample_border(src, sx + hx, sy + hy, sw, sh);
    pixel += post_sample_border(src, sx, sy, sw, sh);
    pixel = __fmaf_rn(post_sample_border(src, sx, sy - step_y, sw, sh), 0.5f, pixel);
    pixel = __fmaf_rn(post_sample_border(src, sx - step_x, sy, sw, sh), 0.5f, pixel);
    pixel = __fmaf_rn(post_sample_border(src, sx + step_x, sy, sw, sh), 0.5f, pixel);
    pixel = __fmaf_rn(post_sample_border(src, sx, sy + step_y, sw, sh), 0.5f, pixel);
    pixel = __fmaf_rn(post_sample_border(src, sx - step_x, sy - step_y, sw, sh), 0.25f, pixel);
    pixel = __fmaf_rn(post_sample_border(src, sx + step_x, sy - step_y, sw, sh), 0.25f, pixel);
    pixel = __fmaf_rn(post_sample_border(src, sx - step_x, sy + step_y, sw, sh), 0.25f, pixel);
    pixel = __fmaf_rn(post_sample_border(src, sx + step_x, sy + step_y, sw, sh), 0.25f, pixel);
    pixel *= 1.0f / 8.0f;
    dst[x + y * tw] = fmaxf(pixel, 0.0f);  // threshold 0
  }
}

// dst may alias base (the reference upsamples in place, device_post.c:106-135); src never does
__global__ void __launch_bounds__(256) k_post_upsample(const float* __restrict__ src, uint32_t sw, uint32_t sh, float* dst, const float* base,
                                                       uint32_t tw, uint32_t th, float sa, float sb) {
  const float scale_x = 1.0f / (tw - 1), scale_y = 1.0f / (th - 1);
  const float step_x = 1.0f / (sw - 1), step_y = 1.0f / (sh - 1);
  for (uint32_t index = blockIdx.x * blockDim.x + threadIdx.x; index < tw * th; index += gridDim.x * blockDim.x) {
    const uint32_t y = index / tw, x = index - y * tw;
    const float sx = scale_x * x, sy = scale_y * y;
    float pixel = post_sample_border(src, sx - step_x, sy - step_y, sw, sh);
    pixel       = __fmaf_rn(post_sample_border(src, sx, sy - step_y, sw, sh), 2.0f, pixel);
    pixel += post_sample_border(src, sx + step_x, sy - step_y, sw, sh);
    pixel = __fmaf_rn(post_sample_border(src, sx - step_x, sy, sw, sh), 2.0f, pixel);
    pixel = __fmaf_rn(post_sample_border(src, sx, sy, sw, sh), 4.0f, pixel);
    pixel = __fmaf_rn(post_sample_border(src, sx + step_x, sy, sw, sh), 2.0f, pixel);
    pixel += post_sample_border(src, sx - step_x, sy + step_y, sw, sh);
    pixel = __fmaf_rn(post_sample_border(src, sx, sy + step_y, sw, sh), 2.0f, pixel);
    pixel += post_sample_border(src, sx + step_x, sy + step_y, sw, sh);
    pixel *= 1.0f / 20.0f;
    pixel *= sa;
    dst[x + y * tw] = __fmaf_rn(base[x + y * tw], sb, pixel);
  }
}

uint32_t lb_bloom_mip_count(uint32_t width, uint32_t height) {  // _device_post_bloom_mip_count, device_post.c:27-41
  uint32_t min_dim = width < height ? width : height, i = 0;
  if (min_dim == 0)
    return 0;
  while (min_dim != 1) {
    i++;
    min_dim >>= 1;
  }
  return i;
}

// _device_post_bloom_apply for undersampling stage 0: result = 3 planes of width * height mean radiance, mips[i] holds (width >> (i + 1)) x
// (height >> (i + 1)) floats. 3 * 2 * mip_count launches.
void lb_launch_bloom(float* result, uint32_t width, uint32_t height, float* const* mips, uint32_t mip_count, float blend, int grid, cudaStream_t s) {
  if (mip_count <= 1)
    return;  // "undersampling_stage + 1 >= bloom_mip_count": the chain is too short
  for (uint32_t ch = 0; ch < 3; ch++) {
    float* plane = result + (size_t) ch * width * height;
    k_post_downsample<<<grid, 256, 0, s>>>(plane, width, height, mips[0], width >> 1, height >> 1);
    for (uint32_t i = 0; i + 1 < mip_count; i++)
      k_post_downsample<<<grid, 256, 0, s>>>(mips[i], width >> (i + 1), height >> (i + 1), mips[i + 1], width >> (i + 2), height >> (i + 2));
    for (uint32_t i = mip_count - 1; i > 0; i--)
      k_post_upsample<<<grid, 256, 0, s>>>(mips[i], width >> (i + 1), height >> (i + 1), mips[i - 1], mips[i - 1], width >> i, height >> i, 1.0f, 1.0f);
    k_post_upsample<<<grid, 256, 0, s>>>(mips[0], width >> 1, height >> 1, plane, plane, width, height, blend / mip_count, 1.0f - blend);
  }
}

void lb_launch_output_argb8(const float* planes, uint32_t width, uint32_t height, uint32_t sample_count, const Lumb200OutputParams& op,
                            const uint16_t* bluenoise_1d, void* dst, int grid, cudaStream_t s, bool raw) {
  k_output_argb8<<<grid, 256, 0, s>>>(planes, width, height, 1.0f / sample_count, op, bluenoise_1d, (uchar4*) dst, raw);
}

// ---------------------------------------------------------------------------------------------
// BSDF directional-albedo LUTs (cuda/bsdf_lut.cuh:20-209): 65 536 samples per texel
// ---------------------------------------------------------------------------------------------
#define LUT_ITERATIONS 0x10000u

__device__ __forceinline__ uint16_t lut_quantise(float sum) { return (uint16_t) (1 + (uint16_t) (ceilf(__saturatef(sum) * 0xFFFE))); }

// The reference sums its 65 536 samples SERIALLY in one fp32 accumulator per texel (bsdf_lut.cuh:41-52), and nvcc contracts every
// `sum += a * b` of those loops into one FFMA (checked in the SASS of the reference build, oracle/_ref/librefdev.so). With terms of
// nearly constant size the round-to-nearest errors of that serial chain do not cancel: the tables carry a bias of up to 8e-4 that a
// pairwise / per-lane summation does not reproduce. To get the reference's tables, lanes evaluate 32 samples in parallel and then
// replay the reference's accumulation chain in sample order: sum = fma(a_j, b_j, sum), j = 0..31 (an invalid sample contributes
// fma(0, 0, sum) == sum).
__device__ __forceinline__ float lut_chain(float sum, float a, float b) {
#pragma unroll
  for (int j = 0; j < 32; j++)
    sum = __fmaf_rn(__shfl_sync(0xFFFFFFFFu, a, j), __shfl_sync(0xFFFFFFFFu, b, j), sum);
  return sum;
}

__global__ void __launch_bounds__(128) k_lut_conductor_glossy(const uint32_t* __restrict__ bluenoise, uint16_t* __restrict__ conductor,
                                                              uint16_t* __restrict__ glossy) {
  // one warp per texel, lanes stride the 65 536 samples; bsdf_generate_ss_lut + bsdf_generate_glossy_lut
  const uint32_t id   = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (id >= LB_LUT_SIZE * LB_LUT_SIZE)
    return;
  const uint32_t y      = id / LB_LUT_SIZE;
  const uint32_t x      = id - y * LB_LUT_SIZE;
  const float NdotV     = fmaxf(32.0f * EPS_F, x * (1.0f / (LB_LUT_SIZE - 1)));
  const float roughness = y * (1.0f / (LB_LUT_SIZE - 1));
  const V3 V            = norm3(v3(0.0f, sqrtf(1.0f - NdotV * NdotV), NdotV));
  const C3 f0           = c3(0.04f, 0.04f, 0.04f);
  const float r4        = pow4(roughness);
  float sum = 0.0f, sum_g = 0.0f;
  for (uint32_t s = lane; s < LUT_ITERATIONS; s += 32) {
    const uint2 q   = lbrng::random_2d_bits(bluenoise, lbrng::T_BSDF_REFLECTION, 0, 0, s, 0);
    const V3 H      = microfacet_sample_normal(V, roughness, make_float2(lbrng::u32_to_float(q.x), lbrng::u32_to_float(q.y)));
    const V3 R      = reflect3(V, H);
    float ca = 0.0f, cb = 0.0f, ga = 0.0f, gb = 0.0f;
    if (R.z > 0.0f) {
      // sum += 2 (k NdotV + t) * G2 * NdotL          -> fma(NdotL, 2 (k NdotV + t) * G2, sum)
      ca = vndf_norm(V, r4, NdotV) * smith_g2(r4, R.z, NdotV);
      cb = R.z;
      // sum += evaluate_sampled_microfacet * lum(F)  -> fma(e, lum, sum)
      ga = ca * cb;
      gb = c_lum(fresnel_schlick(f0, shadowed_f90(f0), fabsf(dot3(H, V))));
    }
    sum   = lut_chain(sum, ca, cb);
    sum_g = lut_chain(sum_g, ga, gb);
  }
  if (lane == 0) {
    const uint16_t c = lut_quantise(sum / LUT_ITERATIONS);
    conductor[id]    = c;
    glossy[id]       = lut_quantise((sum_g / LUT_ITERATIONS) / (c * (1.0f / 0xFFFF)));
  }
}

__global__ void __launch_bounds__(128) k_lut_dielectric(const uint32_t* __restrict__ bluenoise, uint16_t* __restrict__ dielectric,
                                                        uint16_t* __restrict__ dielectric_inv) {
  const uint32_t id   = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (id >= LB_LUT_SIZE * LB_LUT_SIZE * LB_LUT_SIZE)
    return;
  const uint32_t z      = id / (LB_LUT_SIZE * LB_LUT_SIZE);
  const uint32_t y      = (id - z * (LB_LUT_SIZE * LB_LUT_SIZE)) / LB_LUT_SIZE;
  const uint32_t x      = id - y * LB_LUT_SIZE - z * LB_LUT_SIZE * LB_LUT_SIZE;
  const float NdotV     = fmaxf(32.0f * EPS_F, x * (1.0f / (LB_LUT_SIZE - 1)));
  const float roughness = y * (1.0f / (LB_LUT_SIZE - 1));
  const float ior       = 1.0f + z * (1.0f / (LB_LUT_SIZE - 1)) * 2.0f;
  const V3 V            = norm3(v3(0.0f, sqrtf(1.0f - NdotV * NdotV), NdotV));
  const float r4        = pow4(roughness);
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    const float ratio = (pass == 0) ? 1.0f / ior : ior;
    float sum         = 0.0f;
    for (uint32_t s = lane; s < LUT_ITERATIONS; s += 32) {
      bool tot;
      const uint2 q1 = lbrng::random_2d_bits(bluenoise, lbrng::T_BSDF_REFLECTION, 0, 0, s, 0);
      V3 H           = microfacet_sample_normal(V, roughness, make_float2(lbrng::u32_to_float(q1.x), lbrng::u32_to_float(q1.y)));
      const V3 refl  = reflect3(V, H);
      V3 refr        = refract3(V, H, ratio, tot);
      float fresnel  = tot ? 1.0f : bsdf_fresnel(H, V, refr, ratio);
      float a1 = 0.0f, b1 = 0.0f, a2 = 0.0f, b2 = 0.0f;
      if (refl.z > 0.0f) {
        a1 = microfacet_eval_sampled_microfacet(V, roughness, refl.z, NdotV);
        b1 = fresnel;
      }
      const uint2 q2 = lbrng::random_2d_bits(bluenoise, lbrng::T_BSDF_REFRACTION, 0, 0, s, 0);
      H              = refraction_sample_normal(V, roughness, make_float2(lbrng::u32_to_float(q2.x), lbrng::u32_to_float(q2.y)));
      refr           = refract3(V, H, ratio, tot);
      // total reflection counts as fresnel 1 in the first table and 0 in the second (bsdf_lut.cuh:146,186)
      fresnel           = tot ? ((pass == 0) ? 1.0f : 0.0f) : bsdf_fresnel(H, V, refr, ratio);
      const float NdotR = -refr.z;
      if (NdotR > 0.0f) {
        a2 = smith_g2_over_g1(r4, NdotR, NdotV);
        b2 = 1.0f - fresnel;
      }
      // the reference adds the reflection term and then the refraction term of each sample (bsdf_lut.cuh:133-152)
#pragma unroll
      for (int j = 0; j < 32; j++) {
        sum = __fmaf_rn(__shfl_sync(0xFFFFFFFFu, a1, j), __shfl_sync(0xFFFFFFFFu, b1, j), sum);
        sum = __fmaf_rn(__shfl_sync(0xFFFFFFFFu, a2, j), __shfl_sync(0xFFFFFFFFu, b2, j), sum);
      }
    }
    if (lane == 0) {
      if (pass == 0)
        dielectric[id] = lut_quantise(sum / LUT_ITERATIONS);
      else
        dielectric_inv[id] = lut_quantise(sum / LUT_ITERATIONS);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static_assert(LB_RNG_TARGET_COUNT == lbrng::T_COUNT, "table stride must equal the number of random targets");

__global__ void __launch_bounds__(256) k_rng_table(uint4* __restrict__ table, uint32_t sample_id, uint32_t num_dims) {
  const uint32_t dim = blockIdx.x * blockDim.x + threadIdx.x;
  if (dim < num_dims)
    table[dim] = lbrng::table_entry(sample_id, dim);
}

void lb_launch_rng_table(uint4* table, uint32_t sample_id, uint32_t depths, cudaStream_t s) {
  const uint32_t num_dims = depths * LB_RNG_TARGET_COUNT;
  k_rng_table<<<(num_dims + 255) / 256, 256, 0, s>>>(table, sample_id, num_dims);
}

// One launch per material class that the scene uses (class_materials[c] > 0; class ranges live in LbCounters.class_begin), one for the
// misses. The textured variants run only when a material of the scene references a texture.
template <int kClass, bool kSun>
static void launch_shade_class_sun(const LbShadeParams& sp, int grid, cudaStream_t s) {
  if (sp.adaptive) {
    if (sp.textured)
      k_shade<kClass, true, true, false, kSun><<<grid * (128 / LB_SHADE_THREADS), LB_SHADE_THREADS, 0, s>>>(sp);
    else
      k_shade<kClass, false, true, false, kSun><<<grid * (128 / LB_SHADE_THREADS), LB_SHADE_THREADS, 0, s>>>(sp);
  }
  else if (sp.count) {
    if (sp.textured)
      k_shade<kClass, true, false, true, kSun><<<grid * (128 / LB_SHADE_THREADS), LB_SHADE_THREADS, 0, s>>>(sp);
    else
      k_shade<kClass, false, false, true, kSun><<<grid * (128 / LB_SHADE_THREADS), LB_SHADE_THREADS, 0, s>>>(sp);
  }
  else if (sp.textured)
    k_shade<kClass, true, false, false, kSun><<<grid * (128 / LB_SHADE_THREADS), LB_SHADE_THREADS, 0, s>>>(sp);
  else
    k_shade<kClass, false, false, false, kSun><<<grid * (128 / LB_SHADE_THREADS), LB_SHADE_THREADS, 0, s>>>(sp);
}

// the sun's NEE is compiled into a second set of instantiations: scenes under a constant-colour sky keep the leaner kernels
template <int kClass>
static void launch_shade_class(const LbShadeParams& sp, int grid, cudaStream_t s) {
  if (sp.frame.sky_mode != 2)  // direct_lighting_sun_is_allowed: sky.mode != CONSTANT_COLOR
    launch_shade_class_sun<kClass, true>(sp, grid, s);
  else
    launch_shade_class_sun<kClass, false>(sp, grid, s);
}

int lb_launch_shade(const LbShadeParams& sp, int grid, cudaStream_t s, cudaStream_t aux, cudaEvent_t fork, cudaEvent_t join) {
  int launches = 0;
  // the kernels of one bounce touch disjoint paths and append to the queues through atomics: they may run side by side
  cudaStream_t side = s;
  if (aux && fork && join && cudaEventRecord(fork, s) == cudaSuccess && cudaStreamWaitEvent(aux, fork, 0) == cudaSuccess)
    side = aux;
  if (sp.class_materials[LB_CLASS_DIELECTRIC]) {
    launch_shade_class<LB_CLASS_DIELECTRIC>(sp, grid, s);
    launches++;
  }
  if (sp.class_materials[LB_CLASS_METAL]) {
    launch_shade_class<LB_CLASS_METAL>(sp, grid, side);
    launches++;
  }
  if (sp.class_materials[LB_CLASS_GENERIC]) {
    launch_shade_class<LB_CLASS_GENERIC>(sp, grid, s);
    launches++;
  }
  if (sp.frame.sky_mode == 2) {
    k_shade_miss<<<max(grid / 8, 1), 256, 0, side>>>(sp);  // a few instructions per miss: the wide k_shade grid only costs block launches here
    launches++;
  }
  else {
    lb_launch_shade_miss_sky(sp, grid, side);
    launches++;
  }
  if (side != s) {
    cudaEventRecord(join, side);
    cudaStreamWaitEvent(s, join, 0);
  }
  return launches;
}

void lb_launch_enum_finish(const LbShadeParams& sp, int grid, cudaStream_t s) {
  if (sp.textured)
    k_enum_finish<true><<<grid, 128, 0, s>>>(sp);
  else
    k_enum_finish<false><<<grid, 128, 0, s>>>(sp);
}

// parity hook of lumb200_device_sample_texture(_lod): raw tex2DLod<float4> (no flip, no gamma), one uv pair per thread
__global__ void k_sample_texture(const LbTexture* __restrict__ textures, uint32_t num_textures, uint32_t tex, const float2* __restrict__ uv, uint32_t n,
                                 float lod, float4* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  LbTexture t;
  out[i] = lb_texture_valid(textures, num_textures, tex, t) ? tex2DLod<float4>(t.handle, uv[i].x, uv[i].y, lod) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// mipmap_generate_level_2D_RGBA8 / RGBA16 / RGBAF (cuda/mipmap.cuh:39-72,107-140,166-187): every texel of level l + 1 is one
// filtered fetch of level l at the texel's centre; integer formats are re-quantised with round-half-up and keep a non-zero
// alpha non-zero ("for opacity micromaps", :61-62). type: 0 fp32, 1 u8, 2 u16 (Lumb200Texture.type).
__global__ void __launch_bounds__(256) k_mipmap_level(cudaTextureObject_t src, cudaSurfaceObject_t dst, uint32_t width, uint32_t height, uint32_t type) {
  const float scale_x = 1.0f / width, scale_y = 1.0f / height;
  for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < width * height; id += gridDim.x * blockDim.x) {
    const uint32_t y = id / width, x = id - y * width;
    float4 v         = tex2D<float4>(src, scale_x * (x + 0.5f), scale_y * (y + 0.5f));
    if (type == 0) {
      surf2Dwrite(v, dst, x * sizeof(float4), y);
      continue;
    }
    const float full = (type == 1) ? 255.0f : 65535.0f;
    const float top  = full + 0.9f;
    v.w *= full;
    v.w = (v.w > 0.0f) ? fmaxf(v.w, 0.51f) : v.w;
    v.x = fminf(__fmaf_rn(v.x, full, 0.5f), top);
    v.y = fminf(__fmaf_rn(v.y, full, 0.5f), top);
    v.z = fminf(__fmaf_rn(v.z, full, 0.5f), top);
    v.w = fminf(v.w + 0.5f, top);
    if (type == 1)
      surf2Dwrite(make_uchar4((uint8_t) v.x, (uint8_t) v.y, (uint8_t) v.z, (uint8_t) v.w), dst, x * sizeof(uchar4), y);
    else
      surf2Dwrite(make_ushort4((uint16_t) v.x, (uint16_t) v.y, (uint16_t) v.z, (uint16_t) v.w), dst, x * sizeof(ushort4), y);
  }
}

void lb_launch_mipmap_level(cudaTextureObject_t src, cudaSurfaceObject_t dst, uint32_t width, uint32_t height, uint32_t type, cudaStream_t s) {
  const uint32_t n = width * height;
  if (n)
    k_mipmap_level<<<(n + 255u) / 256u, 256, 0, s>>>(src, dst, width, height, type);
}

// ---------------------------------------------------------------------------------------------
// light_compute_intensity (cuda/light.cuh:190-262, light_microtriangle.cuh:8-64): one warp per emitter triangle, two of the
// 64 micro-triangles per lane; each micro-triangle is scanned in texel-sized steps and the largest colour importance wins.
// ---------------------------------------------------------------------------------------------
// light_microtriangle_id_to_bary: row r of the 8-row subdivision covers ids (T[r-1], T[r]] with the reference's thresholds
// T = 15, 28, 39, 48, 55, 60, 63 (light_microtriangle.cuh:12-49 - note that these are not the row starts 0, 15, 28, ...: the
// first id of rows 1..6 is attributed to the row before, which the column formula (id - S[row]) >> 1 reproduces as written).
__device__ __forceinline__ void microtriangle_bary(uint32_t id, float2& b0, float2& b1, float2& b2) {
  const uint32_t T[7] = {15u, 28u, 39u, 48u, 55u, 60u, 63u};
  const uint32_t S[8] = {0u, 15u, 28u, 39u, 48u, 55u, 60u, 63u};
  uint32_t row        = 7;
#pragma unroll
  for (int r = 6; r >= 0; r--)
    if (id <= T[r])
      row = (uint32_t) r;
  const uint32_t col = (row == 7) ? 0u : ((id - S[row]) >> 1);
  const bool is_top  = (id & 1u) == (row & 1u);
  b0 = make_float2((float) row, (float) (col + 1));
  b1 = make_float2((float) (row + 1), (float) col);
  b2 = is_top ? make_float2((float) row, (float) col) : make_float2((float) (row + 1), (float) (col + 1));
  b0.x *= 0.125f, b0.y *= 0.125f, b1.x *= 0.125f, b1.y *= 0.125f, b2.x *= 0.125f, b2.y *= 0.125f;
}

__device__ float microtriangle_max_emission(const LbTexture& t, float2 vertex, float2 e1, float2 e2, uint32_t id) {  // lights_get_max_emission
  float2 b0, b1, b2;
  microtriangle_bary(id, b0, b1, b2);
  const float2 u0 = make_float2(vertex.x + b0.x * e1.x + b0.y * e2.x, vertex.y + b0.x * e1.y + b0.y * e2.y);
  const float2 u1 = make_float2(vertex.x + b1.x * e1.x + b1.y * e2.x, vertex.y + b1.x * e1.y + b1.y * e2.y);
  const float2 u2 = make_float2(vertex.x + b2.x * e1.x + b2.y * e2.x, vertex.y + b2.x * e1.y + b2.y * e2.y);
  const float2 m1 = make_float2(u1.x - u0.x, u1.y - u0.y), m2 = make_float2(u2.x - u0.x, u2.y - u0.y);
  const float su  = fmaxf(fabsf(m1.x), fabsf(m2.x)) * (float) (t.size & 0xFFFFu);
  const float sv  = fmaxf(fabsf(m1.y), fabsf(m2.y)) * (float) (t.size >> 16);
  const float step = 1.0f / ceilf(fmaxf(su, sv));
  C3 mx            = c3(0.0f, 0.0f, 0.0f);
  for (float a = 0.0f; a < 1.0f; a += step)
    for (float b = 0.0f; a + b < 1.0f; b += step) {
      const float4 texel = lb_texture_fetch(t, make_float2(u0.x + a * m1.x + b * m2.x, u0.y + a * m1.y + b * m2.y), true, true);
      mx                 = c3(fmaxf(mx.r, texel.x), fmaxf(mx.g, texel.y), fmaxf(mx.b, texel.z));
    }
  return fmaxf(mx.r, fmaxf(mx.g, mx.b));  // color_importance
}

__global__ void __launch_bounds__(128) k_light_compute_intensity(LbShadeParams P, const uint32_t* __restrict__ mesh_ids,
                                                                 const uint32_t* __restrict__ tri_ids, uint32_t count, float* __restrict__ out) {
  const uint32_t light = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (light >= count)
    return;
  const uint32_t micro = (threadIdx.x & 31u) << 1;
  const uint4 tri      = __ldg(P.mesh_textris[mesh_ids[light]] + tri_ids[light]);
  const float2 t0 = lb_uv_unpack(tri.x), t1 = lb_uv_unpack(tri.y), t2 = lb_uv_unpack(tri.z);
  const Mat m     = load_material(P.materials, tri.w & 0xFFFFu);
  float best      = 0.0f;
  LbTexture t;
  if (lb_texture_valid(P.textures, P.num_textures, m.luminance_tex, t)) {
    const float2 e1 = make_float2(t1.x - t0.x, t1.y - t0.y), e2 = make_float2(t2.x - t0.x, t2.y - t0.y);
    best            = fmaxf(microtriangle_max_emission(t, t0, e1, e2, micro), microtriangle_max_emission(t, t0, e1, e2, micro + 1));
  }
  for (int o = 16; o > 0; o >>= 1)
    best = fmaxf(best, __shfl_xor_sync(0xFFFFFFFFu, best, o));
  if (micro == 0)
    out[light] = best;
}

void lb_launch_light_compute_intensity(const LbShadeParams& sp, const uint32_t* mesh_ids, const uint32_t* tri_ids, uint32_t count, float* out,
                                       cudaStream_t s) {
  if (count)
    k_light_compute_intensity<<<(count * 32u + 127u) / 128u, 128, 0, s>>>(sp, mesh_ids, tri_ids, count, out);
}

void lb_launch_sample_texture(const LbTexture* textures, uint32_t num_textures, uint32_t tex, const float2* uv, uint32_t n, float lod, float4* out,
                              cudaStream_t s) {
  if (n)
    k_sample_texture<<<(n + 127u) / 128u, 128, 0, s>>>(textures, num_textures, tex, uv, n, lod, out);
}

void lb_launch_accumulate(const LbPaths& P, const LbFrame& F, float* planes, LbCounters* counters, int grid, cudaStream_t s) {
  k_accumulate<<<grid, 256, 0, s>>>(P, F.width * F.height, planes, counters);
}

void lb_launch_generate_result(const float* planes, float* result, uint32_t num_pixels, uint32_t sample_count, int grid, cudaStream_t s) {
  k_generate_result<<<grid, 256, 0, s>>>(planes, result, num_pixels, 1.0f / sample_count);
}

// ---------------------------------------------------------------------------------------------
// adaptive sampling (cuda/adaptive_sampling.cuh, cuda/kernels.cuh:195-356, cuda/accumulation.cuh): the image is tiled into 4 x 4
// pixel blocks; byte s of a block's word holds (samples per pixel and execution in stage s + 1) - 1.
// ---------------------------------------------------------------------------------------------
// accumulation_generate_result (cuda/accumulation.cuh:86-190), all four output modes + the camera's local error minimisation.
// A.words == nullptr: adaptive sampling is off, every pixel has `uniform_count` samples.
struct ResolvePixel {
  C3 mean;
  float variance;
  float inv_n;
};

__device__ __forceinline__ ResolvePixel resolve_pixel(const float* __restrict__ planes, uint32_t width, uint32_t height, uint32_t x, uint32_t y,
                                                      const LbAdaptive& A, uint32_t uniform_count) {
  // adaptive_sampling_get_pixel_variance_and_color, adaptive_sampling.cuh:145-164
  const size_t n    = (size_t) width * height;
  const size_t i    = x + (size_t) y * width;
  const uint32_t cnt = A.words ? as_block_samples(__ldg(A.words + (x >> 2) + (y >> 2) * A.bw), A) : uniform_count;
  ResolvePixel r;
  r.inv_n    = 1.0f / (float) cnt;
  r.mean     = c3(planes[i] * r.inv_n, planes[n + i] * r.inv_n, planes[2 * n + i] * r.inv_n);
  r.variance = fmaxf(planes[3 * n + i] * r.inv_n - c_lum(c3(r.mean.r * r.mean.r, r.mean.g * r.mean.g, r.mean.b * r.mean.b)), 0.0f);
  return r;
}

__global__ void __launch_bounds__(256) k_resolve(const float* __restrict__ planes, float* __restrict__ result, uint32_t width, uint32_t height,
                                                 LbAdaptive A, uint32_t uniform_count, uint32_t mode, uint32_t local_error_minimization,
                                                 uint32_t stage, Lumb200OutputParams tm) {
  const uint32_t n = width * height;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t y = i / width, x = i - y * width;
    const ResolvePixel c = resolve_pixel(planes, width, height, x, y, A, uniform_count);
    C3 out = c.mean;
    if (mode == 0 && local_error_minimization) {  // :105-143: blend towards the neighbourhood mean where the pixel's own error dominates
      const float center_error = c.variance * c.inv_n;
      const uint32_t x0 = max(x, 1u) - 1u, x1 = min(x, width - 1u) + 1u, y0 = max(y, 1u) - 1u, y1 = min(y, height - 1u) + 1u;
      C3 nmean     = c3(0.0f, 0.0f, 0.0f);
      float nerror = 0.0f;
      for (uint32_t yi = y0; yi <= y1; yi++)
        for (uint32_t xi = x0; xi <= x1; xi++) {
          if ((xi == x && yi == y) || xi >= width || yi >= height)
            continue;
          const ResolvePixel q = resolve_pixel(planes, width, height, xi, yi, A, uniform_count);
          nmean                = nmean + q.mean;
          nerror += q.variance * q.inv_n;
        }
      // the reference divides by the size of the UNCLAMPED 3 x 3 window minus one (accumulation.cuh:135), border pixels included
      const float nn = 1.0f / (float) ((x1 - x0 + 1u) * (y1 - y0 + 1u) - 1u);
      nmean          = nmean * nn;
      nerror *= nn;
      const float t = __saturatef(center_error / (8.0f * nerror));
      out           = c3(c.mean.r + t * (nmean.r - c.mean.r), c.mean.g + t * (nmean.g - c.mean.g), c.mean.b + t * (nmean.b - c.mean.b));
    }
    else if (mode == 1) {  // variance view
      const float v = 128.0f * c.variance;
      out           = c3(v, v, v);
    }
    else if (mode == 2) {  // error view: standard error after the tone map's compression, false colours (:159-173)
      const float ev    = c_lum(c.mean * tm.exposure);
      const float tv    = c_lum(tonemap_pixel(c.mean, tm));
      const float comp  = (ev > 0.0f) ? tv / ev : 1.0f;
      const float value = 1024.0f * (sqrtf(c.variance * c.inv_n) * comp);
      out = c3(__saturatef(2.0f * value), __saturatef(2.0f * (value - 0.5f)),
               __saturatef((value > 0.5f) ? 4.0f * (0.25f - fabsf(value - 1.0f)) : 4.0f * (0.25f - fabsf(value - 0.25f))));
    }
    else if (mode == 3) {  // sample distribution of the current stage (:175-181)
      const uint32_t tpp = (A.words && stage > 0) ? as_stage_count(__ldg(A.words + (x >> 2) + (y >> 2) * A.bw), stage - 1u) : 1u;
      const float v      = (float) tpp / 256.0f;
      out                = c3(v, v, v);
    }
    result[i]                  = out.r;
    result[(size_t) n + i]     = out.g;
    result[2 * (size_t) n + i] = out.b;
  }
}

void lb_launch_resolve(const float* planes, float* result, uint32_t width, uint32_t height, const LbAdaptive& A, uint32_t uniform_count,
                       uint32_t mode, uint32_t local_error_minimization, uint32_t stage, const Lumb200OutputParams& tm, int grid, cudaStream_t s) {
  k_resolve<<<grid, 256, 0, s>>>(planes, result, width, height, A, uniform_count, mode, local_error_minimization, stage, tm);
}

// adaptive_sampling_block_reduce_variance (:166-199): one thread per block, max pixel variance (x tone-map compression^2 when
// exposure-aware), summed over all blocks with one atomic per block like the reference
__global__ void __launch_bounds__(128) k_as_block_variance(const float* __restrict__ planes, uint32_t width, uint32_t height, LbAdaptive A,
                                                           Lumb200OutputParams tm, float* __restrict__ block_variance, float* __restrict__ sum) {
  const uint32_t bh = (height + 3u) >> 2;
  const uint32_t b  = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= A.bw * bh)
    return;
  const uint32_t by = b / A.bw, bx = b - by * A.bw;
  const size_t n    = (size_t) width * height;
  const float inv   = 1.0f / (float) as_block_samples(__ldg(A.words + b), A);
  float best        = 0.0f;
  for (uint32_t l = 0; l < 16u; l++) {
    const uint32_t x = 4u * bx + (l & 3u), y = 4u * by + (l >> 2);
    if (x >= width || y >= height)
      continue;
    const size_t i = x + (size_t) y * width;
    const C3 m     = c3(planes[i] * inv, planes[n + i] * inv, planes[2 * n + i] * inv);
    float var      = fmaxf(planes[3 * n + i] * inv - c_lum(c3(m.r * m.r, m.g * m.g, m.b * m.b)), 0.0f);
    if (tm.exposure != 0.0f) {  // adaptive_sampling_compute_tonemap_compression_factor, :9-17
      const float ev = c_lum(m * tm.exposure);
      const float tv = c_lum(tonemap_pixel(m, tm));
      const float c  = (ev > 0.0f) ? tv / ev : 1.0f;
      var *= c * c;
    }
    best = fmaxf(best, var);
  }
  best              = fabsf(best);
  block_variance[b] = best;
  atomicAdd(sum, best);
}

// adaptive_sampling_compute_stage_sample_counts (:201-221) + adaptive_sampling_compute_tasks_per_block (:243-256)
__global__ void __launch_bounds__(128) k_as_stage_counts(const float* __restrict__ block_variance, const float* __restrict__ sum,
                                                         uint32_t num_blocks, uint32_t stage, uint32_t max_rate, uint32_t avg_rate,
                                                         uint32_t* __restrict__ words, uint32_t* __restrict__ tasks_per_block) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= num_blocks)
    return;
  const float avg = *sum / (float) num_blocks;
  uint32_t w      = words[b] & ((1u << (stage * 8u)) - 1u);
  uint32_t c      = (uint32_t) (block_variance[b] / avg * (float) avg_rate + 0.5f);  // NaN -> 0
  c               = min(max(c, 1u), max_rate);
  words[b]           = w | ((c - 1u) << (stage * 8u));
  tasks_per_block[b] = c * 16u;
}

// inclusive prefix sum of the tasks per block (adaptive_sampling_compute_block_sum / _prefix_sum, :258-291): one block, chunked
__global__ void __launch_bounds__(1024) k_as_prefix_sum(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0)
    carry = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < n; base += 1024u) {
    const uint32_t i = base + threadIdx.x;
    uint32_t v       = (i < n) ? data[i] : 0u;
    for (uint32_t o = 1; o < 32u; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
      if (lane >= o)
        v += t;
    }
    if (lane == 31u)
      warp_sums[warp] = v;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_sums[lane];
      for (uint32_t o = 1; o < 32u; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o);
        if (lane >= o)
          w += t;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t offset = carry + (warp ? warp_sums[warp - 1] : 0u);
    if (i < n)
      data[i] = v + offset;
    __syncthreads();
    if (threadIdx.x == 1023u)
      carry = v + offset;
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *total = carry;
}

void lb_launch_adaptive_build_stage(const float* planes, uint32_t width, uint32_t height, const LbAdaptive& A, const Lumb200OutputParams& tm,
                                    uint32_t stage, uint32_t max_rate, uint32_t avg_rate, uint32_t* words, float* block_variance, float* sum,
                                    uint32_t* task_prefix, uint32_t* total_tasks, cudaStream_t s) {
  const uint32_t num_blocks = A.bw * ((height + 3u) >> 2);
  cudaMemsetAsync(sum, 0, sizeof(float), s);
  k_as_block_variance<<<(num_blocks + 127u) / 128u, 128, 0, s>>>(planes, width, height, A, tm, block_variance, sum);
  k_as_stage_counts<<<(num_blocks + 127u) / 128u, 128, 0, s>>>(block_variance, sum, num_blocks, stage, max_rate, avg_rate, words, task_prefix);
  k_as_prefix_sum<<<1, 1024, 0, s>>>(task_prefix, num_blocks, total_tasks);
}

// accumulation_collect_results (accumulation.cuh:36-60): several paths per pixel and launch -> atomics
__global__ void __launch_bounds__(256) k_accumulate_adaptive(LbPaths paths, uint32_t n_slots, uint32_t n_pixels, float* __restrict__ planes,
                                                             LbCounters* __restrict__ counters) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) {
    const uint32_t pix = paths.pixel[i];
    if (pix == 0xFFFFFFFFu)
      continue;
    float4 r = paths.result[i];
#pragma unroll
    for (int s = 0; s < LB_NEE_SLOTS; s++) {
      const float4 a = paths.nee[LB_NEE_SLOTS * (size_t) i + s];
      r.x += a.x, r.y += a.y, r.z += a.z;
    }
    if (non_finite(r.x + r.y + r.z)) {
      atomicAdd(&counters->nonfinite_samples, 1u);
      counters->nonfinite_pixel = pix;
      continue;
    }
    atomicAdd(planes + pix, r.x);
    atomicAdd(planes + (size_t) n_pixels + pix, r.y);
    atomicAdd(planes + 2 * (size_t) n_pixels + pix, r.z);
    atomicAdd(planes + 3 * (size_t) n_pixels + pix, c_lum(c3(r.x * r.x, r.y * r.y, r.z * r.z)));
  }
}

void lb_launch_accumulate_adaptive(const LbPaths& P, uint32_t n_slots, uint32_t n_pixels, float* planes, LbCounters* counters, int grid, cudaStream_t s) {
  k_accumulate_adaptive<<<grid, 256, 0, s>>>(P, n_slots, n_pixels, planes, counters);
}

static const size_t kLutElems[4] = {LB_LUT_SIZE * LB_LUT_SIZE, LB_LUT_SIZE* LB_LUT_SIZE, LB_LUT_SIZE* LB_LUT_SIZE* LB_LUT_SIZE,
                                    LB_LUT_SIZE* LB_LUT_SIZE* LB_LUT_SIZE};

void lb_lut_destroy(LbLutTextures* luts) {
  cudaTextureObject_t* tex[4] = {&luts->tex.conductor, &luts->tex.glossy, &luts->tex.dielectric, &luts->tex.dielectric_inv};
  for (int k = 0; k < 4; k++) {
    if (*tex[k])
      cudaDestroyTextureObject(*tex[k]);
    *tex[k] = 0;
    if (luts->arrays[k])
      cudaFreeArray(luts->arrays[k]);
    luts->arrays[k] = nullptr;
    if (luts->d_data[k])
      cudaFree(luts->d_data[k]);
    luts->d_data[k] = nullptr;
  }
  luts->valid = false;
}

static Lumb200Result lut_alloc(LbLutTextures* luts) {
  if (luts->d_data[0])
    return LUMB200_SUCCESS;
  for (int k = 0; k < 4; k++)
    LB_CHECK(cudaMalloc(&luts->d_data[k], sizeof(uint16_t) * kLutElems[k]));
  return LUMB200_SUCCESS;
}

// R16 unorm arrays, linear filtering, clamp, normalised coordinates (device_bsdf.c:7-54, device_texture.c:262-271)
static Lumb200Result lut_make_textures(LbLutTextures* luts, cudaStream_t s) {
  cudaTextureObject_t* tex[4] = {&luts->tex.conductor, &luts->tex.glossy, &luts->tex.dielectric, &luts->tex.dielectric_inv};
  const cudaChannelFormatDesc fmt = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
  LB_CHECK(cudaStreamSynchronize(s));
  for (int k = 0; k < 4; k++) {
    const bool is3d = k >= 2;
    if (*tex[k]) {
      cudaDestroyTextureObject(*tex[k]);
      *tex[k] = 0;
    }
    if (!luts->arrays[k]) {
      if (is3d)
        LB_CHECK(cudaMalloc3DArray(&luts->arrays[k], &fmt, make_cudaExtent(LB_LUT_SIZE, LB_LUT_SIZE, LB_LUT_SIZE)));
      else
        LB_CHECK(cudaMallocArray(&luts->arrays[k], &fmt, LB_LUT_SIZE, LB_LUT_SIZE));
    }
    if (is3d) {
      cudaMemcpy3DParms cp;
      memset(&cp, 0, sizeof(cp));
      cp.srcPtr   = make_cudaPitchedPtr(luts->d_data[k], LB_LUT_SIZE * sizeof(uint16_t), LB_LUT_SIZE, LB_LUT_SIZE);
      cp.dstArray = luts->arrays[k];
      cp.extent   = make_cudaExtent(LB_LUT_SIZE, LB_LUT_SIZE, LB_LUT_SIZE);
      cp.kind     = cudaMemcpyDeviceToDevice;
      LB_CHECK(cudaMemcpy3D(&cp));
    }
    else {
      LB_CHECK(cudaMemcpy2DToArray(luts->arrays[k], 0, 0, luts->d_data[k], LB_LUT_SIZE * sizeof(uint16_t), LB_LUT_SIZE * sizeof(uint16_t),
                                   LB_LUT_SIZE, cudaMemcpyDeviceToDevice));
    }
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType         = cudaResourceTypeArray;
    rd.res.array.array = luts->arrays[k];
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode       = cudaFilterModeLinear;
    td.readMode         = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    LB_CHECK(cudaCreateTextureObject(tex[k], &rd, &td, nullptr));
  }
  luts->valid = true;
  return LUMB200_SUCCESS;
}

Lumb200Result lb_lut_generate(LbLutTextures* luts, const uint32_t* bluenoise, cudaStream_t s) {
  Lumb200Result r = lut_alloc(luts);
  if (r != LUMB200_SUCCESS)
    return r;
  k_lut_conductor_glossy<<<(LB_LUT_SIZE * LB_LUT_SIZE * 32 + 127) / 128, 128, 0, s>>>(bluenoise, luts->d_data[0], luts->d_data[1]);
  k_lut_dielectric<<<(LB_LUT_SIZE * LB_LUT_SIZE * LB_LUT_SIZE * 32 + 127) / 128, 128, 0, s>>>(bluenoise, luts->d_data[2], luts->d_data[3]);
  LB_CHECK(cudaGetLastError());
  return lut_make_textures(luts, s);
}

Lumb200Result lb_lut_upload(LbLutTextures* luts, const uint16_t* conductor, const uint16_t* glossy, const uint16_t* dielectric,
                            const uint16_t* dielectric_inv, cudaStream_t s) {
  Lumb200Result r = lut_alloc(luts);
  if (r != LUMB200_SUCCESS)
    return r;
  const uint16_t* src[4] = {conductor, glossy, dielectric, dielectric_inv};
  for (int k = 0; k < 4; k++)
    LB_CHECK(cudaMemcpyAsync(luts->d_data[k], src[k], sizeof(uint16_t) * kLutElems[k], cudaMemcpyHostToDevice, s));
  return lut_make_textures(luts, s);
}

Lumb200Result lb_lut_download(LbLutTextures* luts, uint16_t* conductor, uint16_t* glossy, uint16_t* dielectric, uint16_t* dielectric_inv,
                              cudaStream_t s) {
  uint16_t* dst[4] = {conductor, glossy, dielectric, dielectric_inv};
  for (int k = 0; k < 4; k++)
    if (dst[k])
      LB_CHECK(cudaMemcpyAsync(dst[k], luts->d_data[k], sizeof(uint16_t) * kLutElems[k], cudaMemcpyDeviceToHost, s));
  LB_CHECK(cudaStreamSynchronize(s));
  return LUMB200_SUCCESS;
}
