// sky.cuh - procedural atmosphere of the miss shader and the sun's direct lighting (LUMINARY_SKY_MODE_DEFAULT).
//
// Replaces, on the reference's per-bounce path:
//   sky_color_main -> sky_get_color -> sky_compute_atmosphere   cuda/sky.cuh:338-515, 567-601   (ray-marched single scattering, 8 wavelengths,
//                                                                transmittance + multiscattering LUTs, sun disc, moon occlusion, stars)
//   sky_get_sun_color                                            cuda/sky_utils.cuh:322-349       (radiance of the sun disc behind the transmittance LUT)
//   sky_compute_transmittance_lut / _multiscattering_lut         cuda/sky.cuh:109-330             (LUT generation, device_sky.c:80-117)
//   sphere / phase function helpers                              cuda/math.cuh:330-347, 620-779, 1167-1239, 1393-1439
// Not on this path: clouds (cloud_shadow), aerial perspective (sky_trace_inscattering), the HDRI mode, the moon's surface textures
// (data/moon/*.png of the reference; the moon occludes the sun and the stars, its disc is black).
// The arithmetic keeps the reference's operation order; compiled with --use_fast_math like the reference's kernels.
#pragma once

#include <float.h>

#include "lumb200_internal.cuh"
#include "texture.cuh"

// sky_defines.h
#define LB_SKY_EARTH_RADIUS 6371.0f
#define LB_SKY_SUN_RADIUS 696340.0f
#define LB_SKY_SUN_DISTANCE 149597870.0f
#define LB_SKY_MOON_RADIUS 1737.4f
#define LB_SKY_MOON_DISTANCE 384399.0f
#define LB_SKY_ATMO_HEIGHT 100.0f
#define LB_SKY_ATMO_RADIUS (LB_SKY_ATMO_HEIGHT + LB_SKY_EARTH_RADIUS)
#define LB_SKY_MS_TEX_SIZE 32
#define LB_SKY_TM_TEX_WIDTH 256
#define LB_SKY_TM_TEX_HEIGHT 64
#define LB_SKY_MS_BASE 16
#define LB_SKY_MS_ITER (LB_SKY_MS_BASE * LB_SKY_MS_BASE)
#define LB_SKY_HEIGHT_OFFSET 0.0005f
#define LB_STARS_GRID_X 64  // STARS_GRID_LD, device_utils.h:41
#define LB_STARS_GRID_Y 32
#define LB_SKY_PI 3.141592653589f

// DeviceSky (device_structs.h:101-125) + the LUT texture objects and the star catalogue of DeviceConstantMemory
struct LbSkyDev {
  uint32_t mode;  // LuminarySkyMode: 0 procedural, 1 baked table (HDRI), 2 constant colour
  uint32_t steps;
  uint32_t ozone_absorption;
  uint32_t has_stars;
  float geometry_offset[3];
  float sun_strength, base_density, stars_intensity, rayleigh_density, mie_density, ozone_density, rayleigh_falloff, mie_falloff, mie_diameter,
    ground_visibility, ozone_layer_thickness, multiscattering_factor;
  float sun_pos[3], moon_pos[3];
  cudaTextureObject_t tm_low, tm_high, ms_low, ms_high;  // 256 x 64 and 32 x 32, float4, linear, clamp, normalised coordinates
  const float4* stars;                                   // Star {altitude, azimuth, radius, intensity}, sorted by grid cell
  const uint32_t* stars_offsets;                         // [64 * 32 + 1]
  cudaTextureObject_t hdri;                              // mode 1: hdri_dim^2 float4, point filter, wrap, normalised coordinates
  uint32_t aerial_perspective;                           // in-scattering along every hit segment (sky_process_inscattering_events)
  float moon_tex_offset;
  LbTexture moon_albedo, moon_normal;                    // the moon's surface (device_load_embedded_data); handle 0 = absent: a black disc
};

namespace lbsky {

struct Spectrum {
  float v[8];
};

#define LB_SPECTRUM_OP(expr)     \
  Spectrum r;                    \
  _Pragma("unroll") for (int i = 0; i < 8; i++) r.v[i] = (expr); \
  return r;

__device__ __forceinline__ Spectrum s_set1(float a) { LB_SPECTRUM_OP(a) }
__device__ __forceinline__ Spectrum s_add(const Spectrum& a, const Spectrum& b) { LB_SPECTRUM_OP(a.v[i] + b.v[i]) }
__device__ __forceinline__ Spectrum s_sub(const Spectrum& a, const Spectrum& b) { LB_SPECTRUM_OP(a.v[i] - b.v[i]) }
__device__ __forceinline__ Spectrum s_mul(const Spectrum& a, const Spectrum& b) { LB_SPECTRUM_OP(a.v[i] * b.v[i]) }
__device__ __forceinline__ Spectrum s_scale(const Spectrum& a, float b) { LB_SPECTRUM_OP(a.v[i] * b) }
__device__ __forceinline__ Spectrum s_inv(const Spectrum& a) { LB_SPECTRUM_OP(1.0f / a.v[i]) }
__device__ __forceinline__ Spectrum s_exp(const Spectrum& a) { LB_SPECTRUM_OP(expf(a.v[i])) }

__device__ __forceinline__ Spectrum s_set(float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7) {
  Spectrum r;
  r.v[0] = v0, r.v[1] = v1, r.v[2] = v2, r.v[3] = v3, r.v[4] = v4, r.v[5] = v5, r.v[6] = v6, r.v[7] = v7;
  return r;
}
__device__ __forceinline__ Spectrum s_merge(float4 low, float4 high) { return s_set(low.x, low.y, low.z, low.w, high.x, high.y, high.z, high.w); }
__device__ __forceinline__ float4 s_low(const Spectrum& a) { return make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }
__device__ __forceinline__ float4 s_high(const Spectrum& a) { return make_float4(a.v[4], a.v[5], a.v[6], a.v[7]); }

// sky_utils.cuh:105-119, 255-272
__device__ __forceinline__ Spectrum s_ident() {
  return s_set(8.4205e-03f, 2.6449e-01f, 4.0273e-01f, 1.6624e-01f, 2.4324e-01f, 3.5849e-01f, 3.6342e-01f, 2.4177e-01f);
}
__device__ __forceinline__ Spectrum sun_radiance() {
  return s_set(2.463170e+04f, 2.888721e+04f, 2.795153e+04f, 2.629836e+04f, 2.667237e+04f, 2.638737e+04f, 2.490630e+04f, 2.338930e+04f);
}
__device__ __forceinline__ Spectrum rayleigh_scattering() {
  return s_set(3.945800e-02f, 2.939289e-02f, 2.235060e-02f, 1.730112e-02f, 1.360286e-02f, 1.084340e-02f, 8.750306e-03f, 7.139216e-03f);
}
__device__ __forceinline__ Spectrum ozone_extinction() {
  return s_set(1.484836e-05f, 8.501668e-05f, 2.646158e-04f, 7.953520e-04f, 1.661103e-03f, 2.510733e-03f, 2.697211e-03f, 1.727741e-03f);
}
#define LB_SKY_MIE_SCATTERING (3.996f * 0.001f)
#define LB_SKY_MIE_EXTINCTION (4.440f * 0.001f)

__device__ __forceinline__ float length3(V3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ V3 normalize3(V3 a) { return a * rsqrtf(dot3(a, a)); }  // math.cuh:174-182

__device__ __forceinline__ V3 sun_pos(const LbSkyDev& S) { return v3(S.sun_pos[0], S.sun_pos[1], S.sun_pos[2]); }
__device__ __forceinline__ V3 moon_pos(const LbSkyDev& S) { return v3(S.moon_pos[0], S.moon_pos[1], S.moon_pos[2]); }

// world_to_sky_transform, sky_utils.cuh:21-31
__device__ __forceinline__ V3 world_to_sky(const LbSkyDev& S, V3 p) {
  V3 r = v3(p.x * 0.001f, p.y * 0.001f + LB_SKY_EARTH_RADIUS, p.z * 0.001f);
  return r + v3(S.geometry_offset[0], S.geometry_offset[1], S.geometry_offset[2]);
}

__device__ __forceinline__ float sky_height(V3 p) { return length3(p) - LB_SKY_EARTH_RADIUS; }

// ---- spheres, math.cuh:620-779 ----
__device__ __forceinline__ float sphere_ray_intersection(V3 ray, V3 origin, V3 p, float r) {
  const V3 diff   = origin - p;
  const float dot = dot3(diff, ray);
  const float r2  = r * r;
  const float c   = dot3(diff, diff) - r2;
  const V3 k      = diff - ray * dot;
  const float d   = r2 - dot3(k, k);
  if (d < 0.0f)
    return FLT_MAX;
  const float sd = sqrtf(d);
  const float q  = -dot - copysignf(sd, dot);
  const float t0 = c / q;
  if (t0 >= 0.0f)
    return t0;
  return (q >= 0.0f) ? q : FLT_MAX;
}
__device__ __forceinline__ bool sphere_ray_hit(V3 ray, V3 origin, V3 p, float r) {
  const V3 diff   = origin - p;
  const float dot = dot3(diff, ray);
  const float r2  = r * r;
  const float c   = dot3(diff, diff) - r2;
  const V3 k      = diff - ray * dot;
  const float d   = r2 - dot3(k, k);
  if (d < 0.0f)
    return false;
  const float sd = sqrtf(d);
  const float q  = -dot - copysignf(sd, dot);
  return (c / q) >= 0.0f;
}
__device__ __forceinline__ float sph_ray_int_p0(V3 ray, V3 origin, float r) {
  const float dot = dot3(origin, ray);
  const float r2  = r * r;
  const V3 k      = origin - ray * dot;
  const float d   = r2 - dot3(k, k);
  if (d < 0.0f)
    return FLT_MAX;
  const float sd = sqrtf(d);
  const float q  = -dot - copysignf(sd, dot);
  const float c  = dot3(origin, origin) - r2;
  const float t0 = c / q;
  if (t0 >= 0.0f)
    return t0;
  return (q >= 0.0f) ? q : FLT_MAX;
}
__device__ __forceinline__ float sph_ray_int_back_p0(V3 ray, V3 origin, float r) {
  const float dot = dot3(origin, ray);
  const float r2  = r * r;
  const V3 k      = origin - ray * dot;
  const float d   = r2 - dot3(k, k);
  if (d < 0.0f)
    return FLT_MAX;
  const float sd = sqrtf(d);
  const float q  = -dot - copysignf(sd, dot);
  const float c  = dot3(origin, origin) - r2;
  if (q >= 0.0f)
    return q;
  const float t0 = c / q;
  return (t0 >= 0.0f) ? t0 : FLT_MAX;
}
__device__ __forceinline__ bool sph_ray_hit_p0(V3 ray, V3 origin, float r) {
  const float dot = dot3(origin, ray);
  const float r2  = r * r;
  const V3 k      = origin - ray * dot;
  const float d   = r2 - dot3(k, k);
  if (d < 0.0f)
    return false;
  const float sd = sqrtf(d);
  const float q  = -dot - copysignf(sd, dot);
  const float c  = dot3(origin, origin) - r2;
  return (c / q) >= 0.0f;
}

// sample_ray_sphere, math.cuh:330-347
__device__ __forceinline__ V3 sample_ray_sphere(float alpha, float beta) {
  if (fabsf(alpha) > 1.0f - FLT_EPSILON)
    return v3(0.0f, 0.0f, copysignf(1.0f, alpha));
  const float a = sqrtf(1.0f - alpha * alpha);
  const float b = 2.0f * LB_SKY_PI * beta;
  return v3(a * cosf(b), a * sinf(b), alpha);
}

// sample_hemisphere_basis, math.cuh:277-299
__device__ __forceinline__ V3 sample_hemisphere_basis(float altitude, float azimuth, V3 basis) {
  const float sign = copysignf(1.0f, basis.z);
  const float a    = -1.0f / (sign + basis.z);
  const float b    = basis.x * basis.y * a;
  const V3 u1      = v3(1.0f + sign * basis.x * basis.x * a, sign * b, -sign * basis.x);
  const V3 u2      = v3(b, sign + basis.y * basis.y * a, -basis.y);
  const float c1   = sinf(altitude) * cosf(azimuth);
  const float c2   = sinf(altitude) * sinf(azimuth);
  const float c3   = cosf(altitude);
  V3 result;
  result.x = c1 * u1.x + c2 * u2.x + c3 * basis.x;
  result.y = c1 * u1.y + c2 * u2.y + c3 * basis.y;
  result.z = c1 * u1.z + c2 * u2.z + c3 * basis.z;
  return normalize3(result);
}

// sample_sphere, math.cuh:1393-1419
__device__ __forceinline__ V3 sample_sphere(V3 p, float r, V3 origin, float2 random, float& area) {
  float r1 = random.x;
  float r2 = random.y;
  V3 dir        = p - origin;
  const float d = length3(dir);
  if (d < r) {
    area = 4.0f * LB_SKY_PI;
    return normalize3(sample_ray_sphere(2.0f * r1 - 1.0f, r2));
  }
  r1  = 0.999f * r1;
  r2  = 0.999f * r2;
  dir = dir * (1.0f / d);
  const float angle = asinf(__saturatef(r / d));
  area              = 2.0f * LB_SKY_PI * angle * angle;
  const float u     = sqrtf(r1) * angle;
  const float v     = 2.0f * LB_SKY_PI * r2;
  return normalize3(sample_hemisphere_basis(u, v, dir));
}

// sample_sphere_solid_angle, math.cuh:1429-1439
__device__ __forceinline__ float sample_sphere_solid_angle(V3 p, float r, V3 origin) {
  const float d = length3(p - origin);
  if (d < r)
    return 2.0f * LB_SKY_PI;
  const float a = asinf(r / d);
  return 2.0f * LB_SKY_PI * a * a;
}

// ---- phase functions, math.cuh:1167-1239 ----
struct JendersieEon {
  float g_hg, g_d, alpha, w_d;
};
__device__ __forceinline__ float henyey_greenstein(float cos_angle, float g) {
  const float g2         = g * g;
  const float denom_term = 1.0f + g2 - 2.0f * g * cos_angle;
  const float pow15      = denom_term * sqrtf(denom_term);
  return (1.0f - g * g) / (4.0f * LB_SKY_PI * pow15);
}
__device__ __forceinline__ float draine(float cos_angle, float g, float alpha) {
  return henyey_greenstein(cos_angle, g) * ((1.0f + alpha * cos_angle * cos_angle) / (1.0f + (alpha / 3.0f) * (1.0f + 2.0f * g * g)));
}
__device__ __forceinline__ JendersieEon jendersie_eon_parameters(float d) {
  JendersieEon p;
  p.g_hg = p.g_d = p.alpha = p.w_d = 0.0f;  // the reference leaves them uninitialised for a NaN diameter
  if (d >= 5.0f && d <= 50.0f) {
    p.g_hg  = expf(-0.0990567f / (d - 1.67154f));
    p.g_d   = expf(-(2.20679f / (d + 3.91029f)) - 0.428934f);
    p.alpha = expf(3.62489f - (8.29288f / (d + 5.52825f)));
    p.w_d   = expf(-(0.599085f / (d - 0.641583f)) - 0.665888f);
  }
  else if (d >= 1.5f && d < 5.0f) {
    p.g_hg  = 0.0604931f * logf(logf(d)) + 0.940256f;
    p.g_d   = 0.500411f - (0.081287f / (-2.0f * logf(d) + tanf(logf(d)) + 1.27551f));
    p.alpha = 7.30354f * logf(d) + 6.31675f;
    p.w_d   = 0.026914f * (logf(d) - cosf(5.68947f * (logf(logf(d)) - 0.0292149f))) + 0.376475f;
  }
  else if (d >= 0.1f && d < 1.5f) {
    p.g_hg = 0.862f - 0.143f * logf(d) * logf(d);
    p.g_d  = 0.379685f
              * cosf(1.19692f * cosf(((logf(d) - 0.238604f) * (logf(d) + 1.00667f)) / (0.507522f - 0.15677f * logf(d))) + 1.37932f * logf(d)
                     + 0.0625835f)
            + 0.344213f;
    p.alpha = 250.0f;
    p.w_d   = 0.146209f * cosf(3.38707f * logf(d) + 2.11193f) + 0.316072f + 0.0778917f * logf(d);
  }
  else if (d < 0.1f) {
    p.g_hg  = 13.8f * d * d;
    p.g_d   = 1.1456f * d * sinf(9.29044f * d);
    p.alpha = 250.0f;
    p.w_d   = 0.252977f - 312.983f * powf(d, 4.3f);
  }
  return p;
}
__device__ __forceinline__ float jendersie_eon_phase(float cos_angle, const JendersieEon& p) {
  const float phase_hg = henyey_greenstein(cos_angle, p.g_hg);
  const float phase_d  = draine(cos_angle, p.g_d, p.alpha);
  return (1.0f - p.w_d) * phase_hg + p.w_d * phase_d;
}

// ---- densities, sky_utils.cuh:85-91, sky.cuh:46-70 ----
__device__ __forceinline__ float rayleigh_phase(float cos_angle) { return 3.0f * (1.0f + cos_angle * cos_angle) / (16.0f * 3.1415926535f); }
__device__ __forceinline__ float rayleigh_density(const LbSkyDev& S, float height) {
  return 2.5f * S.base_density * expf(-height * (1.0f / S.rayleigh_falloff));
}
__device__ __forceinline__ float mie_density(const LbSkyDev& S, float height) {
  const float INSO = expf(-height * (1.0f / S.mie_falloff));
  float WASO       = 0.0f;
  if (height < 2.0f)
    WASO = 1.0f + 0.125f * (2.0f - height);
  else if (height < 3.0f)
    WASO = 3.0f - height;
  WASO *= 60.0f / S.ground_visibility;
  return S.base_density * (INSO + WASO);
}
__device__ __forceinline__ float ozone_density(const LbSkyDev& S, float height) {
  if (!S.ozone_absorption)
    return 0.0f;
  const float min_val = (height > 25.0f) ? 0.0f : 0.1f;
  return S.base_density * fmaxf(min_val, 1.0f - fabsf(height - 25.0f) / S.ozone_layer_thickness);
}

// sky_compute_path, sky.cuh:78-101
__device__ __forceinline__ float2 compute_path(V3 origin, V3 ray, float min_height, float max_height) {
  const float height = length3(origin);
  if (height <= min_height)
    return make_float2(0.0f, -FLT_MAX);
  float distance;
  float start = 0.0f;
  if (height > max_height) {
    const float earth_dist = sph_ray_int_p0(ray, origin, min_height);
    const float atmo_dist  = sph_ray_int_p0(ray, origin, max_height);
    const float atmo_dist2 = sph_ray_int_back_p0(ray, origin, max_height);
    distance               = fminf(earth_dist - atmo_dist, atmo_dist2 - atmo_dist);
    start                  = atmo_dist;
  }
  else {
    const float earth_dist = sph_ray_int_p0(ray, origin, min_height);
    const float atmo_dist  = sph_ray_int_p0(ray, origin, max_height);
    distance               = fminf(earth_dist, atmo_dist);
  }
  return make_float2(start, distance);
}

// [Hil20] LUT parametrisations, sky_utils.cuh:75-83, 278-294
__device__ __forceinline__ float sub_to_unit_uv(float u, float resolution) { return (u - 0.5f / resolution) * (resolution / (resolution - 1.0f)); }
__device__ __forceinline__ float2 transmittance_lut_uv(float height, float zenith_cos_angle) {
  height += LB_SKY_EARTH_RADIUS;
  const float H   = sqrtf(fmaxf(0.0f, LB_SKY_ATMO_RADIUS * LB_SKY_ATMO_RADIUS - LB_SKY_EARTH_RADIUS * LB_SKY_EARTH_RADIUS));
  const float rho = sqrtf(fmaxf(0.0f, height * height - LB_SKY_EARTH_RADIUS * LB_SKY_EARTH_RADIUS));
  const float discriminant = height * height * (zenith_cos_angle * zenith_cos_angle - 1.0f) + LB_SKY_ATMO_RADIUS * LB_SKY_ATMO_RADIUS;
  const float d            = fmaxf(0.0f, (-height * zenith_cos_angle + sqrtf(discriminant)));
  const float d_min = LB_SKY_ATMO_RADIUS - height;
  const float d_max = rho + H;
  return make_float2((d - d_min) / (d_max - d_min), rho / H);
}

// sky_compute_color_from_spectrum, sky_utils.cuh:297-320
__device__ __forceinline__ float3 color_from_spectrum(const Spectrum& radiance) {
  const float r = 0.00640271f * radiance.v[0] + 0.179441f * radiance.v[1] + 0.04852f * radiance.v[2] - 0.43822f * radiance.v[3]
                  - 0.920721f * radiance.v[4] - 0.0226871f * radiance.v[5] + 1.83443f * radiance.v[6] + 2.36265f * radiance.v[7];
  const float g = -0.00550232f * radiance.v[0] - 0.164f * radiance.v[1] - 0.119836f * radiance.v[2] + 0.365423f * radiance.v[3]
                  + 1.28952f * radiance.v[4] + 1.41809f * radiance.v[5] + 0.629138f * radiance.v[6] - 0.0816028f * radiance.v[7];
  const float b = 0.0386558f * radiance.v[0] + 1.21426f * radiance.v[1] + 1.80395f * radiance.v[2] + 0.475181f * radiance.v[3]
                  - 0.0638328f * radiance.v[4] - 0.169502f * radiance.v[5] - 0.114583f * radiance.v[6] - 0.0374822f * radiance.v[7];
  return make_float3(fmaxf(r, 0.0f), fmaxf(g, 0.0f), fmaxf(b, 0.0f));
}

// the medium's coefficients at one height (shared by the LUT integrators and the ray march)
struct Medium {
  Spectrum scattering_rayleigh;
  float scattering_mie;
  Spectrum extinction;
};
__device__ __forceinline__ Medium medium_at(const LbSkyDev& S, float height) {
  const float density_rayleigh = rayleigh_density(S, height) * S.rayleigh_density;
  const float density_mie      = mie_density(S, height) * S.mie_density;
  const float density_ozone    = ozone_density(S, height) * S.ozone_density;
  Medium m;
  m.scattering_rayleigh              = s_scale(rayleigh_scattering(), density_rayleigh);
  m.scattering_mie                   = LB_SKY_MIE_SCATTERING * density_mie;
  const Spectrum extinction_rayleigh = s_scale(rayleigh_scattering(), density_rayleigh);  // SKY_RAYLEIGH_EXTINCTION == SKY_RAYLEIGH_SCATTERING
  const float extinction_mie         = LB_SKY_MIE_EXTINCTION * density_mie;
  const Spectrum extinction_ozone    = s_scale(ozone_extinction(), density_ozone);
  m.extinction                       = s_add(s_add(extinction_rayleigh, s_set1(extinction_mie)), extinction_ozone);
  return m;
}

// sky_get_sun_color, sky_utils.cuh:322-349 (no cloud HDRI)
__device__ __forceinline__ float3 sun_color(const LbSkyDev& S, V3 origin, V3 ray) {
  const float height           = sky_height(origin);
  const float zenith_cos_angle = dot3(normalize3(origin), ray);
  const float2 uv              = transmittance_lut_uv(height, zenith_cos_angle);
  const float4 low             = tex2D<float4>(S.tm_low, uv.x, uv.y);
  const float4 high            = tex2D<float4>(S.tm_high, uv.x, uv.y);
  const Spectrum extinction_sun = s_mul(s_ident(), s_merge(low, high));
  const Spectrum radiance       = s_mul(extinction_sun, s_scale(sun_radiance(), S.sun_strength));
  return color_from_spectrum(radiance);
}

// sky_compute_atmosphere without cloud shadows, sky.cuh:338-502; random_offset = random_1D(RANDOM_TARGET_SKY_STEP_OFFSET)
// transmittance_out (aerial perspective only): multiplied by the transmittance of the marched segment (sky.cuh:499)
__device__ inline Spectrum compute_atmosphere(const LbSkyDev& S, V3 origin, V3 ray, float limit, bool celestials, int steps, float random_offset,
                                              Spectrum* transmittance_out = nullptr) {
  Spectrum result = s_set1(0.0f);
  const float2 path    = compute_path(origin, ray, LB_SKY_EARTH_RADIUS, LB_SKY_ATMO_RADIUS);
  const float start    = path.x;
  const float distance = fminf(path.y, limit - start);
  Spectrum transmittance = s_ident();
  const V3 sun           = sun_pos(S);

  if (distance > 0.0f) {
    float reach = start;
    float step_size;
    const float light_angle = sample_sphere_solid_angle(sun, LB_SKY_SUN_RADIUS, origin);
    const JendersieEon mie  = jendersie_eon_parameters(S.mie_diameter);
#pragma unroll 1
    for (int i = 0; i < steps; i++) {
      const float new_reach = start + distance * (i + random_offset) / steps;
      step_size             = new_reach - reach;
      reach                 = new_reach;

      const V3 pos       = origin + ray * reach;
      const float height = sky_height(pos);

      const V3 ray_scatter         = normalize3(sun - pos);
      const float cos_angle        = dot3(ray, ray_scatter);
      const float zenith_cos_angle = dot3(normalize3(pos), ray_scatter);
      const float phase_rayleigh   = rayleigh_phase(cos_angle);
      const float phase_mie        = jendersie_eon_phase(cos_angle, mie);
      const float shadow           = sph_ray_hit_p0(ray_scatter, pos, LB_SKY_EARTH_RADIUS) ? 0.0f : 1.0f;

      const float2 tm_uv            = transmittance_lut_uv(height, zenith_cos_angle);
      const Spectrum extinction_sun = s_merge(tex2D<float4>(S.tm_low, tm_uv.x, tm_uv.y), tex2D<float4>(S.tm_high, tm_uv.x, tm_uv.y));

      const Medium m            = medium_at(S, height);
      const Spectrum scattering = s_add(m.scattering_rayleigh, s_set1(m.scattering_mie));
      const Spectrum phase_times_scattering = s_add(s_scale(m.scattering_rayleigh, phase_rayleigh), s_set1(m.scattering_mie * phase_mie));
      const Spectrum ss_radiance            = s_scale(s_mul(extinction_sun, phase_times_scattering), shadow * light_angle);

      const float ms_u = zenith_cos_angle * 0.5f + 0.5f, ms_v = height / LB_SKY_ATMO_HEIGHT;
      const Spectrum ms_tex      = s_merge(tex2D<float4>(S.ms_low, ms_u, ms_v), tex2D<float4>(S.ms_high, ms_u, ms_v));
      const Spectrum ms_radiance = s_mul(ms_tex, scattering);
      const Spectrum Ssum        = s_add(ss_radiance, ms_radiance);

      const Spectrum step_transmittance = s_exp(s_scale(m.extinction, -step_size));
      const Spectrum Sint               = s_mul(s_sub(Ssum, s_mul(Ssum, step_transmittance)), s_inv(m.extinction));
      result                            = s_add(result, s_mul(Sint, transmittance));
      transmittance                     = s_mul(transmittance, step_transmittance);
    }
    result = s_mul(result, s_scale(sun_radiance(), S.sun_strength));
  }

  if (celestials) {
    const V3 moon         = moon_pos(S);
    const float sun_hit   = sphere_ray_intersection(ray, origin, sun, LB_SKY_SUN_RADIUS);
    const float earth_hit = sph_ray_int_p0(ray, origin, LB_SKY_EARTH_RADIUS);
    const float moon_hit  = sphere_ray_intersection(ray, origin, moon, LB_SKY_MOON_RADIUS);
    if (earth_hit > sun_hit && moon_hit > sun_hit)
      result = s_add(result, s_mul(transmittance, s_scale(sun_radiance(), S.sun_strength)));
    else if (earth_hit > moon_hit) {  // the moon's surface lit by the sun, sky.cuh:440-475
      const V3 mp         = origin + ray * moon_hit;
      const V3 bounce_ray = normalize3(sun - mp);
      if (!sphere_ray_hit(bounce_ray, mp, v3(0.0f, 0.0f, 0.0f), LB_SKY_EARTH_RADIUS)) {
        V3 normal         = normalize3(mp - moon);
        const float tex_u = 0.5f + S.moon_tex_offset + atan2f(normal.z, normal.x) * (1.0f / (2.0f * LB_SKY_PI));
        const float tex_v = 0.5f + asinf(normal.y) * (1.0f / LB_SKY_PI);
        // texture_load with its default arguments (texture_utils.cuh:13-45): flipped v, gamma applied, (0, 0, 0, 0) when the texture is absent
        const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const float4 nv   = S.moon_normal.handle ? lb_texture_fetch(S.moon_normal, make_float2(tex_u, tex_v), true, true) : zero;
        // create_basis + transform_vec3, math.cuh:301-321, 445-453
        const float sign = copysignf(1.0f, normal.z);
        const float a    = -1.0f / (sign + normal.z);
        const float b    = normal.x * normal.y * a;
        const V3 u1      = v3(1.0f + sign * normal.x * normal.x * a, sign * b, -sign * normal.x);
        const V3 u2      = v3(b, sign + normal.y * normal.y * a, -normal.y);
        const V3 mn      = v3(nv.x * 2.0f - 1.0f, nv.y * 2.0f - 1.0f, nv.z * 2.0f - 1.0f);
        normal = normalize3(v3(u1.x * mn.x + u2.x * mn.y + normal.x * mn.z, u1.y * mn.x + u2.y * mn.y + normal.y * mn.z,
                               u1.z * mn.x + u2.z * mn.y + normal.z * mn.z));
        const float NdotL = dot3(normal, bounce_ray);
        if (NdotL > 0.0f) {
          const float albedo      = S.moon_albedo.handle ? lb_texture_fetch(S.moon_albedo, make_float2(tex_u, tex_v), true, true).x : 0.0f;
          const float light_angle = sample_sphere_solid_angle(sun, LB_SKY_SUN_RADIUS, mp);
          const float weight      = albedo * S.sun_strength * NdotL * light_angle / (2.0f * LB_SKY_PI);
          const Spectrum flux     = s_set(1.7f, 1.8f, 2.0f, 1.9f, 1.87f, 1.7f, 1.65f, 1.55f);  // SKY_MOON_SOLAR_FLUX
          result                  = s_add(result, s_mul(transmittance, s_mul(flux, s_scale(sun_radiance(), weight))));
        }
      }
    }
    if (S.has_stars && sun_hit == FLT_MAX && earth_hit == FLT_MAX && moon_hit == FLT_MAX) {
      const float ray_altitude = asinf(ray.y);
      const float ray_azimuth  = atan2f(-ray.z, -ray.x) + LB_SKY_PI;
      const uint32_t x    = (uint32_t) (ray_azimuth * 10.0f);
      const uint32_t y    = (uint32_t) ((ray_altitude + LB_SKY_PI * 0.5f) * 10.0f);
      const uint32_t grid = x + y * LB_STARS_GRID_X;
      const uint32_t a    = __ldg(S.stars_offsets + grid);
      const uint32_t b    = __ldg(S.stars_offsets + grid + 1);
      for (uint32_t i = a; i < b; i++) {
        const float4 star = __ldg(S.stars + i);  // altitude, azimuth, radius, intensity
        const V3 star_pos = v3(cosf(star.y) * cosf(star.x), sinf(star.x), sinf(star.y) * cosf(star.x));
        if (sphere_ray_hit(ray, v3(0.0f, 0.0f, 0.0f), star_pos, star.z))
          result = s_add(result, s_scale(transmittance, star.w * S.stars_intensity));
      }
    }
  }
  if (transmittance_out)
    *transmittance_out = s_mul(*transmittance_out, transmittance);
  return result;
}

// sky_color_main (DEFAULT mode), sky.cuh:567-576: `state` decides whether the sun disc / stars are visible
__device__ __forceinline__ float3 sky_color(const LbSkyDev& S, V3 origin_world, V3 ray, bool include_sun, float random_offset) {
  const V3 sky_origin = world_to_sky(S, origin_world);
  return color_from_spectrum(compute_atmosphere(S, sky_origin, ray, FLT_MAX, include_sun, (int) S.steps, random_offset));
}

// sky_color_main (HDRI mode), sky.cuh:577-596 + sky_hdri_sample, sky_utils.cuh:49-63: the baked table plus the sun's disc
__device__ __forceinline__ float3 sky_color_hdri(const LbSkyDev& S, V3 origin_world, V3 ray, bool include_sun) {
  const float theta = atan2f(ray.z, ray.x);
  const float phi   = asinf(ray.y);
  const float u     = (theta + LB_SKY_PI) / (2.0f * LB_SKY_PI);
  const float v     = 1.0f - ((phi + 0.5f * LB_SKY_PI) / LB_SKY_PI);
  const float4 h    = tex2DLod<float4>(S.hdri, u, v, 0.0f);
  float3 sky        = make_float3(h.x, h.y, h.z);
  if (include_sun) {
    const V3 sky_origin = world_to_sky(S, origin_world);
    if (sphere_ray_hit(ray, sky_origin, sun_pos(S), LB_SKY_SUN_RADIUS) && !sph_ray_hit_p0(ray, sky_origin, LB_SKY_EARTH_RADIUS)) {
      const float3 sc = sun_color(S, sky_origin, ray);
      sky.x += sc.x, sky.y += sc.y, sky.z += sc.z;
    }
  }
  return sky;
}

}  // namespace lbsky
