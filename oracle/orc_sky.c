/* orc_sky.c - CPU restatement of the reference's procedural atmosphere (LUMINARY_SKY_MODE_DEFAULT). TEST INFRASTRUCTURE ONLY.
 *
 *   sky LUTs              sky_compute_transmittance_lut / sky_compute_multiscattering_lut, cuda/sky.cuh:109-330 (launch: device_sky.c:80-117)
 *   miss shading          sky_color_main -> sky_get_color -> sky_compute_atmosphere, cuda/sky.cuh:338-515, 567-601
 *   sun radiance          sky_get_sun_color, cuda/sky_utils.cuh:322-349
 *   helpers               cuda/sky_utils.cuh:9-320, cuda/math.cuh:277-347 (sampling), :620-779 (spheres), :1167-1239 (phase functions),
 *                         :1393-1439 (sample_sphere)
 *   host side             device_struct_sky_convert (device_structs.c:107-172: sun / moon positions in double), sky_get_default (sky.c:6-42),
 *                         _sky_stars_generate (device_sky.c:484-546: glibc rand() catalogue, 64 x 32 grid)
 * Plain IEEE single precision with libm (the reference's kernels run under --use_fast_math: rsqrtf, __expf, approximate division
 * are the documented fp32 tolerance of the sky parity tests). The LUT textures are fetched through orc_texture_fetch, the measured
 * model of the B200 texture unit (float4, linear, clamp, normalised coordinates - texture defaults + device_sky.c:46-60).
 * Not restated, like in the product: clouds, aerial perspective, the HDRI mode, the moon's surface textures (black disc).
 * Parity pinning: tests/test_sky_gpu.py compares LUTs, sky colours and the sun's NEE task with the reference's own kernels
 * (oracle/_ref/librefdev.so). */
#include <float.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "orc_internal.h"

#define SKY_EARTH_RADIUS 6371.0f
#define SKY_SUN_RADIUS 696340.0f
#define SKY_SUN_DISTANCE 149597870.0f
#define SKY_MOON_RADIUS 1737.4f
#define SKY_MOON_DISTANCE 384399.0f
#define SKY_ATMO_HEIGHT 100.0f
#define SKY_ATMO_RADIUS (SKY_ATMO_HEIGHT + SKY_EARTH_RADIUS)
#define SKY_MS_TEX_SIZE 32
#define SKY_TM_TEX_WIDTH 256
#define SKY_TM_TEX_HEIGHT 64
#define SKY_MS_BASE 16
#define SKY_MS_ITER (SKY_MS_BASE * SKY_MS_BASE)
#define SKY_HEIGHT_OFFSET 0.0005f
#define STARS_GRID_X 64
#define STARS_GRID_Y 32
#define PI_F 3.141592653589f
#define SKY_MIE_SCATTERING (3.996f * 0.001f)
#define SKY_MIE_EXTINCTION (4.440f * 0.001f)

typedef struct {
  float v[8];
} Spectrum;

static const Spectrum S_IDENT        = {{8.4205e-03f, 2.6449e-01f, 4.0273e-01f, 1.6624e-01f, 2.4324e-01f, 3.5849e-01f, 3.6342e-01f, 2.4177e-01f}};
static const Spectrum S_SUN_RADIANCE = {{2.463170e+04f, 2.888721e+04f, 2.795153e+04f, 2.629836e+04f, 2.667237e+04f, 2.638737e+04f, 2.490630e+04f,
                                         2.338930e+04f}};
static const Spectrum S_RAYLEIGH     = {{3.945800e-02f, 2.939289e-02f, 2.235060e-02f, 1.730112e-02f, 1.360286e-02f, 1.084340e-02f, 8.750306e-03f,
                                         7.139216e-03f}};
static const Spectrum S_OZONE        = {{1.484836e-05f, 8.501668e-05f, 2.646158e-04f, 7.953520e-04f, 1.661103e-03f, 2.510733e-03f, 2.697211e-03f,
                                         1.727741e-03f}};

#define S_OP(expr)              \
  Spectrum r;                   \
  for (int i = 0; i < 8; i++)   \
    r.v[i] = (expr);            \
  return r;
static inline Spectrum s_set1(float a) { S_OP(a) }
static inline Spectrum s_add(Spectrum a, Spectrum b) { S_OP(a.v[i] + b.v[i]) }
static inline Spectrum s_sub(Spectrum a, Spectrum b) { S_OP(a.v[i] - b.v[i]) }
static inline Spectrum s_mul(Spectrum a, Spectrum b) { S_OP(a.v[i] * b.v[i]) }
static inline Spectrum s_scale(Spectrum a, float b) { S_OP(a.v[i] * b) }
static inline Spectrum s_inv(Spectrum a) { S_OP(1.0f / a.v[i]) }
static inline Spectrum s_exp(Spectrum a) { S_OP(expf(a.v[i])) }
static inline Spectrum s_merge(const float low[4], const float high[4]) {
  Spectrum r = {{low[0], low[1], low[2], low[3], high[0], high[1], high[2], high[3]}};
  return r;
}

/* ------------------------------------------------------------------ */
/* host side                                                            */
/* ------------------------------------------------------------------ */
void orc_sky_params_default(OrcSkyParams* p) { /* sky.c:6-42 */
  memset(p, 0, sizeof(*p));
  p->geometry_offset        = v_get(0.0f, 0.1f, 0.0f);
  p->altitude               = 0.5f;
  p->azimuth                = 3.141f;
  p->moon_altitude          = -0.5f;
  p->moon_azimuth           = 0.0f;
  p->moon_tex_offset        = 0.0f;
  p->sun_strength           = 1.0f;
  p->base_density           = 1.0f;
  p->rayleigh_density       = 1.0f;
  p->mie_density            = 1.0f;
  p->ozone_density          = 1.0f;
  p->ground_visibility      = 60.0f;
  p->mie_diameter           = 2.0f;
  p->ozone_layer_thickness  = 15.0f;
  p->rayleigh_falloff       = 8.0f;
  p->mie_falloff            = 1.7f;
  p->multiscattering_factor = 1.0f;
  p->steps                  = 40;
  p->ozone_absorption       = 1;
  p->stars_seed             = 0;
  p->stars_count            = 10000;
  p->stars_intensity        = 1.0f;
  p->aerial_perspective     = 0;
}

/* device_struct_sky_convert, device_structs.c:132-170 */
static OrcVec3 celestial_position(float azimuth, float altitude, double distance, OrcVec3 offset) {
  double x = cos(azimuth) * cos(altitude);
  double y = sin(altitude);
  double z = sin(azimuth) * cos(altitude);
  const double scale = 1.0 / (sqrt(x * x + y * y + z * z));
  x *= scale * distance;
  y *= scale * distance;
  z *= scale * distance;
  y -= SKY_EARTH_RADIUS;
  x -= offset.x;
  y -= offset.y;
  z -= offset.z;
  return v_get((float) x, (float) y, (float) z);
}

/* _sky_stars_generate, device_sky.c:484-546 */
static float stars_random_float(void) { return (float) (((double) rand()) / RAND_MAX); }

static void stars_generate(OrcSky* sky, uint32_t count, uint32_t seed) {
  free(sky->stars);
  sky->stars       = NULL;
  sky->stars_count = count;
  sky->has_stars   = 0;
  if (count == 0)
    return;
  srand(seed);
  float* buffer    = (float*) malloc(sizeof(float) * 4 * count);
  uint32_t* counts = (uint32_t*) calloc(STARS_GRID_X * STARS_GRID_Y, sizeof(uint32_t));
  for (uint32_t i = 0; i < count; i++) {
    /* designated initialisers evaluate in order of appearance with gcc: altitude, azimuth, radius, intensity */
    const float altitude  = -PI_F * 0.5f + PI_F * (1.0f - sqrtf(stars_random_float()));
    const float azimuth   = 2.0f * PI_F * stars_random_float();
    const float radius    = 0.0001f + 0.0004f * (1.0f - sqrtf(stars_random_float()));
    const float intensity = 0.0001f + 0.0015f * (0.1f + 0.9f * (1.0f - sqrtf(stars_random_float())));
    const uint32_t x      = (uint32_t) (azimuth * 10.0f);
    const uint32_t y      = (uint32_t) ((altitude + PI_F * 0.5f) * 10.0f);
    if (x < STARS_GRID_X && y < STARS_GRID_Y)
      counts[x + y * STARS_GRID_X]++;
    buffer[4 * i + 0] = altitude, buffer[4 * i + 1] = azimuth, buffer[4 * i + 2] = radius, buffer[4 * i + 3] = intensity;
  }
  uint32_t offset = 0;
  for (uint32_t i = 0; i < STARS_GRID_X * STARS_GRID_Y; i++) {
    sky->stars_offsets[i] = offset;
    offset += counts[i];
    counts[i] = 0;
  }
  sky->stars_offsets[STARS_GRID_X * STARS_GRID_Y] = offset;
  sky->stars = (float*) calloc((size_t) 4 * count, sizeof(float));
  for (uint32_t i = 0; i < count; i++) {
    const uint32_t x = (uint32_t) (buffer[4 * i + 1] * 10.0f);
    const uint32_t y = (uint32_t) ((buffer[4 * i + 0] + PI_F * 0.5f) * 10.0f);
    if (x >= STARS_GRID_X || y >= STARS_GRID_Y)
      continue; /* the reference raises "Star generation exception." */
    const uint32_t p = x + y * STARS_GRID_X;
    memcpy(sky->stars + 4 * (size_t) (sky->stars_offsets[p] + counts[p]++), buffer + 4 * i, sizeof(float) * 4);
  }
  free(buffer);
  free(counts);
  sky->has_stars = 1;
}

/* ------------------------------------------------------------------ */
/* device side                                                          */
/* ------------------------------------------------------------------ */
static inline float sky_height(OrcVec3 p) { return v_len(p) - SKY_EARTH_RADIUS; }

OrcVec3 orc_world_to_sky(const OrcSky* sky, OrcVec3 p) { /* sky_utils.cuh:21-31 */
  OrcVec3 r = v_get(p.x * 0.001f, p.y * 0.001f + SKY_EARTH_RADIUS, p.z * 0.001f);
  return v_add(r, sky->p.geometry_offset);
}

static float sphere_ray_intersection(OrcVec3 ray, OrcVec3 origin, OrcVec3 p, float r) { /* math.cuh:620-641 */
  const OrcVec3 diff = v_sub(origin, p);
  const float dot    = v_dot(diff, ray);
  const float r2     = r * r;
  const float c      = v_dot(diff, diff) - r2;
  const OrcVec3 k    = v_sub(diff, v_scale(ray, dot));
  const float d      = r2 - v_dot(k, k);
  if (d < 0.0f)
    return ORC_FLT_MAX;
  const float sd = sqrtf(d);
  const float q  = -dot - copysignf(sd, dot);
  const float t0 = c / q;
  if (t0 >= 0.0f)
    return t0;
  return (q >= 0.0f) ? q : ORC_FLT_MAX;
}

bool orc_sphere_ray_hit(OrcVec3 ray, OrcVec3 origin, OrcVec3 p, float r) { /* math.cuh:679-696 */
  const OrcVec3 diff = v_sub(origin, p);
  const float dot    = v_dot(diff, ray);
  const float r2     = r * r;
  const float c      = v_dot(diff, diff) - r2;
  const OrcVec3 k    = v_sub(diff, v_scale(ray, dot));
  const float d      = r2 - v_dot(k, k);
  if (d < 0.0f)
    return false;
  const float sd = sqrtf(d);
  const float q  = -dot - copysignf(sd, dot);
  return (c / q) >= 0.0f;
}

static float sph_ray_int_p0(OrcVec3 ray, OrcVec3 origin, float r) { /* math.cuh:650-669 */
  const float dot = v_dot(origin, ray);
  const float r2  = r * r;
  const OrcVec3 k = v_sub(origin, v_scale(ray, dot));
  const float d   = r2 - v_dot(k, k);
  if (d < 0.0f)
    return ORC_FLT_MAX;
  const float sd = sqrtf(d);
  const float q  = -dot - copysignf(sd, dot);
  const float c  = v_dot(origin, origin) - r2;
  const float t0 = c / q;
  if (t0 >= 0.0f)
    return t0;
  return (q >= 0.0f) ? q : ORC_FLT_MAX;
}

static float sph_ray_int_back_p0(OrcVec3 ray, OrcVec3 origin, float r) { /* math.cuh:760-779 */
  const float dot = v_dot(origin, ray);
  const float r2  = r * r;
  const OrcVec3 k = v_sub(origin, v_scale(ray, dot));
  const float d   = r2 - v_dot(k, k);
  if (d < 0.0f)
    return ORC_FLT_MAX;
  const float sd = sqrtf(d);
  const float q  = -dot - copysignf(sd, dot);
  const float c  = v_dot(origin, origin) - r2;
  if (q >= 0.0f)
    return q;
  const float t0 = c / q;
  return (t0 >= 0.0f) ? t0 : ORC_FLT_MAX;
}

bool orc_sph_ray_hit_p0(OrcVec3 ray, OrcVec3 origin, float r) { /* math.cuh:705-720 */
  const float dot = v_dot(origin, ray);
  const float r2  = r * r;
  const OrcVec3 k = v_sub(origin, v_scale(ray, dot));
  const float d   = r2 - v_dot(k, k);
  if (d < 0.0f)
    return false;
  const float sd = sqrtf(d);
  const float q  = -dot - copysignf(sd, dot);
  const float c  = v_dot(origin, origin) - r2;
  return (c / q) >= 0.0f;
}

static OrcVec3 sample_ray_sphere(float alpha, float beta) { /* math.cuh:330-347 */
  if (fabsf(alpha) > 1.0f - FLT_EPSILON)
    return v_get(0.0f, 0.0f, copysignf(1.0f, alpha));
  const float a = sqrtf(1.0f - alpha * alpha);
  const float b = 2.0f * PI_F * beta;
  return v_get(a * cosf(b), a * sinf(b), alpha);
}

static OrcVec3 sample_hemisphere_basis(float altitude, float azimuth, OrcVec3 basis) { /* math.cuh:277-299 */
  const float sign = copysignf(1.0f, basis.z);
  const float a    = -1.0f / (sign + basis.z);
  const float b    = basis.x * basis.y * a;
  const OrcVec3 u1 = v_get(1.0f + sign * basis.x * basis.x * a, sign * b, -sign * basis.x);
  const OrcVec3 u2 = v_get(b, sign + basis.y * basis.y * a, -basis.y);
  const float c1   = sinf(altitude) * cosf(azimuth);
  const float c2   = sinf(altitude) * sinf(azimuth);
  const float c3   = cosf(altitude);
  OrcVec3 result;
  result.x = c1 * u1.x + c2 * u2.x + c3 * basis.x;
  result.y = c1 * u1.y + c2 * u2.y + c3 * basis.y;
  result.z = c1 * u1.z + c2 * u2.z + c3 * basis.z;
  return v_normalize(result);
}

OrcVec3 orc_sample_sphere(OrcVec3 p, float r, OrcVec3 origin, OrcFloat2 random, float* area) { /* math.cuh:1393-1419 */
  float r1 = random.x;
  float r2 = random.y;
  OrcVec3 dir   = v_sub(p, origin);
  const float d = v_len(dir);
  if (d < r) {
    *area = 4.0f * PI_F;
    return v_normalize(sample_ray_sphere(2.0f * r1 - 1.0f, r2));
  }
  r1  = 0.999f * r1;
  r2  = 0.999f * r2;
  dir = v_scale(dir, 1.0f / d);
  const float angle = asinf(orc_saturate(r / d));
  *area             = 2.0f * PI_F * angle * angle;
  const float u     = sqrtf(r1) * angle;
  const float v     = 2.0f * PI_F * r2;
  return v_normalize(sample_hemisphere_basis(u, v, dir));
}

static float sample_sphere_solid_angle(OrcVec3 p, float r, OrcVec3 origin) { /* math.cuh:1429-1439 */
  const float d = v_len(v_sub(p, origin));
  if (d < r)
    return 2.0f * PI_F;
  const float a = asinf(r / d);
  return 2.0f * PI_F * a * a;
}

typedef struct {
  float g_hg, g_d, alpha, w_d;
} JendersieEon;

static float henyey_greenstein(float cos_angle, float g) { /* math.cuh:1167-1173 */
  const float g2         = g * g;
  const float denom_term = 1.0f + g2 - 2.0f * g * cos_angle;
  const float pow15      = denom_term * sqrtf(denom_term);
  return (1.0f - g * g) / (4.0f * PI_F * pow15);
}
static float draine(float cos_angle, float g, float alpha) { /* math.cuh:1175-1178 */
  return henyey_greenstein(cos_angle, g) * ((1.0f + alpha * cos_angle * cos_angle) / (1.0f + (alpha / 3.0f) * (1.0f + 2.0f * g * g)));
}
static JendersieEon jendersie_eon_parameters(float d) { /* math.cuh:1189-1223 */
  JendersieEon p = {0.0f, 0.0f, 0.0f, 0.0f};
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
static float jendersie_eon_phase(float cos_angle, const JendersieEon* p) { /* math.cuh:1234-1239 */
  const float phase_hg = henyey_greenstein(cos_angle, p->g_hg);
  const float phase_d  = draine(cos_angle, p->g_d, p->alpha);
  return (1.0f - p->w_d) * phase_hg + p->w_d * phase_d;
}

static float rayleigh_phase(float cos_angle) { return 3.0f * (1.0f + cos_angle * cos_angle) / (16.0f * 3.1415926535f); }
static float rayleigh_density(const OrcSkyParams* S, float height) { return 2.5f * S->base_density * expf(-height * (1.0f / S->rayleigh_falloff)); }
static float mie_density(const OrcSkyParams* S, float height) { /* sky.cuh:47-62 */
  const float INSO = expf(-height * (1.0f / S->mie_falloff));
  float WASO       = 0.0f;
  if (height < 2.0f)
    WASO = 1.0f + 0.125f * (2.0f - height);
  else if (height < 3.0f)
    WASO = 3.0f - height;
  WASO *= 60.0f / S->ground_visibility;
  return S->base_density * (INSO + WASO);
}
static float ozone_density(const OrcSkyParams* S, float height) { /* sky.cuh:64-70 */
  if (!S->ozone_absorption)
    return 0.0f;
  const float min_val = (height > 25.0f) ? 0.0f : 0.1f;
  return S->base_density * fmaxf(min_val, 1.0f - fabsf(height - 25.0f) / S->ozone_layer_thickness);
}

typedef struct {
  Spectrum scattering_rayleigh;
  float scattering_mie;
  Spectrum extinction;
} Medium;
static Medium medium_at(const OrcSkyParams* S, float height) {
  const float density_rayleigh = rayleigh_density(S, height) * S->rayleigh_density;
  const float density_mie      = mie_density(S, height) * S->mie_density;
  const float density_ozone    = ozone_density(S, height) * S->ozone_density;
  Medium m;
  m.scattering_rayleigh              = s_scale(S_RAYLEIGH, density_rayleigh);
  m.scattering_mie                   = SKY_MIE_SCATTERING * density_mie;
  const Spectrum extinction_rayleigh = s_scale(S_RAYLEIGH, density_rayleigh);
  const float extinction_mie         = SKY_MIE_EXTINCTION * density_mie;
  const Spectrum extinction_ozone    = s_scale(S_OZONE, density_ozone);
  m.extinction                       = s_add(s_add(extinction_rayleigh, s_set1(extinction_mie)), extinction_ozone);
  return m;
}

static void compute_path(OrcVec3 origin, OrcVec3 ray, float min_height, float max_height, float* start_out, float* distance_out) { /* sky.cuh:78-101 */
  const float height = v_len(origin);
  if (height <= min_height) {
    *start_out = 0.0f, *distance_out = -ORC_FLT_MAX;
    return;
  }
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
  *start_out = start, *distance_out = distance;
}

static float sub_to_unit_uv(float u, float resolution) { return (u - 0.5f / resolution) * (resolution / (resolution - 1.0f)); }

static void transmittance_lut_uv(float height, float zenith_cos_angle, float* u, float* v) { /* sky_utils.cuh:279-294 */
  height += SKY_EARTH_RADIUS;
  const float H   = sqrtf(fmaxf(0.0f, SKY_ATMO_RADIUS * SKY_ATMO_RADIUS - SKY_EARTH_RADIUS * SKY_EARTH_RADIUS));
  const float rho = sqrtf(fmaxf(0.0f, height * height - SKY_EARTH_RADIUS * SKY_EARTH_RADIUS));
  const float discriminant = height * height * (zenith_cos_angle * zenith_cos_angle - 1.0f) + SKY_ATMO_RADIUS * SKY_ATMO_RADIUS;
  const float d            = fmaxf(0.0f, (-height * zenith_cos_angle + sqrtf(discriminant)));
  const float d_min = SKY_ATMO_RADIUS - height;
  const float d_max = rho + H;
  *u = (d - d_min) / (d_max - d_min);
  *v = rho / H;
}

static OrcTexture lut_texture(const float* data, uint32_t width, uint32_t height) {
  OrcTexture t;
  memset(&t, 0, sizeof(t));
  t.width = width, t.height = height, t.pitch = width * 16u, t.type = ORC_TEX_FP32, t.num_components = 4;
  t.wrap_u = t.wrap_v = 1; /* clamp */
  t.filter = 1;            /* linear */
  t.gamma  = 1.0f;
  t.data   = data;
  return t;
}

static Spectrum fetch_transmittance(const OrcSky* sky, float height, float zenith_cos_angle) {
  float u, v, low[4], high[4];
  transmittance_lut_uv(height, zenith_cos_angle, &u, &v);
  const OrcTexture tl = lut_texture(sky->tm_low, SKY_TM_TEX_WIDTH, SKY_TM_TEX_HEIGHT), th = lut_texture(sky->tm_high, SKY_TM_TEX_WIDTH, SKY_TM_TEX_HEIGHT);
  orc_texture_fetch(&tl, u, v, low);
  orc_texture_fetch(&th, u, v, high);
  return s_merge(low, high);
}

static OrcRGB color_from_spectrum(Spectrum radiance) { /* sky_utils.cuh:297-320 */
  const float r = 0.00640271f * radiance.v[0] + 0.179441f * radiance.v[1] + 0.04852f * radiance.v[2] - 0.43822f * radiance.v[3]
                  - 0.920721f * radiance.v[4] - 0.0226871f * radiance.v[5] + 1.83443f * radiance.v[6] + 2.36265f * radiance.v[7];
  const float g = -0.00550232f * radiance.v[0] - 0.164f * radiance.v[1] - 0.119836f * radiance.v[2] + 0.365423f * radiance.v[3]
                  + 1.28952f * radiance.v[4] + 1.41809f * radiance.v[5] + 0.629138f * radiance.v[6] - 0.0816028f * radiance.v[7];
  const float b = 0.0386558f * radiance.v[0] + 1.21426f * radiance.v[1] + 1.80395f * radiance.v[2] + 0.475181f * radiance.v[3]
                  - 0.0638328f * radiance.v[4] - 0.169502f * radiance.v[5] - 0.114583f * radiance.v[6] - 0.0374822f * radiance.v[7];
  return c_get(fmaxf(r, 0.0f), fmaxf(g, 0.0f), fmaxf(b, 0.0f));
}

OrcRGB orc_sky_sun_color(const OrcSky* sky, OrcVec3 origin, OrcVec3 ray) { /* sky_get_sun_color, sky_utils.cuh:322-349 */
  const float height            = sky_height(origin);
  const float zenith_cos_angle  = v_dot(v_normalize(origin), ray);
  const Spectrum extinction_sun = s_mul(S_IDENT, fetch_transmittance(sky, height, zenith_cos_angle));
  const Spectrum radiance       = s_mul(extinction_sun, s_scale(S_SUN_RADIANCE, sky->p.sun_strength));
  return color_from_spectrum(radiance);
}

/* ---- LUTs ---- */
static Spectrum transmittance_optical_depth(const OrcSkyParams* S, float r, float mu) { /* sky.cuh:110-141 */
  const int steps       = 2500;
  const float disc      = r * r * (mu * mu - 1.0f) + SKY_ATMO_RADIUS * SKY_ATMO_RADIUS;
  const float dist      = fmaxf(-r * mu + sqrtf(fmaxf(0.0f, disc)), 0.0f);
  const float step_size = dist / steps;
  Spectrum depth        = s_set1(0.0f);
  for (int i = 0; i <= steps; i++) {
    const float reach  = i * step_size;
    const float height = sqrtf(reach * reach + 2.0f * r * mu * reach + r * r) - SKY_EARTH_RADIUS;
    const Medium m     = medium_at(S, height);
    const float w      = (i == 0 || i == steps) ? 0.5f : 1.0f;
    depth              = s_add(depth, s_scale(m.extinction, w * step_size));
  }
  return depth;
}

static void build_transmittance_lut(OrcSky* sky) { /* sky.cuh:144-178 */
#pragma omp parallel for schedule(dynamic, 64)
  for (int id = 0; id < SKY_TM_TEX_WIDTH * SKY_TM_TEX_HEIGHT; id++) {
    const int y = id / SKY_TM_TEX_WIDTH;
    const int x = id - y * SKY_TM_TEX_WIDTH;
    float fx    = ((float) x + 0.5f) / SKY_TM_TEX_WIDTH;
    float fy    = ((float) y + 0.5f) / SKY_TM_TEX_HEIGHT;
    fx          = sub_to_unit_uv(fx, SKY_TM_TEX_WIDTH);
    fy          = sub_to_unit_uv(fy, SKY_TM_TEX_HEIGHT);
    const float H   = sqrtf(SKY_ATMO_RADIUS * SKY_ATMO_RADIUS - SKY_EARTH_RADIUS * SKY_EARTH_RADIUS);
    const float rho = H * fy;
    const float r   = sqrtf(rho * rho + SKY_EARTH_RADIUS * SKY_EARTH_RADIUS);
    const float d_min = SKY_ATMO_RADIUS - r;
    const float d_max = rho + H;
    const float d     = d_min + fx * (d_max - d_min);
    float mu = (d == 0.0f) ? 1.0f : (H * H - rho * rho - d * d) / (2.0f * r * d);
    mu       = fminf(1.0f, fmaxf(-1.0f, mu));
    const Spectrum t = s_exp(s_scale(transmittance_optical_depth(&sky->p, r, mu), -1.0f));
    memcpy(sky->tm_low + 4 * (size_t) id, t.v, 16);
    memcpy(sky->tm_high + 4 * (size_t) id, t.v + 4, 16);
  }
}

static void multiscattering_integration(const OrcSky* sky, OrcVec3 origin, OrcVec3 ray, OrcVec3 sun, Spectrum* L, Spectrum* as1) { /* sky.cuh:186-272 */
  *L   = s_set1(0.0f);
  *as1 = s_set1(0.0f);
  float start, distance;
  compute_path(origin, ray, SKY_EARTH_RADIUS, SKY_ATMO_RADIUS, &start, &distance);
  if (distance == -ORC_FLT_MAX)
    return;
  if (distance > 0.0f) {
    const int steps = 500;
    float reach     = start;
    float step_size;
    const float light_angle = sample_sphere_solid_angle(sun, SKY_SUN_RADIUS, origin);
    Spectrum transmittance  = s_set1(1.0f);
    const JendersieEon mie  = jendersie_eon_parameters(sky->p.mie_diameter);
    for (int i = 0; i < steps; i++) {
      const float new_reach = start + distance * (i + 0.3f) / steps;
      step_size             = new_reach - reach;
      reach                 = new_reach;
      const OrcVec3 pos     = v_add(origin, v_scale(ray, reach));
      const float height    = sky_height(pos);
      const OrcVec3 ray_scatter    = v_normalize(v_sub(sun, pos));
      const float cos_angle        = v_dot(ray, ray_scatter);
      const float phase_rayleigh   = rayleigh_phase(cos_angle);
      const float phase_mie        = jendersie_eon_phase(cos_angle, &mie);
      const float zenith_cos_angle = v_dot(v_normalize(pos), ray_scatter);
      const Spectrum extinction_sun = fetch_transmittance(sky, height, zenith_cos_angle);
      const Medium m                = medium_at(&sky->p, height);
      const Spectrum scattering     = s_add(m.scattering_rayleigh, s_set1(m.scattering_mie));
      const Spectrum phase_times_scattering = s_add(s_scale(m.scattering_rayleigh, phase_rayleigh), s_set1(m.scattering_mie * phase_mie));
      const float shadow   = orc_sph_ray_hit_p0(ray_scatter, pos, SKY_EARTH_RADIUS) ? 0.0f : 1.0f;
      const Spectrum Sterm = s_scale(s_mul(extinction_sun, phase_times_scattering), shadow * light_angle);
      const Spectrum step_transmittance = s_exp(s_scale(m.extinction, -step_size));
      const Spectrum ss_int = s_mul(s_sub(Sterm, s_mul(Sterm, step_transmittance)), s_inv(m.extinction));
      const Spectrum ms_int = s_mul(s_sub(scattering, s_mul(scattering, step_transmittance)), s_inv(m.extinction));
      *L            = s_add(*L, s_mul(ss_int, transmittance));
      *as1          = s_add(*as1, s_mul(ms_int, transmittance));
      transmittance = s_mul(transmittance, step_transmittance);
    }
  }
}

static void build_multiscattering_lut(OrcSky* sky) { /* sky.cuh:276-330 */
#pragma omp parallel for schedule(dynamic, 4)
  for (int id = 0; id < SKY_MS_TEX_SIZE * SKY_MS_TEX_SIZE; id++) {
    const int y = id / SKY_MS_TEX_SIZE;
    const int x = id - y * SKY_MS_TEX_SIZE;
    float fx    = ((float) x + 0.5f) / SKY_MS_TEX_SIZE;
    float fy    = ((float) y + 0.5f) / SKY_MS_TEX_SIZE;
    fx          = sub_to_unit_uv(fx, SKY_MS_TEX_SIZE);
    fy          = sub_to_unit_uv(fy, SKY_MS_TEX_SIZE);
    const float cos_angle = fx * 2.0f - 1.0f;
    const OrcVec3 sun_dir = v_get(0.0f, cos_angle, sqrtf(orc_saturate(1.0f - cos_angle * cos_angle)));
    const float height    = SKY_EARTH_RADIUS + orc_saturate(fy + SKY_HEIGHT_OFFSET) * (SKY_ATMO_HEIGHT - SKY_HEIGHT_OFFSET);
    const OrcVec3 pos     = v_get(0.0f, height, 0.0f);
    const OrcVec3 sun     = v_scale(sun_dir, SKY_SUN_DISTANCE);
    const float sqrt_sample = (float) SKY_MS_BASE;
    static _Thread_local Spectrum lum[SKY_MS_ITER], ms[SKY_MS_ITER];
    for (int t = 0; t < SKY_MS_ITER; t++) {
      const float a     = t / SKY_MS_BASE;
      const float b     = (t - ((t / SKY_MS_BASE) * SKY_MS_BASE));
      const float randA = a / sqrt_sample;
      const float randB = b / sqrt_sample;
      const OrcVec3 ray = sample_ray_sphere(2.0f * randA - 1.0f, randB);
      multiscattering_integration(sky, pos, ray, sun, &lum[t], &ms[t]);
    }
    /* the block's shared-memory tree reduction, in its order */
    for (int i = SKY_MS_ITER >> 1; i > 0; i >>= 1)
      for (int t = 0; t < i; t++) {
        lum[t] = s_add(lum[t], lum[t + i]);
        ms[t]  = s_add(ms[t], ms[t + i]);
      }
    const Spectrum luminance       = s_scale(lum[0], 1.0f / (sqrt_sample * sqrt_sample));
    const Spectrum multiscattering = s_scale(ms[0], 1.0f / (sqrt_sample * sqrt_sample));
    const Spectrum contribution    = s_inv(s_sub(s_set1(1.0f), multiscattering));
    const Spectrum out             = s_scale(s_mul(luminance, contribution), sky->p.multiscattering_factor);
    memcpy(sky->ms_low + 4 * (size_t) id, out.v, 16);
    memcpy(sky->ms_high + 4 * (size_t) id, out.v + 4, 16);
  }
}

/* sky_compute_atmosphere without cloud shadows, sky.cuh:338-502 */
static Spectrum compute_atmosphere_t(const OrcSky* sky, OrcVec3 origin, OrcVec3 ray, float limit, bool celestials, int steps, float random_offset,
                                     Spectrum* transmittance_out) {
  const OrcSkyParams* S = &sky->p;
  Spectrum result       = s_set1(0.0f);
  float start, path_len;
  compute_path(origin, ray, SKY_EARTH_RADIUS, SKY_ATMO_RADIUS, &start, &path_len);
  const float distance   = fminf(path_len, limit - start);
  Spectrum transmittance = S_IDENT;
  const OrcTexture ml = lut_texture(sky->ms_low, SKY_MS_TEX_SIZE, SKY_MS_TEX_SIZE), mh = lut_texture(sky->ms_high, SKY_MS_TEX_SIZE, SKY_MS_TEX_SIZE);

  if (distance > 0.0f) {
    float reach = start;
    float step_size;
    const float light_angle = sample_sphere_solid_angle(sky->sun_pos, SKY_SUN_RADIUS, origin);
    const JendersieEon mie  = jendersie_eon_parameters(S->mie_diameter);
    for (int i = 0; i < steps; i++) {
      const float new_reach = start + distance * (i + random_offset) / steps;
      step_size             = new_reach - reach;
      reach                 = new_reach;
      const OrcVec3 pos     = v_add(origin, v_scale(ray, reach));
      const float height    = sky_height(pos);
      const OrcVec3 ray_scatter    = v_normalize(v_sub(sky->sun_pos, pos));
      const float cos_angle        = v_dot(ray, ray_scatter);
      const float zenith_cos_angle = v_dot(v_normalize(pos), ray_scatter);
      const float phase_rayleigh   = rayleigh_phase(cos_angle);
      const float phase_mie        = jendersie_eon_phase(cos_angle, &mie);
      const float shadow           = orc_sph_ray_hit_p0(ray_scatter, pos, SKY_EARTH_RADIUS) ? 0.0f : 1.0f;
      const Spectrum extinction_sun = fetch_transmittance(sky, height, zenith_cos_angle);
      const Medium m                = medium_at(S, height);
      const Spectrum scattering     = s_add(m.scattering_rayleigh, s_set1(m.scattering_mie));
      const Spectrum phase_times_scattering = s_add(s_scale(m.scattering_rayleigh, phase_rayleigh), s_set1(m.scattering_mie * phase_mie));
      const Spectrum ss_radiance            = s_scale(s_mul(extinction_sun, phase_times_scattering), shadow * light_angle);
      float low[4], high[4];
      orc_texture_fetch(&ml, zenith_cos_angle * 0.5f + 0.5f, height / SKY_ATMO_HEIGHT, low);
      orc_texture_fetch(&mh, zenith_cos_angle * 0.5f + 0.5f, height / SKY_ATMO_HEIGHT, high);
      const Spectrum ms_radiance = s_mul(s_merge(low, high), scattering);
      const Spectrum Ssum        = s_add(ss_radiance, ms_radiance);
      const Spectrum step_transmittance = s_exp(s_scale(m.extinction, -step_size));
      const Spectrum Sint               = s_mul(s_sub(Ssum, s_mul(Ssum, step_transmittance)), s_inv(m.extinction));
      result                            = s_add(result, s_mul(Sint, transmittance));
      transmittance                     = s_mul(transmittance, step_transmittance);
    }
    result = s_mul(result, s_scale(S_SUN_RADIANCE, S->sun_strength));
  }

  if (celestials) {
    const float sun_hit   = sphere_ray_intersection(ray, origin, sky->sun_pos, SKY_SUN_RADIUS);
    const float earth_hit = sph_ray_int_p0(ray, origin, SKY_EARTH_RADIUS);
    const float moon_hit  = sphere_ray_intersection(ray, origin, sky->moon_pos, SKY_MOON_RADIUS);
    if (earth_hit > sun_hit && moon_hit > sun_hit)
      result = s_add(result, s_mul(transmittance, s_scale(S_SUN_RADIANCE, S->sun_strength)));
    else if (earth_hit > moon_hit) { /* the moon's surface lit by the sun, sky.cuh:440-475 */
      const OrcVec3 mp         = v_add(origin, v_scale(ray, moon_hit));
      const OrcVec3 bounce_ray = v_normalize(v_sub(sky->sun_pos, mp));
      if (!orc_sphere_ray_hit(bounce_ray, mp, v_get(0.0f, 0.0f, 0.0f), SKY_EARTH_RADIUS)) {
        OrcVec3 normal    = v_normalize(v_sub(mp, sky->moon_pos));
        const float tex_u = 0.5f + S->moon_tex_offset + atan2f(normal.z, normal.x) * (1.0f / (2.0f * PI_F));
        const float tex_v = 0.5f + asinf(normal.y) * (1.0f / PI_F);
        /* texture_load with default arguments (texture_utils.cuh:13-45): v flipped, gamma applied, (0, 0, 0, 0) for an absent texture */
        float nv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (sky->moon_normal.data) {
          orc_texture_fetch(&sky->moon_normal, tex_u, 1.0f - tex_v, nv);
          for (int k = 0; k < 3; k++)
            nv[k] = powf(nv[k], sky->moon_normal.gamma);
        }
        /* create_basis + transform_vec3, math.cuh:301-321, 445-453 */
        const float sign = copysignf(1.0f, normal.z);
        const float a    = -1.0f / (sign + normal.z);
        const float b    = normal.x * normal.y * a;
        const OrcVec3 u1 = v_get(1.0f + sign * normal.x * normal.x * a, sign * b, -sign * normal.x);
        const OrcVec3 u2 = v_get(b, sign + normal.y * normal.y * a, -normal.y);
        const OrcVec3 mn = v_get(nv[0] * 2.0f - 1.0f, nv[1] * 2.0f - 1.0f, nv[2] * 2.0f - 1.0f);
        normal = v_normalize(v_get(u1.x * mn.x + u2.x * mn.y + normal.x * mn.z, u1.y * mn.x + u2.y * mn.y + normal.y * mn.z,
                                   u1.z * mn.x + u2.z * mn.y + normal.z * mn.z));
        const float NdotL = v_dot(normal, bounce_ray);
        if (NdotL > 0.0f) {
          float av[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          if (sky->moon_albedo.data) {
            orc_texture_fetch(&sky->moon_albedo, tex_u, 1.0f - tex_v, av);
            av[0] = powf(av[0], sky->moon_albedo.gamma);
          }
          const float light_angle = sample_sphere_solid_angle(sky->sun_pos, SKY_SUN_RADIUS, mp);
          const float weight      = av[0] * S->sun_strength * NdotL * light_angle / (2.0f * PI_F);
          const Spectrum flux     = {{1.7f, 1.8f, 2.0f, 1.9f, 1.87f, 1.7f, 1.65f, 1.55f}}; /* SKY_MOON_SOLAR_FLUX */
          result                  = s_add(result, s_mul(transmittance, s_mul(flux, s_scale(S_SUN_RADIANCE, weight))));
        }
      }
    }
    if (sky->has_stars && sun_hit == ORC_FLT_MAX && earth_hit == ORC_FLT_MAX && moon_hit == ORC_FLT_MAX) {
      const float ray_altitude = asinf(ray.y);
      const float ray_azimuth  = atan2f(-ray.z, -ray.x) + PI_F;
      const uint32_t x    = (uint32_t) (ray_azimuth * 10.0f);
      const uint32_t y    = (uint32_t) ((ray_altitude + PI_F * 0.5f) * 10.0f);
      const uint32_t grid = x + y * STARS_GRID_X;
      if (grid < STARS_GRID_X * STARS_GRID_Y) {
        const uint32_t a = sky->stars_offsets[grid], b = sky->stars_offsets[grid + 1];
        for (uint32_t i = a; i < b; i++) {
          const float* star      = sky->stars + 4 * (size_t) i;
          const OrcVec3 star_pos = v_get(cosf(star[1]) * cosf(star[0]), sinf(star[0]), sinf(star[1]) * cosf(star[0]));
          if (orc_sphere_ray_hit(ray, v_get(0.0f, 0.0f, 0.0f), star_pos, star[2]))
            result = s_add(result, s_scale(transmittance, star[3] * S->stars_intensity));
        }
      }
    }
  }
  if (transmittance_out)
    *transmittance_out = s_mul(*transmittance_out, transmittance); /* sky.cuh:499 */
  return result;
}

static Spectrum compute_atmosphere(const OrcSky* sky, OrcVec3 origin, OrcVec3 ray, float limit, bool celestials, int steps, float random_offset) {
  return compute_atmosphere_t(sky, origin, ray, limit, celestials, steps, random_offset, NULL);
}

/* sky_trace_inscattering (sky.cuh:517-532) as sky_process_inscattering_events calls it (kernels.cuh:356-389): aerial perspective of the
 * segment origin_world + [0, t] * ray. Returns the in-scattered radiance (not yet multiplied by the throughput) and the colour of the
 * segment's transmittance; depth = device.state.depth (IS_PRIMARY_RAY = depth 0). */
OrcRGB orc_sky_inscattering(const OrcSky* sky, OrcVec3 origin_world, OrcVec3 ray, float t, uint32_t depth, float random_steps, float random_offset,
                            OrcRGB* transmittance_rgb) {
  const OrcVec3 sky_origin = orc_world_to_sky(sky, origin_world);
  const float limit        = t * 0.001f;
  const float base_range   = (depth == 0) ? 40.0f : 80.0f;
  const int steps          = (int) (fminf(fmaxf(0.5f, limit / base_range), 2.0f) * (float) (int) (sky->p.steps / 6) + random_steps - 0.5f);
  Spectrum transmittance   = s_set1(1.0f);
  const Spectrum radiance  = compute_atmosphere_t(sky, sky_origin, ray, limit, false, steps, random_offset, &transmittance);
  *transmittance_rgb       = color_from_spectrum(transmittance);
  return color_from_spectrum(radiance);
}

/* sky_color_main, DEFAULT mode (sky.cuh:567-576) */
OrcRGB orc_sky_color(const OrcSky* sky, OrcVec3 origin_world, OrcVec3 ray, bool include_sun, float random_offset) {
  const OrcVec3 sky_origin = orc_world_to_sky(sky, origin_world);
  return color_from_spectrum(compute_atmosphere(sky, sky_origin, ray, ORC_FLT_MAX, include_sun, (int) sky->p.steps, random_offset));
}

/* ------------------------------------------------------------------ */
/* HDRI mode: the sky baked into a latitude / longitude table (cuda/sky_hdri.cuh, device_sky.c:232-346)                    */
/* ------------------------------------------------------------------ */
/* sky_hdri_warp_apply_median_of_means, sky_hdri.cuh:13-58: insertion sort of the per-lane means, Gini-weighted trimmed mean */
static float median_of_means(float* buckets, uint32_t num_buckets) {
  for (uint32_t i = 1; i < num_buckets; i++) {
    const float x = buckets[i];
    uint32_t j    = i;
    while (j > 0 && buckets[j - 1] > x) {
      buckets[j] = buckets[j - 1];
      j--;
    }
    buckets[j] = x;
  }
  float num = 0.0f, denom = 0.0f;
  for (uint32_t b = 0; b < num_buckets; b++) {
    num += b * buckets[b];
    denom += buckets[b];
  }
  num *= 2.0f;
  denom *= num_buckets;
  const float G    = fminf(fmaxf((num / denom) - (num_buckets + 1.0f) / num_buckets, 0.0f), 1.0f); /* __saturatef: NaN -> 0 */
  const uint32_t k = num_buckets >> 1;
  const uint32_t c = (uint32_t) (k - (1.0f - ((G == G) ? G : 0.0f)) * k);
  float output = 0.0f;
  for (uint32_t b = c; b < num_buckets - c; b++)
    output += buckets[b];
  output /= num_buckets - 2 * c;
  return output;
}

/* sky_hdri_sample, sky_utils.cuh:49-63: float4 texture, point filter, wrap addressing, normalised coordinates */
static OrcRGB hdri_sample(const OrcSky* sky, OrcVec3 ray) {
  const float theta = atan2f(ray.z, ray.x);
  const float phi   = asinf(ray.y);
  const float u     = (theta + PI_F) / (2.0f * PI_F);
  const float v     = 1.0f - ((phi + 0.5f * PI_F) / PI_F);
  OrcTexture t;
  memset(&t, 0, sizeof(t));
  t.width = t.height = sky->hdri_dim, t.pitch = sky->hdri_dim * 16u, t.type = ORC_TEX_FP32, t.num_components = 4;
  t.wrap_u = t.wrap_v = 0; /* wrap */
  t.filter = 0;            /* point */
  t.gamma  = 1.0f;
  t.data   = sky->hdri_color;
  float c[4];
  orc_texture_fetch(&t, u, v, c);
  const OrcRGB r = {c[0], c[1], c[2]};
  return r;
}

/* sky_color_main (sky.cuh:567-601): mode 0 marches the atmosphere, mode 1 reads the baked table and adds the sun's disc */
OrcRGB orc_sky_color_mode(const OrcSky* sky, uint32_t mode, OrcVec3 origin_world, OrcVec3 ray, bool include_sun, float random_offset) {
  if (mode != 1 || !sky->hdri_color)
    return orc_sky_color(sky, origin_world, ray, include_sun, random_offset);
  OrcRGB c = hdri_sample(sky, ray);
  if (include_sun) {
    const OrcVec3 sky_origin  = orc_world_to_sky(sky, origin_world);
    const bool ray_hits_sun   = orc_sphere_ray_hit(ray, sky_origin, sky->sun_pos, SKY_SUN_RADIUS);
    const bool ray_hits_earth = orc_sph_ray_hit_p0(ray, sky_origin, SKY_EARTH_RADIUS);
    if (ray_hits_sun && !ray_hits_earth) {
      const OrcRGB sun = orc_sky_sun_color(sky, sky_origin, ray);
      c.r += sun.r, c.g += sun.g, c.b += sun.b;
    }
  }
  return c;
}

/* sky_compute_hdri (sky_hdri.cuh:60-158, no clouds): one warp per texel, lane l integrates the samples l, l + 32, ... */
void orc_scene_build_sky_hdri(OrcScene* s, const float origin_world[3], uint32_t dim, uint32_t sample_count, int num_threads) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#endif
  OrcSky* sky = s->sky;
  free(sky->hdri_color);
  dim              = dim ? dim : 1;
  sample_count     = sample_count ? sample_count : 1;
  sky->hdri_dim    = dim;
  sky->hdri_color  = (float*) calloc((size_t) 4 * dim * dim, sizeof(float));
  const OrcVec3 origin    = v_get(origin_world[0], origin_world[1], origin_world[2]);
  const OrcVec3 sky_origin = orc_world_to_sky(sky, origin);
  const float step_size   = 1.0f / (dim - 1);
  const uint32_t buckets  = sample_count < 32u ? sample_count : 32u;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8)
#endif
  for (int64_t pixel = 0; pixel < (int64_t) dim * dim; pixel++) {
    const uint32_t y = (uint32_t) (pixel / dim), x = (uint32_t) (pixel - (int64_t) y * dim);
    float mean[3][32];
    for (uint32_t lane = 0; lane < 32; lane++) {
      OrcRGB color = {0.0f, 0.0f, 0.0f};
      uint32_t n   = 0;
      for (uint32_t sample_id = lane; sample_id < sample_count; sample_id += 32) {
        const OrcPathID pid    = orc_path_id_get(x, y, sample_id);
        const OrcFloat2 jitter = orc_random_2d(ORC_RT_CAMERA_JITTER, pid, 0);
        const float u          = (((float) x) + jitter.x) * step_size;
        const float v          = 1.0f - (((float) y) + jitter.y) * step_size;
        const float altitude   = PI_F * v - 0.5f * PI_F;
        const float azimuth    = 2.0f * PI_F * u - PI_F;
        const OrcVec3 ray      = v_get(cosf(azimuth) * cosf(altitude), sinf(altitude), sinf(azimuth) * cosf(altitude));
        const OrcRGB sky_color = color_from_spectrum(
          compute_atmosphere(sky, sky_origin, ray, ORC_FLT_MAX, false, (int) sky->p.steps, orc_random_1d(ORC_RT_SKY_STEP_OFFSET, pid, 0)));
        color.r += sky_color.r, color.g += sky_color.g, color.b += sky_color.b;
        n++;
      }
      mean[0][lane] = n ? color.r / n : 0.0f;
      mean[1][lane] = n ? color.g / n : 0.0f;
      mean[2][lane] = n ? color.b / n : 0.0f;
    }
    float* dst = sky->hdri_color + 4 * (size_t) pixel;
    for (int c = 0; c < 3; c++)
      dst[c] = median_of_means(mean[c], buckets);
    dst[3] = 0.0f;
  }
}

/* the moon's surface textures (device_embedded_data.c:62-100); NULL = absent: the disc is a black occluder. Texels are copied. */
static void copy_texture(OrcTexture* dst, const OrcTexture* src) {
  free((void*) dst->data);
  memset(dst, 0, sizeof(*dst));
  if (!src || !src->data)
    return;
  *dst          = *src;
  const size_t n = (size_t) src->pitch * src->height;
  void* copy     = malloc(n);
  memcpy(copy, src->data, n);
  dst->data = copy;
}
void orc_scene_set_moon_textures(OrcScene* s, const OrcTexture* albedo, const OrcTexture* normal) {
  copy_texture(&s->sky->moon_albedo, albedo);
  copy_texture(&s->sky->moon_normal, normal);
}

void orc_scene_sky_hdri(const OrcScene* s, const float** color, uint32_t* dim) { *color = s->sky->hdri_color, *dim = s->sky->hdri_dim; }

void orc_scene_set_sky_hdri(OrcScene* s, const float* color, uint32_t dim) {
  free(s->sky->hdri_color);
  s->sky->hdri_dim   = dim;
  s->sky->hdri_color = (float*) malloc(sizeof(float) * 4 * (size_t) dim * dim);
  memcpy(s->sky->hdri_color, color, sizeof(float) * 4 * (size_t) dim * dim);
}

/* ------------------------------------------------------------------ */
/* public                                                               */
/* ------------------------------------------------------------------ */
void orc_scene_set_sky(OrcScene* s, const OrcSkyParams* p, int num_threads) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#endif
  if (!p) {
    orc_sky_free(s->sky);
    s->sky = NULL;
    return;
  }
  /* the LUTs depend on the medium only (SCENE_DIRTY_FLAG_INTEGRATION fields of sky_check_for_dirty) */
  bool rebuild = true;
  if (s->sky) {
    const OrcSkyParams* o = &s->sky->p;
    rebuild = o->base_density != p->base_density || o->rayleigh_density != p->rayleigh_density || o->mie_density != p->mie_density
              || o->ozone_density != p->ozone_density || o->rayleigh_falloff != p->rayleigh_falloff || o->mie_falloff != p->mie_falloff
              || o->mie_diameter != p->mie_diameter || o->ground_visibility != p->ground_visibility
              || o->ozone_layer_thickness != p->ozone_layer_thickness || o->multiscattering_factor != p->multiscattering_factor
              || o->ozone_absorption != p->ozone_absorption;
  }
  else {
    s->sky = (OrcSky*) calloc(1, sizeof(OrcSky));
    s->sky->tm_low  = (float*) calloc(4 * SKY_TM_TEX_WIDTH * SKY_TM_TEX_HEIGHT, sizeof(float));
    s->sky->tm_high = (float*) calloc(4 * SKY_TM_TEX_WIDTH * SKY_TM_TEX_HEIGHT, sizeof(float));
    s->sky->ms_low  = (float*) calloc(4 * SKY_MS_TEX_SIZE * SKY_MS_TEX_SIZE, sizeof(float));
    s->sky->ms_high = (float*) calloc(4 * SKY_MS_TEX_SIZE * SKY_MS_TEX_SIZE, sizeof(float));
    s->sky->stars_count = 0xFFFFFFFFu;
  }
  OrcSky* sky       = s->sky;
  const bool stars  = sky->stars_count != p->stars_count || sky->p.stars_seed != p->stars_seed;
  sky->p            = *p;
  sky->sun_pos      = celestial_position(p->azimuth, p->altitude, SKY_SUN_DISTANCE, p->geometry_offset);
  sky->moon_pos     = celestial_position(p->moon_azimuth, p->moon_altitude, SKY_MOON_DISTANCE, p->geometry_offset);
  if (stars)
    stars_generate(sky, p->stars_count, p->stars_seed);
  if (rebuild) {
    /* one-entry process cache: test scenes share the default medium, the tables take seconds */
    static OrcSkyParams cached_params;
    static float* cached[4] = {NULL, NULL, NULL, NULL};
    const size_t tm = sizeof(float) * 4 * SKY_TM_TEX_WIDTH * SKY_TM_TEX_HEIGHT, ms = sizeof(float) * 4 * SKY_MS_TEX_SIZE * SKY_MS_TEX_SIZE;
    const OrcSkyParams* o = &cached_params;
    const bool hit = cached[0] && o->base_density == p->base_density && o->rayleigh_density == p->rayleigh_density && o->mie_density == p->mie_density
                     && o->ozone_density == p->ozone_density && o->rayleigh_falloff == p->rayleigh_falloff && o->mie_falloff == p->mie_falloff
                     && o->mie_diameter == p->mie_diameter && o->ground_visibility == p->ground_visibility
                     && o->ozone_layer_thickness == p->ozone_layer_thickness && o->multiscattering_factor == p->multiscattering_factor
                     && o->ozone_absorption == p->ozone_absorption;
    if (hit) {
      memcpy(sky->tm_low, cached[0], tm), memcpy(sky->tm_high, cached[1], tm), memcpy(sky->ms_low, cached[2], ms), memcpy(sky->ms_high, cached[3], ms);
    }
    else {
      build_transmittance_lut(sky);
      build_multiscattering_lut(sky);
      if (!cached[0])
        cached[0] = (float*) malloc(tm), cached[1] = (float*) malloc(tm), cached[2] = (float*) malloc(ms), cached[3] = (float*) malloc(ms);
      memcpy(cached[0], sky->tm_low, tm), memcpy(cached[1], sky->tm_high, tm), memcpy(cached[2], sky->ms_low, ms), memcpy(cached[3], sky->ms_high, ms);
      cached_params = *p;
    }
  }
}

void orc_sky_free(OrcSky* sky) {
  if (!sky)
    return;
  free(sky->tm_low);
  free(sky->tm_high);
  free(sky->ms_low);
  free(sky->ms_high);
  free(sky->stars);
  free(sky->hdri_color);
  free((void*) sky->moon_albedo.data);
  free((void*) sky->moon_normal.data);
  free(sky);
}

void orc_scene_sky_luts(const OrcScene* s, const float** tm_low, const float** tm_high, const float** ms_low, const float** ms_high) {
  *tm_low = s->sky->tm_low, *tm_high = s->sky->tm_high, *ms_low = s->sky->ms_low, *ms_high = s->sky->ms_high;
}

void orc_scene_set_sky_luts(OrcScene* s, const float* tm_low, const float* tm_high, const float* ms_low, const float* ms_high) {
  memcpy(s->sky->tm_low, tm_low, sizeof(float) * 4 * SKY_TM_TEX_WIDTH * SKY_TM_TEX_HEIGHT);
  memcpy(s->sky->tm_high, tm_high, sizeof(float) * 4 * SKY_TM_TEX_WIDTH * SKY_TM_TEX_HEIGHT);
  memcpy(s->sky->ms_low, ms_low, sizeof(float) * 4 * SKY_MS_TEX_SIZE * SKY_MS_TEX_SIZE);
  memcpy(s->sky->ms_high, ms_high, sizeof(float) * 4 * SKY_MS_TEX_SIZE * SKY_MS_TEX_SIZE);
}

void orc_scene_sky_info(const OrcScene* s, float sun_pos[3], float moon_pos[3], const float** stars, const uint32_t** stars_offsets, uint32_t* stars_count) {
  sun_pos[0] = s->sky->sun_pos.x, sun_pos[1] = s->sky->sun_pos.y, sun_pos[2] = s->sky->sun_pos.z;
  moon_pos[0] = s->sky->moon_pos.x, moon_pos[1] = s->sky->moon_pos.y, moon_pos[2] = s->sky->moon_pos.z;
  *stars = s->sky->stars, *stars_offsets = s->sky->stars_offsets, *stars_count = s->sky->has_stars ? s->sky->stars_count : 0;
}

/* explicit segments through sky_trace_inscattering: -> in-scattered radiance (n, 3) and transmittance colour (n, 3) */
void orc_sky_inscatter_segments(const OrcScene* s, uint32_t n, const float* origins_world, const float* rays, const float* t, uint32_t depth,
                                const float* random_steps, const float* random_offsets, float* inscattering, float* transmittance) {
  for (uint32_t i = 0; i < n; i++) {
    OrcRGB tr;
    const OrcRGB c = orc_sky_inscattering(s->sky, v_get(origins_world[3 * i], origins_world[3 * i + 1], origins_world[3 * i + 2]),
                                          v_get(rays[3 * i], rays[3 * i + 1], rays[3 * i + 2]), t[i], depth, random_steps[i], random_offsets[i], &tr);
    inscattering[3 * i + 0] = c.r, inscattering[3 * i + 1] = c.g, inscattering[3 * i + 2] = c.b;
    transmittance[3 * i + 0] = tr.r, transmittance[3 * i + 1] = tr.g, transmittance[3 * i + 2] = tr.b;
  }
}

void orc_sky_colors_mode(const OrcScene* s, uint32_t mode, uint32_t n, const float* origins_world, const float* rays, const uint32_t* include_sun,
                         const float* random_offsets, float* rgb, int num_threads) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int64_t i = 0; i < (int64_t) n; i++) {
    const OrcRGB c = orc_sky_color_mode(s->sky, mode, v_get(origins_world[3 * i], origins_world[3 * i + 1], origins_world[3 * i + 2]),
                                        v_get(rays[3 * i], rays[3 * i + 1], rays[3 * i + 2]), include_sun[i] != 0, random_offsets[i]);
    rgb[3 * i + 0] = c.r, rgb[3 * i + 1] = c.g, rgb[3 * i + 2] = c.b;
  }
}

void orc_sky_colors(const OrcScene* s, uint32_t n, const float* origins_world, const float* rays, const uint32_t* include_sun,
                    const float* random_offsets, float* rgb, int num_threads) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
  for (int64_t i = 0; i < (int64_t) n; i++) {
    const OrcRGB c = orc_sky_color(s->sky, v_get(origins_world[3 * i], origins_world[3 * i + 1], origins_world[3 * i + 2]),
                                   v_get(rays[3 * i], rays[3 * i + 1], rays[3 * i + 2]), include_sun[i] != 0, random_offsets[i]);
    rgb[3 * i + 0] = c.r, rgb[3 * i + 1] = c.g, rgb[3 * i + 2] = c.b;
  }
}
