/*
 * orc_core.c - oracle: PathID, counter-based RNG, packing, transforms, thin-lens camera.
 * TEST INFRASTRUCTURE ONLY (see lum_oracle.h). Every function restates one reference function.
 */
#include <math.h>
#include <string.h>

#include "lum_oracle.h"
#include "orc_internal.h"

/* ------------------------------------------------------------------ */
/* PathID: cuda/utils.cuh:142-178, constants device_utils.h:37-39      */
/* ------------------------------------------------------------------ */
#define PID_SENSOR_BITS 14
#define PID_SAMPLE_BITS 20
#define PID_EXTRA_BITS (16 - PID_SENSOR_BITS)
#define PID_SENSOR_MASK ((1u << PID_SENSOR_BITS) - 1)
#define PID_SAMPLE_MASK ((1u << PID_SAMPLE_BITS) - 1)
#define PID_EXTRA_MASK ((1u << PID_EXTRA_BITS) - 1)

OrcPathID orc_path_id_get(uint32_t x, uint32_t y, uint32_t sample_id) {
  OrcPathID id;
  id.x = (uint16_t) ((x & PID_SENSOR_MASK) | (((sample_id >> (16 + PID_EXTRA_BITS * 0)) & PID_EXTRA_MASK) << PID_SENSOR_BITS));
  id.y = (uint16_t) ((y & PID_SENSOR_MASK) | (((sample_id >> (16 + PID_EXTRA_BITS * 1)) & PID_EXTRA_MASK) << PID_SENSOR_BITS));
  /* the reference stores `sample_id & PATH_ID_SAMPLE_MASK` into a uint16_t, i.e. the low 16 bits survive */
  id.z = (uint16_t) (sample_id & PID_SAMPLE_MASK);
  return id;
}

void orc_path_id_pixel(OrcPathID id, uint32_t* x, uint32_t* y) {
  *x = id.x & PID_SENSOR_MASK;
  *y = id.y & PID_SENSOR_MASK;
}

uint32_t orc_path_id_sample(OrcPathID id) {
  uint32_t s = id.z;
  s |= ((uint32_t) id.x >> PID_SENSOR_BITS) << (16 + PID_EXTRA_BITS * 0);
  s |= ((uint32_t) id.y >> PID_SENSOR_BITS) << (16 + PID_EXTRA_BITS * 1);
  return s;
}

/* ------------------------------------------------------------------ */
/* RNG: cuda/random.cuh                                                 */
/* ------------------------------------------------------------------ */
static inline uint32_t swap16(uint32_t a) { return (a >> 16) | (a << 16); } /* intrinsics.cuh:37-45 */

static inline uint32_t brev32(uint32_t x) {
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
  x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
  return (x >> 16) | (x << 16);
}

/* random_uint32_t_base, random.cuh:172-194 (Squares, Widynski 2020, 32-bit variant) */
uint32_t orc_squares32(uint32_t key, uint32_t counter) {
  uint32_t x = counter * key;
  uint32_t y = counter * key;
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

/* random_uint16_t_base, random.cuh:197-213 */
uint16_t orc_squares16(uint32_t key, uint32_t counter) {
  uint32_t x = counter * key;
  uint32_t y = counter * key;
  uint32_t z = y + key;

  x = x * x + y;
  x = swap16(x);
  x = x * x + z;
  x = swap16(x);

  return (uint16_t) ((x * x + y) >> 16);
}

/* random_laine_karras_permutation, random.cuh:238-245 */
static inline uint32_t lk_perm(uint32_t x, uint32_t seed) {
  x += seed;
  x ^= x * 0x6c50b47cu;
  x ^= x * 0xb82f1e52u;
  x ^= x * 0xc7afe638u;
  x ^= x * 0x8d22f6e6u;
  return x;
}

/* random_nested_uniform_scramble_base2, random.cuh:247-252 */
static inline uint32_t owen_scramble(uint32_t x, uint32_t seed) { return brev32(lk_perm(brev32(x), seed)); }

/* random_hash_combine, random.cuh:254-256 */
static inline uint32_t hash_combine(uint32_t seed, uint32_t v) { return seed ^ (v + (seed << 6) + (seed >> 2)); }

/* random_sobol_P, random.cuh:258-265 */
static inline uint32_t sobol_P(uint32_t v) {
  v ^= v << 16;
  v ^= (v & 0x00FF00FFu) << 8;
  v ^= (v & 0x0F0F0F0Fu) << 4;
  v ^= (v & 0x33333333u) << 2;
  v ^= (v & 0x55555555u) << 1;
  return v;
}

/* random_sobol, random.cuh:267-287 */
OrcUint2 orc_sobol(uint32_t offset, uint32_t dimension) {
  const uint32_t seed = orc_squares32(0xfcbd6e15u, dimension);
  const uint32_t J    = lk_perm(brev32(offset), seed);
  const uint32_t sx   = J;
  const uint32_t sy   = sobol_P(J);

  OrcUint2 r;
  r.x = owen_scramble(sx, hash_combine(seed, 0));
  r.y = owen_scramble(sy, hash_combine(seed, 1));
  return r;
}

static const uint32_t* g_bluenoise = NULL;

void orc_set_bluenoise(const uint32_t* table) { g_bluenoise = table; }

#define R2_PHI1 3242174889u
#define R2_PHI2 2447445413u

/* random_2D_base, random.cuh:320-333 with random_r2 (:226-231) and random_blue_noise_mask_2D (:309-314) */
OrcUint2 orc_random_2d_base(uint32_t target, uint32_t px, uint32_t py, uint32_t sequence_id, uint32_t depth) {
  const uint32_t dim = target + depth * ORC_RT_COUNT;

  OrcUint2 quasi = orc_sobol(sequence_id, dim);

  const uint32_t ox = ((1u + dim) * R2_PHI1) >> 24;
  const uint32_t oy = ((1u + dim) * R2_PHI2) >> 24;

  const uint32_t bx    = (px + ox) & 0xFFu;
  const uint32_t by    = (py + oy) & 0xFFu;
  const uint32_t noise = g_bluenoise ? g_bluenoise[bx + by * 256u] : 0u;

  quasi.x += noise & 0xFFFF0000u;
  quasi.y += noise << 16;
  return quasi;
}

/* random_uint32_t_to_float, random.cuh:144-148 */
float orc_u32_to_float(uint32_t v) {
  const uint32_t i = 0x3F800000u | (v >> 9);
  float f;
  memcpy(&f, &i, 4);
  return f - 1.0f;
}

/* random_uint16_t_to_float, random.cuh:150-154 */
float orc_u16_to_float(uint16_t v) {
  const uint32_t i = 0x3F800000u | (((uint32_t) v) << 7);
  float f;
  memcpy(&f, &i, 4);
  return f - 1.0f;
}

/* random_2D / random_1D, random.cuh:343-368 (depth = device.state.depth passed explicitly) */
OrcFloat2 orc_random_2d(uint32_t target, OrcPathID id, uint32_t depth) {
  uint32_t px, py;
  orc_path_id_pixel(id, &px, &py);
  const OrcUint2 q = orc_random_2d_base(target, px, py, orc_path_id_sample(id), depth);
  OrcFloat2 r      = {orc_u32_to_float(q.x), orc_u32_to_float(q.y)};
  return r;
}

float orc_random_1d(uint32_t target, OrcPathID id, uint32_t depth) {
  uint32_t px, py;
  orc_path_id_pixel(id, &px, &py);
  return orc_u32_to_float(orc_random_2d_base(target, px, py, orc_path_id_sample(id), depth).x);
}

/* random_saturate, random.cuh:163-165 */
float orc_random_saturate(float r) {
  const uint32_t mb = 0x3F7FFFFFu;
  float mx;
  memcpy(&mx, &mb, 4);
  return fminf(fmaxf(r, 0.0f), mx);
}

/* ------------------------------------------------------------------ */
/* Packing                                                              */
/* ------------------------------------------------------------------ */
static inline uint32_t f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float u2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline float saturatef(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

uint32_t orc_pack_normal_host(OrcVec3 n) {
  double x = n.x, y = n.y, z = n.z;
  const double rn = 1.0 / (fabs(x) + fabs(y) + fabs(z));
  x *= rn;
  y *= rn;
  z *= rn;
  const double t = fmax(fmin(-z, 1.0), 0.0);
  x += (x >= 0.0) ? t : -t;
  y += (y >= 0.0) ? t : -t;
  x = fmax(fmin(x, 1.0), -1.0);
  y = fmax(fmin(y, 1.0), -1.0);
  x = (x + 1.0) * 0.5;
  y = (y + 1.0) * 0.5;
  const uint32_t xu = (uint32_t) (x * 0xFFFF + 0.5);
  const uint32_t yu = (uint32_t) (y * 0xFFFF + 0.5);
  return (yu << 16) | xu;
}

uint32_t orc_pack_normal(OrcVec3 n) {
  float x = n.x, y = n.y, z = n.z;
  const float rn = 1.0f / (fabsf(x) + fabsf(y) + fabsf(z));
  x *= rn;
  y *= rn;
  z *= rn;
  const float t = fmaxf(fminf(-z, 1.0f), 0.0f);
  x += (x >= 0.0f) ? t : -t;
  y += (y >= 0.0f) ? t : -t;
  x = fmaxf(fminf(x, 1.0f), -1.0f);
  y = fmaxf(fminf(y, 1.0f), -1.0f);
  x = (x + 1.0f) * 0.5f;
  y = (y + 1.0f) * 0.5f;
  const uint32_t xu = (uint32_t) (x * 0xFFFF + 0.5f);
  const uint32_t yu = (uint32_t) (y * 0xFFFF + 0.5f);
  return (yu << 16) | xu;
}

OrcVec3 orc_unpack_normal(uint32_t data) {
  float x = (data & 0xFFFF) * (1.0f / 0xFFFF);
  float y = (data >> 16) * (1.0f / 0xFFFF);
  x       = (x * 2.0f) - 1.0f;
  y       = (y * 2.0f) - 1.0f;
  OrcVec3 n     = {x, y, 1.0f - fabsf(x) - fabsf(y)};
  const float t = saturatef(-n.z);
  n.x += (n.x >= 0.0f) ? -t : t;
  n.y += (n.y >= 0.0f) ? -t : t;
  return v_normalize(n);
}

uint32_t orc_pack_uv(float u, float v) { return (f2u(u) & 0xFFFF0000u) | (f2u(v) >> 16); }

OrcFloat2 orc_unpack_uv(uint32_t p) {
  OrcFloat2 r = {u2f(p & 0xFFFF0000u), u2f(p << 16)};
  return r;
}

OrcUint2 orc_record_pack(OrcRGB c) {
  const uint32_t r = f2u(c.r) >> 11, g = f2u(c.g) >> 11, b = f2u(c.b) >> 11;
  OrcUint2 p;
  p.x = r | (g << 21);
  p.y = (g >> 11) | (b << 10);
  return p;
}

OrcRGB orc_record_unpack(OrcUint2 p) {
  const uint32_t r = p.x & 0x1FFFFF;
  const uint32_t g = (p.x >> 21) | ((p.y & 0x3FF) << 11);
  const uint32_t b = p.y >> 10;
  OrcRGB c         = {u2f(r << 11), u2f(g << 11), u2f(b << 11)};
  return c;
}

OrcUint2 orc_ray_pack(OrcVec3 ray) {
  float x = ray.x, y = ray.y, z = ray.z;
  const float rn = 1.0f / (fabsf(x) + fabsf(y) + fabsf(z));
  x *= rn;
  y *= rn;
  z *= rn;
  const float t = saturatef(-z);
  x += (x >= 0.0f) ? t : -t;
  y += (y >= 0.0f) ? t : -t;
  x = fminf(1.0f, fmaxf(-1.0f, x));
  y = fminf(1.0f, fmaxf(-1.0f, y));
  x = (x + 1.0f) * 0.5f;
  y = (y + 1.0f) * 0.5f;
  OrcUint2 p;
  /* float -> uint32 conversion saturates on the device; 0xFFFFFFFF as float is 2^32 */
  const float fx = x * 4294967296.0f + 0.5f;
  const float fy = y * 4294967296.0f + 0.5f;
  p.x            = (fx >= 4294967296.0f) ? 0xFFFFFFFFu : (uint32_t) fx;
  p.y            = (fy >= 4294967296.0f) ? 0xFFFFFFFFu : (uint32_t) fy;
  return p;
}

OrcVec3 orc_ray_unpack(OrcUint2 p) {
  float x = (float) p.x * (1.0f / 4294967296.0f);
  float y = (float) p.y * (1.0f / 4294967296.0f);
  x       = (x * 2.0f) - 1.0f;
  y       = (y * 2.0f) - 1.0f;
  OrcVec3 r     = {x, y, 1.0f - fabsf(x) - fabsf(y)};
  const float t = saturatef(-r.z);
  r.x += (r.x >= 0.0f) ? -t : t;
  r.y += (r.y >= 0.0f) ? -t : t;
  return v_normalize(r);
}

uint32_t orc_ior_compress(float ior) { return (f2u((0.5f * (ior - 1.0f)) + 1.0f) >> 15) & 0xFF; }

float orc_ior_decompress(uint32_t c) { return ((u2f(0x3F800000u | (c << 15)) - 1.0f) * 2.0f) + 1.0f; }

OrcQuat orc_euler_to_quat(OrcVec3 rot) {
  const float cr = cosf(rot.x * 0.5f), sr = sinf(rot.x * 0.5f);
  const float cp = cosf(rot.y * 0.5f), sp = sinf(rot.y * 0.5f);
  const float cy = cosf(rot.z * 0.5f), sy = sinf(rot.z * 0.5f);
  OrcQuat q;
  q.w = cr * cp * cy + sr * sp * sy;
  q.x = sr * cp * cy - cr * sp * sy;
  q.y = cr * sp * cy + sr * cp * sy;
  q.z = cr * cp * sy - sr * sp * cy;
  return q;
}

OrcQuat16 orc_quat_pack(OrcQuat q) {
  OrcQuat16 d;
  d.x = (uint16_t) (((1.0f - q.x) * 0x7FFF) + 0.5f);
  d.y = (uint16_t) (((1.0f - q.y) * 0x7FFF) + 0.5f);
  d.z = (uint16_t) (((1.0f - q.z) * 0x7FFF) + 0.5f);
  d.w = (uint16_t) (((1.0f + q.w) * 0x7FFF) + 0.5f);
  return d;
}

OrcVec3 orc_quat_apply(OrcQuat q, OrcVec3 v) {
  const OrcVec3 u    = {q.x, q.y, q.z};
  const float s      = q.w;
  const float dot_uv = v_dot(u, v);
  const float dot_uu = v_dot(u, u);
  const OrcVec3 c    = v_cross(u, v);
  OrcVec3 r          = v_scale(u, 2.0f * dot_uv);
  r                  = v_add(r, v_scale(v, s * s - dot_uu));
  r                  = v_add(r, v_scale(c, 2.0f * s));
  return r;
}

/* quaternion16_apply / _inv, cuda/math.cuh:429-449 */
static OrcVec3 quat16_apply(OrcQuat16 q, OrcVec3 v) {
  OrcQuat f;
  f.x = (q.x * (1.0f / 0x7FFF)) - 1.0f;
  f.y = (q.y * (1.0f / 0x7FFF)) - 1.0f;
  f.z = (q.z * (1.0f / 0x7FFF)) - 1.0f;
  f.w = (q.w * (1.0f / 0x7FFF)) - 1.0f;
  return orc_quat_apply(f, v);
}

static OrcVec3 quat16_apply_inv(OrcQuat16 q, OrcVec3 v) {
  OrcQuat f;
  f.x = 1.0f - (q.x * (1.0f / 0x7FFF));
  f.y = 1.0f - (q.y * (1.0f / 0x7FFF));
  f.z = 1.0f - (q.z * (1.0f / 0x7FFF));
  f.w = (q.w * (1.0f / 0x7FFF)) - 1.0f;
  return orc_quat_apply(f, v);
}

OrcVec3 orc_transform_apply_rotation(const OrcTransform* t, OrcVec3 v) { return quat16_apply(t->rotation, v); }
OrcVec3 orc_transform_apply_rotation_inv(const OrcTransform* t, OrcVec3 v) { return quat16_apply_inv(t->rotation, v); }
OrcVec3 orc_transform_apply_relative(const OrcTransform* t, OrcVec3 v) { return v_mul(quat16_apply(t->rotation, v), t->scale); }

OrcVec3 orc_transform_apply(const OrcTransform* t, OrcVec3 v) { return v_add(orc_transform_apply_relative(t, v), t->translation); }

OrcVec3 orc_transform_apply_inv(const OrcTransform* t, OrcVec3 v) {
  const OrcVec3 inv = {1.0f / t->scale.x, 1.0f / t->scale.y, 1.0f / t->scale.z};
  return quat16_apply_inv(t->rotation, v_mul(v_sub(v, t->translation), inv));
}

/* device_struct_material_convert, device_structs.c:257-330 */
static uint16_t f01_to_u16(float f) {
  /* _device_struct_convert_float01_to_uint16, device_structs.c:250-252 (no clamp in the reference) */
  return (uint16_t) (f * 65535.0f + 0.5f);
}

void orc_material_pack(const OrcMaterialDesc* m, OrcMaterialPacked* d) {
  memset(d, 0, sizeof(*d));
  d->flags |= m->emission_active ? 0x02 : 0;
  d->flags |= m->thin_walled ? 0x04 : 0;
  d->flags |= m->metallic ? 0x08 : 0;
  d->flags |= m->colored_transparency ? 0x10 : 0;
  d->flags |= m->roughness_as_smoothness ? 0x20 : 0;
  d->flags |= m->normal_map_is_compressed ? 0x40 : 0;
  d->flags |= m->bidirectional_emission ? 0x80 : 0;
  d->flags |= (m->base_substrate == 1) ? 0x01 : 0;

  d->roughness_clamp  = (uint8_t) (f01_to_u16(m->roughness_clamp) >> 8);
  d->roughness        = f01_to_u16(m->roughness);
  d->refraction_index = f01_to_u16(0.5f * (m->refraction_index - 1.0f));

  float er = m->emission[0], eg = m->emission[1], eb = m->emission[2];
  const float en = 1.0f / fminf(fmaxf(fmaxf(er, eg), eb) + 1.0f, 65535.0f);
  er *= en;
  eg *= en;
  eb *= en;

  d->albedo_r       = f01_to_u16(m->albedo[0]);
  d->albedo_g       = f01_to_u16(m->albedo[1]);
  d->albedo_b       = f01_to_u16(m->albedo[2]);
  d->albedo_a       = f01_to_u16(m->albedo[3]);
  d->emission_r     = f01_to_u16(er);
  d->emission_g     = f01_to_u16(eg);
  d->emission_b     = f01_to_u16(eb);
  d->emission_scale = (uint16_t) ((f2u(m->emission_scale / en) >> 15) & 0xFFFF);
  d->albedo_tex = d->luminance_tex = d->roughness_tex = d->metallic_tex = d->normal_tex = 0xFFFF; /* TEXTURE_NONE */
}

/* ------------------------------------------------------------------ */
/* Camera: cuda/camera.cuh:11-38, camera_thin_lens.cuh:8-86, camera_utils.cuh:23-27 */
/* ------------------------------------------------------------------ */
void orc_camera_sample(const OrcCamera* cam, const OrcSettings* s, OrcPathID id, OrcVec3* origin, OrcVec3* dir) {
  const uint32_t sample_id = orc_path_id_sample(id);
  /* camera_get_jitter: pixel (0,0), depth 0 => one sub-pixel offset per pass */
  const OrcUint2 jq    = orc_random_2d_base(ORC_RT_CAMERA_JITTER, 0, 0, sample_id, 0);
  const float jx       = orc_u32_to_float(jq.x);
  const float jy       = orc_u32_to_float(jq.y);

  const float step = 2.0f * (cam->fov / s->width);
  const float vfov = step * s->height * 0.5f;

  uint32_t px, py;
  orc_path_id_pixel(id, &px, &py);

  OrcVec3 sensor;
  sensor.x = cam->fov - step * (px + jx);
  sensor.y = -vfov + step * (py + jy);
  sensor.z = 1.0f;

  const OrcVec3 zero            = {0.0f, 0.0f, 0.0f};
  const OrcVec3 sensor_to_focal = v_normalize(v_sub(zero, sensor));

  const float focal_length = fmaxf(cam->object_distance * (1.0f / 0.001f), 0.01f); /* CAMERA_COMMON_INV_SCALE folds to 999.99994f */
  const OrcVec3 focal_point = v_scale(sensor_to_focal, -focal_length / sensor_to_focal.z);

  OrcVec3 aperture = zero;
  if (cam->aperture_size != 0.0f) {
    const OrcFloat2 r   = orc_random_2d(ORC_RT_LENS, id, 0);
    const float ap_size = cam->aperture_size * (1.0f / 0.001f);
    float sx, sy;
    if (cam->aperture_shape == 1) {
      const int blade        = (int) (orc_random_1d(ORC_RT_LENS_BLADE, id, 0) * cam->aperture_blade_count);
      const float alpha      = sqrtf(r.x);
      const float beta       = r.y;
      const float u          = 1.0f - alpha;
      const float v          = alpha * beta;
      const float angle_step = (2.0f * ORC_PI) / cam->aperture_blade_count;
      const float a1         = angle_step * blade;
      const float a2         = angle_step * (blade + 1);
      sx                     = (sinf(a1) * u + sinf(a2) * v) * ap_size;
      sy                     = (cosf(a1) * u + cosf(a2) * v) * ap_size;
    }
    else {
      const float alpha = r.x * 2.0f * ORC_PI;
      const float beta  = sqrtf(r.y) * ap_size;
      sx                = cosf(alpha) * beta;
      sy                = sinf(alpha) * beta;
    }
    aperture.x = sx;
    aperture.y = sy;
  }

  OrcVec3 o = aperture;
  OrcVec3 d = v_normalize(v_sub(focal_point, aperture));

  /* camera_sample: to world space */
  o = orc_quat_apply(cam->rotation, o);
  o = v_scale(o, cam->camera_scale * 0.001f);
  o = v_add(o, cam->pos);
  d = orc_quat_apply(cam->rotation, d);

  *origin = o;
  *dir    = d;
}
