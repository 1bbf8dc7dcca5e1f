/* orc_internal.h - small vector helpers shared by the oracle translation units. TEST INFRASTRUCTURE ONLY.
 * Semantics follow cuda/math.cuh:14-213 of the reference; rsqrtf() there is an approximation under
 * --use_fast_math, here it is 1/sqrtf (IEEE), which is the documented fp32 tolerance of the shading parity. */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H

#include <math.h>
#include <string.h>

#include "lum_oracle.h"

static inline OrcVec3 v_get(float x, float y, float z) {
  OrcVec3 r = {x, y, z};
  return r;
}
static inline OrcVec3 v_add(OrcVec3 a, OrcVec3 b) { return v_get(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline OrcVec3 v_sub(OrcVec3 a, OrcVec3 b) { return v_get(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline OrcVec3 v_mul(OrcVec3 a, OrcVec3 b) { return v_get(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline OrcVec3 v_scale(OrcVec3 a, float s) { return v_get(a.x * s, a.y * s, a.z * s); }
static inline float v_dot(OrcVec3 a, OrcVec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline OrcVec3 v_cross(OrcVec3 a, OrcVec3 b) {
  return v_get(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float v_len(OrcVec3 a) { return sqrtf(v_dot(a, a)); }
static inline OrcVec3 v_normalize(OrcVec3 a) { return v_scale(a, 1.0f / sqrtf(v_dot(a, a))); }

static inline OrcRGB c_get(float r, float g, float b) {
  OrcRGB c = {r, g, b};
  return c;
}
static inline OrcRGB c_splat(float v) { return c_get(v, v, v); }
static inline OrcRGB c_add(OrcRGB a, OrcRGB b) { return c_get(a.r + b.r, a.g + b.g, a.b + b.b); }
static inline OrcRGB c_sub(OrcRGB a, OrcRGB b) { return c_get(a.r - b.r, a.g - b.g, a.b - b.b); }
static inline OrcRGB c_mul(OrcRGB a, OrcRGB b) { return c_get(a.r * b.r, a.g * b.g, a.b * b.b); }
static inline OrcRGB c_scale(OrcRGB a, float s) { return c_get(a.r * s, a.g * s, a.b * s); }
static inline int c_any(OrcRGB a) { return a.r > 0.0f || a.g > 0.0f || a.b > 0.0f; } /* math.cuh:944-946: NaN fails the test */
static inline float c_importance(OrcRGB a) { return fmaxf(a.r, fmaxf(a.g, a.b)); }        /* math.cuh:1066-1068 */
static inline float c_luminance(OrcRGB v) { return 0.212655f * v.r + 0.715158f * v.g + 0.072187f * v.b; }

static inline float orc_saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
static inline uint32_t orc_f2u(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float orc_u2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}

/* procedural sky attached to a scene (orc_sky.c) */
typedef struct OrcSky {
  OrcSkyParams p;
  OrcVec3 sun_pos, moon_pos;                /* device_struct_sky_convert, device_structs.c:132-170 */
  float *tm_low, *tm_high, *ms_low, *ms_high; /* 256 x 64 and 32 x 32 float4 LUTs */
  float* stars;                             /* Star {altitude, azimuth, radius, intensity}, sorted by grid cell */
  uint32_t stars_offsets[64 * 32 + 1];
  uint32_t stars_count;
  int has_stars;
  float* hdri_color;                        /* HDRI mode: hdri_dim x hdri_dim float4 (sky_compute_hdri), NULL until built */
  uint32_t hdri_dim;
  OrcTexture moon_albedo, moon_normal;      /* data == NULL: absent */
} OrcSky;
void orc_sky_free(OrcSky* sky);
OrcVec3 orc_world_to_sky(const OrcSky* sky, OrcVec3 p);
bool orc_sphere_ray_hit(OrcVec3 ray, OrcVec3 origin, OrcVec3 p, float r);
bool orc_sph_ray_hit_p0(OrcVec3 ray, OrcVec3 origin, float r);
OrcVec3 orc_sample_sphere(OrcVec3 p, float r, OrcVec3 origin, OrcFloat2 random, float* area);
OrcRGB orc_sky_sun_color(const OrcSky* sky, OrcVec3 origin_sky, OrcVec3 ray);
OrcRGB orc_sky_color(const OrcSky* sky, OrcVec3 origin_world, OrcVec3 ray, bool include_sun, float random_offset);
OrcRGB orc_sky_inscattering(const OrcSky* sky, OrcVec3 origin_world, OrcVec3 ray, float t, uint32_t depth, float random_steps, float random_offset,
                            OrcRGB* transmittance_rgb);
OrcRGB orc_sky_color_mode(const OrcSky* sky, uint32_t mode, OrcVec3 origin_world, OrcVec3 ray, bool include_sun, float random_offset);
#define ORC_SKY_EARTH_RADIUS 6371.0f
#define ORC_SKY_SUN_RADIUS 696340.0f

/* scene internals (orc_trace.c) used by the shading units */
typedef struct {
  float lo[3], hi[3];
  uint32_t left;  /* inner: index of left child (right = left + 1); leaf: first prim slot */
  uint32_t count; /* 0 for inner nodes */
} OrcBvhNode;

struct OrcScene {
  uint32_t num_meshes, num_instances, num_materials, num_prims;
  OrcMesh* meshes;
  OrcInstance* instances;
  OrcMaterialPacked* materials;
  uint32_t* instance_prim_offset; /* first flattened prim of each instance */
  float* world;                   /* 9 floats per flattened prim */
  uint32_t* prim_instance;
  uint32_t* prim_tri;
  OrcBvhNode* nodes;
  uint32_t num_nodes;
  uint32_t* prim_order; /* leaf slot -> flattened prim */
  /* lights */
  OrcLightTree light_tree;
  int has_lights;
  OrcBvhNode* light_nodes;
  uint32_t num_light_nodes;
  uint32_t* light_order;
  float* light_world; /* 9 floats per light id */
  /* textures (orc_texture.c) */
  OrcTexture* textures;
  uint32_t num_textures;
  /* procedural sky (orc_sky.c); NULL: LUMINARY_SKY_MODE_DEFAULT renders black (scenes of the constant-colour tests) */
  OrcSky* sky;
  /* LUTs */
  const uint16_t *lut_conductor, *lut_glossy, *lut_dielectric, *lut_dielectric_inv;
};

#endif
