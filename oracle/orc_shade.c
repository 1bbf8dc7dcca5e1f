/*
 * orc_shade.c - oracle: surface shading, BSDF, light-tree NEE and the per-bounce path loop.
 * TEST INFRASTRUCTURE ONLY (see lum_oracle.h).
 *
 * Restates, function by function, the reference's shading kernels for the configuration the hot path covers
 * (no textures, no fog / ocean / particles / clouds, sky either constant colour or off):
 *   geometry_process_tasks        cuda/geometry.cuh:11-180
 *   geometry_get_context          cuda/geometry_utils.cuh:54-221
 *   material params quantisation  cuda/material.cuh:36-44,186-330
 *   BSDF                          cuda/bsdf.cuh:11-301, cuda/bsdf_utils.cuh:79-587, cuda/bsdf_lut.cuh:20-209
 *   light tree / RIS / triangles  cuda/light.cuh:49-159, light_tree.cuh:68-320, ris.cuh:22-157, light_triangle.cuh
 *   BSDF-sampled lights, MIS      cuda/light_bsdf.cuh:24-146, mis.cuh:19-57
 *   NEE evaluation                cuda/direct_lighting.cuh:445-669, optix_anyhit.cuh:49-205
 *   Russian roulette              cuda/directives.cuh:11-32
 *   miss shading                  cuda/sky.cuh:534-633 (constant colour branch)
 *   accumulation                  cuda/accumulation.cuh:36-84
 * The reference builds its CUDA with --use_fast_math; this file uses libm, which is the stated fp32 tolerance
 * of the image-level parity tests. Quirks of the reference that influence results are kept and marked QUIRK.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "lum_oracle.h"
#include "orc_internal.h"

OrcHit orc_bvh_closest(
  const OrcBvhNode* nodes, const uint32_t* order, const float* tris, OrcVec3 origin, OrcVec3 ray, float tmin, float tmax, uint32_t ignore_prim,
  uint64_t* nodes_visited, uint64_t* tris_tested);
void orc_build_bvh_public(const float* tris, uint32_t n, OrcBvhNode** nodes_out, uint32_t* num_nodes_out, uint32_t** order_out);

#define GEOMETRY_DELTA_PATH_CUTOFF 0.05f /* cuda/utils.cuh:45 */
#define BSDF_ROUGHNESS_CLAMP 2e-2f       /* cuda/utils.cuh:46 */
#define RUSSIAN_ROULETTE_CLAMP (1.0f / 8.0f)
#define LIGHT_TREE_NUM_OUTPUTS 8

/* MaterialFlag, device_utils.h:252-259 */
#define MF_TRANSLUCENT 1u
#define MF_REFRACTION_IS_INSIDE 2u
#define MF_METALLIC 4u
#define MF_COLORED_TRANSPARENCY 8u

/* DeviceMaterialFlags, device_structs.h:218-230 */
#define DMF_TRANSLUCENT 0x01
#define DMF_EMISSION 0x02
#define DMF_METALLIC 0x08
#define DMF_COLORED_TRANSPARENCY 0x10
#define DMF_ROUGHNESS_AS_SMOOTHNESS 0x20
#define DMF_NORMAL_MAP_COMPRESSED 0x40
#define DMF_BIDIRECTIONAL_EMISSION 0x80

typedef enum { HINT_GENERAL = 0, HINT_MICROFACET = 1, HINT_DIFFUSE = 2, HINT_REFRACTION = 3 } Hint;

/* ------------------------------------------------------------------ */
/* material                                                             */
/* ------------------------------------------------------------------ */
typedef struct {
  uint8_t flags;
  float roughness_clamp, roughness, refraction_index;
  float albedo[4];
  OrcRGB emission;
  float emission_scale;
} Material;

static float normed_u16(uint32_t v) { return v * (1.0f / 0xFFFF); }

/* load_material, cuda/memory.cuh:442-469 */
static Material load_material(const OrcMaterialPacked* p) {
  Material m;
  m.flags = p->flags;
  /* QUIRK: the reference masks the flags word with 0xFF00 without shifting, so an 8-bit clamp c decodes to c*256/65535 */
  m.roughness_clamp  = normed_u16(((uint32_t) p->roughness_clamp) << 8);
  m.roughness        = normed_u16(p->roughness);
  m.refraction_index = normed_u16(p->refraction_index) * 2.0f + 1.0f;
  m.albedo[0]        = normed_u16(p->albedo_r);
  m.albedo[1]        = normed_u16(p->albedo_g);
  m.albedo[2]        = normed_u16(p->albedo_b);
  m.albedo[3]        = normed_u16(p->albedo_a);
  m.emission_scale   = orc_u2f(((uint32_t) p->emission_scale) << 15);
  m.emission         = c_scale(c_get(normed_u16(p->emission_r), normed_u16(p->emission_g), normed_u16(p->emission_b)), m.emission_scale);
  return m;
}

/* MaterialParams after the set/get round trip of cuda/material.cuh (10/10/10-bit albedo, 8-bit opacity, 10-bit
 * roughness, 8-bit IOR over [0,3], shared-exponent emission). */
typedef struct {
  uint32_t flags;
  OrcRGB albedo;
  float opacity, roughness, ior;
  OrcRGB emission;
} Params;

static float quant_norm(float v, uint32_t bits) {
  const uint32_t maxv = (1u << bits) - 1u;
  const uint32_t q    = (uint32_t) (orc_saturate(v) * maxv + 0.5f);
  return q * (1.0f / maxv);
}

static OrcRGB quant_color(OrcRGB value) { /* material_set_color / material_get_color, MATERIAL_PARAM_TYPE_COLOR */
  uint32_t max_component;
  float max_value, lower, higher;
  if (value.r > value.g && value.r > value.b) {
    max_component = 0, max_value = value.r, lower = value.g, higher = value.b;
  }
  else if (value.g > value.b) {
    max_component = 1, max_value = value.g, lower = value.r, higher = value.b;
  }
  else {
    max_component = 2, max_value = value.b, lower = value.r, higher = value.g;
  }
  max_value = orc_saturate(max_value * (1.0f / 1023.0f)) * 2.0f;
  lower     = orc_saturate(lower * (2.0f / 1023.0f) * (1.0f / max_value));
  higher    = orc_saturate(higher * (2.0f / 1023.0f) * (1.0f / max_value));

  const uint32_t dmax = (orc_f2u(max_value) >= 0x30000000u) ? (orc_f2u(max_value) >> 14) & 0x3FFF : 0;
  const uint32_t dlo  = (uint32_t) (lower * 0xFF + 0.5f);
  const uint32_t dhi  = (uint32_t) (higher * 0xFF + 0.5f);

  const float mv = (dmax > 0) ? orc_u2f((dmax << 14) | 0x30000000u) * (1023.0f / 2.0f) : 0.0f;
  const float lo = dlo * (1.0f / 0xFF) * mv;
  const float hi = dhi * (1.0f / 0xFF) * mv;
  if (max_component == 0)
    return c_get(mv, lo, hi);
  if (max_component == 1)
    return c_get(lo, mv, hi);
  return c_get(lo, hi, mv);
}

typedef struct {
  uint32_t instance_id, tri_id, prim;
  OrcVec3 position, V, normal;
  uint32_t face_normal; /* packed */
  uint16_t state;
  Params params;
} Ctx;

/* ------------------------------------------------------------------ */
/* LUT texture fetch: CUDA linear filtering, normalised coordinates, clamp; R16 unorm                      */
/* ------------------------------------------------------------------ */
static void lin_setup(float x, int n, int* i0, int* i1, float* a) {
  const float xb = x * n - 0.5f;
  const float fl = floorf(xb);
  /* 9-bit fixed point weight with 8 fractional bits (CUDA programming guide, texture fetching) */
  const float fr = floorf((xb - fl) * 256.0f + 0.5f) * (1.0f / 256.0f);
  int a0 = (int) fl, a1 = (int) fl + 1;
  if (a0 < 0)
    a0 = 0;
  if (a1 < 0)
    a1 = 0;
  if (a0 > n - 1)
    a0 = n - 1;
  if (a1 > n - 1)
    a1 = n - 1;
  *i0 = a0, *i1 = a1, *a = fr;
}

static float tex2d(const uint16_t* t, float x, float y) {
  if (!t)
    return 1.0f;
  int x0, x1, y0, y1;
  float a, b;
  lin_setup(x, 32, &x0, &x1, &a);
  lin_setup(y, 32, &y0, &y1, &b);
  const float s = 1.0f / 65535.0f;
  const float v00 = t[x0 + 32 * y0] * s, v10 = t[x1 + 32 * y0] * s, v01 = t[x0 + 32 * y1] * s, v11 = t[x1 + 32 * y1] * s;
  return (1 - a) * (1 - b) * v00 + a * (1 - b) * v10 + (1 - a) * b * v01 + a * b * v11;
}

static float tex3d(const uint16_t* t, float x, float y, float z) {
  if (!t)
    return 1.0f;
  int x0, x1, y0, y1, z0, z1;
  float a, b, c;
  lin_setup(x, 32, &x0, &x1, &a);
  lin_setup(y, 32, &y0, &y1, &b);
  lin_setup(z, 32, &z0, &z1, &c);
  const float s = 1.0f / 65535.0f;
#define T3(X, Y, Z) (t[(X) + 32 * (Y) + 1024 * (Z)] * s)
  const float lo = (1 - a) * (1 - b) * T3(x0, y0, z0) + a * (1 - b) * T3(x1, y0, z0) + (1 - a) * b * T3(x0, y1, z0) + a * b * T3(x1, y1, z0);
  const float hi = (1 - a) * (1 - b) * T3(x0, y0, z1) + a * (1 - b) * T3(x1, y0, z1) + (1 - a) * b * T3(x0, y1, z1) + a * b * T3(x1, y1, z1);
#undef T3
  return (1 - c) * lo + c * hi;
}

/* ------------------------------------------------------------------ */
/* math helpers of cuda/math.cuh                                        */
/* ------------------------------------------------------------------ */
static OrcVec3 reflect_vector(OrcVec3 V, OrcVec3 n) { /* math.cuh:192-197 */
  const float d = v_dot(V, n);
  return v_normalize(v_sub(v_scale(n, 2.0f * d), V));
}

static OrcVec3 refract_vector(OrcVec3 V, OrcVec3 n, float index_ratio, bool* total_reflection) { /* math.cuh:799-819 */
  if (index_ratio < ORC_EPS) {
    *total_reflection = false;
    return v_scale(V, -1.0f);
  }
  const float d = fabsf(v_dot(n, V));
  const float b = 1.0f - index_ratio * index_ratio * (1.0f - d * d);
  *total_reflection = b < 0.0f;
  if (*total_reflection)
    return reflect_vector(V, n);
  return v_normalize(v_sub(v_scale(n, index_ratio * d - sqrtf(b)), v_scale(V, index_ratio)));
}

static OrcQuat rotation_to_z(OrcVec3 v) { /* quaternion_rotation_to_z_canonical, math.cuh:385-409 */
  OrcQuat r;
  if (v.z < -1.0f + ORC_EPS) {
    r.x = 1.0f, r.y = 0.0f, r.z = 0.0f, r.w = 0.0f;
    return r;
  }
  r.x = v.y, r.y = -v.x, r.z = 0.0f, r.w = 1.0f + v.z;
  const float norm = 1.0f / sqrtf(r.x * r.x + r.y * r.y + r.w * r.w);
  r.x *= norm, r.y *= norm, r.w *= norm;
  return r;
}

static OrcQuat quat_inverse(OrcQuat q) {
  OrcQuat r = {-q.x, -q.y, -q.z, q.w};
  return r;
}

static OrcVec3 sample_ray_sphere(float alpha, float beta) { /* math.cuh:339-358 */
  if (fabsf(alpha) > 1.0f - ORC_EPS)
    return v_get(0.0f, 0.0f, copysignf(1.0f, alpha));
  const float a = sqrtf(1.0f - alpha * alpha);
  const float b = 2.0f * ORC_PI * beta;
  return v_get(a * cosf(b), a * sinf(b), alpha);
}

static OrcFloat2 coords_in_triangle(OrcVec3 vertex, OrcVec3 e1, OrcVec3 e2, OrcVec3 point) { /* math.cuh:203-213 */
  const OrcVec3 diff = v_sub(point, vertex);
  const float d00 = v_dot(e1, e1), d01 = v_dot(e1, e2), d11 = v_dot(e2, e2);
  const float d20 = v_dot(diff, e1), d21 = v_dot(diff, e2);
  const float denom = 1.0f / (d00 * d11 - d01 * d01);
  OrcFloat2 r       = {(d11 * d20 - d01 * d21) * denom, (d00 * d21 - d01 * d20) * denom};
  return r;
}

static OrcVec3 normal_adaptation(OrcVec3 V, OrcVec3 shading, OrcVec3 geometry) { /* math.cuh:1547-1569 */
  if (v_dot(shading, geometry) < 0.0f)
    shading = v_scale(shading, -1.0f);
  if (v_dot(V, shading) < 0.0f) {
    const OrcVec3 proj = v_scale(V, v_dot(shading, V));
    return v_normalize(v_sub(shading, v_scale(proj, 1.1f)));
  }
  return shading;
}

/* ------------------------------------------------------------------ */
/* geometry_get_context, cuda/geometry_utils.cuh:54-221                 */
/* ------------------------------------------------------------------ */
static Ctx get_context(const OrcScene* s, uint32_t prim, OrcVec3 hit_point, OrcVec3 ray_world, uint16_t state, uint32_t medium_ior) {
  const uint32_t inst   = s->prim_instance[prim];
  const uint32_t tri    = s->prim_tri[prim];
  const OrcInstance* in = &s->instances[inst];
  const OrcMesh* mesh   = &s->meshes[in->mesh_id];
  const OrcTransform* t = &in->transform;

  const float* vb      = mesh->vertex + 9 * (size_t) tri;
  const OrcVec3 vertex = v_get(vb[0], vb[1], vb[2]);
  const OrcVec3 edge1  = v_sub(v_get(vb[3], vb[4], vb[5]), vertex);
  const OrcVec3 edge2  = v_sub(v_get(vb[6], vb[7], vb[8]), vertex);

  OrcVec3 position  = orc_transform_apply_inv(t, hit_point);
  const OrcVec3 ray = orc_transform_apply_rotation_inv(t, ray_world);

  OrcVec3 face_normal = v_normalize(v_cross(edge1, edge2));
  const OrcFloat2 co  = coords_in_triangle(vertex, edge1, edge2, position);

  position = v_add(vertex, v_add(v_scale(edge1, co.x), v_scale(edge2, co.y)));
  position = orc_transform_apply(t, position);

  const uint16_t material_id = mesh->material[tri];
  const Material mat         = load_material(&s->materials[material_id]);

  /* normals are stored octahedron-packed per vertex (device_structs.c:345-355) */
  const float* nb   = mesh->normal + 9 * (size_t) tri;
  const OrcVec3 n0  = orc_unpack_normal(orc_pack_normal_host(v_get(nb[0], nb[1], nb[2])));
  const OrcVec3 n1  = orc_unpack_normal(orc_pack_normal_host(v_get(nb[3], nb[4], nb[5])));
  const OrcVec3 n2  = orc_unpack_normal(orc_pack_normal_host(v_get(nb[6], nb[7], nb[8])));
  const OrcVec3 en1 = v_sub(n1, n0);
  const OrcVec3 en2 = v_sub(n2, n0);

  const OrcMaterialPacked* mp = &s->materials[material_id];
  const OrcFloat2 tex_coords  = orc_prim_tex_coords(s, prim, co.x, co.y); /* lerp_uv of the bfloat16 vertex uvs */

  /* geometry_compute_normal, geometry_utils.cuh:13-52 */
  const bool is_inside = v_dot(face_normal, ray) > 0.0f;
  if (is_inside)
    face_normal = v_scale(face_normal, -1.0f);
  OrcVec3 normal = v_get(n0.x + co.x * en1.x + co.y * en2.x, n0.y + co.x * en1.y + co.y * en2.y, n0.z + co.x * en1.z + co.y * en2.z);
  {
    const float len = v_len(normal); /* lerp_normals, math.cuh:215-227 */
    normal          = (len < ORC_EPS) ? face_normal : v_scale(normal, 1.0f / len);
  }
  if (mp->normal_tex != 0xFFFF) { /* normal map, geometry_utils.cuh:26-49 */
    const float def[4] = {0.0f, 0.0f, 1.0f, 0.0f};
    float nf[4];
    orc_texture_load(s, mp->normal_tex, tex_coords.x, tex_coords.y, true, false, def, nf);
    OrcVec3 map_normal = v_get(nf[0], nf[1], nf[2]);
    if ((mat.flags & DMF_NORMAL_MAP_COMPRESSED) && orc_texture_valid(s, mp->normal_tex))
      map_normal = v_sub(v_scale(map_normal, 2.0f), v_get(1.0f, 1.0f, 1.0f));
    map_normal      = v_normalize(map_normal);
    const OrcQuat q = rotation_to_z(normal);
    normal          = orc_quat_apply(quat_inverse(q), map_normal);
  }
  normal = normal_adaptation(v_scale(ray, -1.0f), normal, face_normal);

  float albedo[4] = {mat.albedo[0], mat.albedo[1], mat.albedo[2], mat.albedo[3]};
  if (mp->albedo_tex != 0xFFFF) {
    const float def[4] = {0.9f, 0.9f, 0.9f, 1.0f};
    orc_texture_load(s, mp->albedo_tex, tex_coords.x, tex_coords.y, true, true, def, albedo);
  }

  const bool emissive_side    = (!is_inside) || (mat.flags & DMF_BIDIRECTIONAL_EMISSION);
  const bool has_emission     = (mat.flags & DMF_EMISSION) && emissive_side;
  const bool include_emission = has_emission && ((state & ORC_STATE_ALLOW_EMISSION) != 0);
  OrcRGB emission             = include_emission ? mat.emission : c_splat(0.0f);
  if (include_emission && mp->luminance_tex != 0xFFFF) {
    const float def[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float lf[4];
    orc_texture_load(s, mp->luminance_tex, tex_coords.x, tex_coords.y, true, true, def, lf);
    emission = c_scale(c_get(lf[0], lf[1], lf[2]), albedo[3] * mat.emission_scale);
  }

  float roughness = mat.roughness;
  if (mp->roughness_tex != 0xFFFF) {
    const float def[4] = {0.5f, 0.0f, 0.0f, 0.0f};
    float rf[4];
    orc_texture_load(s, mp->roughness_tex, tex_coords.x, tex_coords.y, true, true, def, rf);
    roughness = rf[0];
  }
  if (mat.flags & DMF_ROUGHNESS_AS_SMOOTHNESS)
    roughness = 1.0f - roughness;
  roughness = fmaxf(roughness, BSDF_ROUGHNESS_CLAMP);
  if ((state & ORC_STATE_DELTA_PATH) == 0)
    roughness = fmaxf(roughness, mat.roughness_clamp);

  uint32_t flags = mat.flags & DMF_TRANSLUCENT;
  if (mp->metallic_tex != 0xFFFF) {
    /* the reference leaves metallic textures unimplemented ("TODO: Stochastic filtering", geometry_utils.cuh:160-162):
     * a material with a metallic map is NOT metallic */
  }
  else if (mat.flags & DMF_METALLIC)
    flags |= MF_METALLIC;
  if (mat.flags & DMF_COLORED_TRANSPARENCY)
    flags |= MF_COLORED_TRANSPARENCY;
  if (is_inside)
    flags |= MF_REFRACTION_IS_INSIDE;

  /* medium_stack_ior_peek, medium_stack.cuh:10-16 */
  const uint32_t cior   = (is_inside ? (medium_ior >> 8) : medium_ior) & 0xFF;
  const float other_ior = orc_ior_decompress(cior);
  const float ior_in    = is_inside ? mat.refraction_index : other_ior;
  const float ior_out   = is_inside ? other_ior : mat.refraction_index;

  if ((flags & MF_TRANSLUCENT) && (fabsf(1.0f - ior_in / ior_out) < 1e-4f)) {
    if ((flags & MF_COLORED_TRANSPARENCY) == 0) {
      albedo[0] = 1.0f + albedo[3] * (albedo[0] - 1.0f);
      albedo[1] = 1.0f + albedo[3] * (albedo[1] - 1.0f);
      albedo[2] = 1.0f + albedo[3] * (albedo[2] - 1.0f);
    }
    albedo[3] = 0.0f;
    flags |= MF_COLORED_TRANSPARENCY;
  }

  Ctx ctx;
  ctx.instance_id = inst;
  ctx.tri_id      = tri;
  ctx.prim        = prim;
  ctx.normal      = orc_transform_apply_rotation(t, normal);
  ctx.face_normal = orc_pack_normal(face_normal); /* QUIRK: packed in mesh space, never rotated (geometry_utils.cuh:206) */
  ctx.position    = position;
  ctx.V           = v_scale(ray_world, -1.0f);
  ctx.state       = state;

  ctx.params.flags     = flags;
  ctx.params.albedo    = c_get(quant_norm(albedo[0], 10), quant_norm(albedo[1], 10), quant_norm(albedo[2], 10));
  ctx.params.opacity   = quant_norm(albedo[3], 8);
  ctx.params.roughness = quant_norm(roughness, 10);
  ctx.params.emission  = quant_color(emission);
  ctx.params.ior       = quant_norm((ior_in / ior_out) * (1.0f / 3.0f), 8) * 3.0f;
  return ctx;
}

/* ------------------------------------------------------------------ */
/* BSDF, cuda/bsdf_utils.cuh                                            */
/* ------------------------------------------------------------------ */
typedef struct {
  OrcVec3 V;
  float fresnel_dielectric, NdotH, NdotL, NdotV, HdotL, HdotV;
  bool is_refraction;
} RayCtx;

static float bsdf_fresnel(OrcVec3 normal, OrcVec3 V, OrcVec3 refraction, float ior) { /* :79-96 */
  const float NdotV = v_dot(V, normal);
  const float NdotT = -v_dot(refraction, normal);
  const float s1 = ior * NdotV, s2 = 1.0f * NdotT;
  const float p1 = ior * NdotT, p2 = 1.0f * NdotV;
  float rs = (s1 - s2) / (s1 + s2);
  float rp = (p1 - p2) / (p1 + p2);
  rs *= rs;
  rp *= rp;
  return orc_saturate(0.5f * (rs + rp));
}

static OrcRGB fresnel_schlick(OrcRGB f0, float f90, float HdotV) { /* :105-118 */
  const float o  = 1.0f - fabsf(HdotV);
  const float p2 = o * o;
  const float t  = p2 * p2 * o;
  return c_get(f0.r + (f90 - f0.r) * t, f0.g + (f90 - f0.g) * t, f0.b + (f90 - f0.b) * t);
}

static float shadowed_f90(OrcRGB f0) { return fminf(1.0f, (1.0f / 0.04f) * c_luminance(f0)); }

static OrcVec3 normal_from_pair(OrcVec3 L, OrcVec3 V, float ior) { /* :137-143 */
  const OrcVec3 n = v_add(L, v_scale(V, ior));
  const float len = v_len(n);
  return (len > 0.0f) ? v_scale(n, 1.0f / len) : V;
}

static float smith_g1(float r4, float NdotS) {
  const float n2 = fmaxf(0.0001f, NdotS * NdotS);
  return 2.0f / (sqrtf(((r4 * (1.0f - n2)) + n2) / n2) + 1.0f);
}
static float smith_g2(float r4, float NdotL, float NdotV) {
  const float a = NdotV * sqrtf(r4 + NdotL * (NdotL - r4 * NdotL));
  const float b = NdotL * sqrtf(r4 + NdotV * (NdotV - r4 * NdotV));
  return 0.5f / (a + b);
}
static float smith_g2_over_g1(float r4, float NdotL, float NdotV) {
  const float g1v = smith_g1(r4, NdotV), g1l = smith_g1(r4, NdotL);
  return g1l / (g1v + g1l - g1v * g1l);
}
static float ggx_d(float NdotH, float r4) {
  const float n2 = fminf(NdotH * NdotH, 1.0f);
  const float a  = 1.0f - n2 + r4 * n2;
  return r4 / (ORC_PI * a * a);
}

static OrcVec3 microfacet_sample_normal(OrcVec3 V, float roughness, OrcFloat2 rnd) { /* bounded VNDF, :185-203 */
  const float r2 = roughness * roughness, r4 = r2 * r2;
  const OrcVec3 v   = v_normalize(v_get(r2 * V.x, r2 * V.y, V.z));
  const float phi   = 2.0f * ORC_PI * rnd.x;
  const float s     = 1.0f + sqrtf(V.x * V.x + V.y * V.y);
  const float s2    = s * s;
  const float k     = (1.0f - r4) * s2 / (s2 + r4 * V.z * V.z);
  const float b     = k * v.z;
  const float z     = (1.0f - rnd.y) * (1.0f + b) - b;
  const float st    = sqrtf(orc_saturate(1.0f - z * z));
  const OrcVec3 smp = v_add(v_get(st * cosf(phi), st * sinf(phi), z), v);
  return v_normalize(v_get(smp.x * r2, smp.y * r2, smp.z));
}

static float vndf_norm(OrcVec3 V, float r4, float NdotV) { /* 2 (k NdotV + t), shared by several evaluators */
  const float len2 = r4 * (V.x * V.x + V.y * V.y);
  const float t    = sqrtf(len2 + V.z * V.z);
  const float s    = 1.0f + sqrtf(V.x * V.x + V.y * V.y);
  const float s2   = s * s;
  const float k    = (1.0f - r4) * s2 / (s2 + r4 * V.z * V.z);
  return 2.0f * (k * NdotV + t);
}

static float microfacet_pdf(OrcVec3 V, float roughness, float NdotH, float NdotV) { /* :205-220 */
  const float r2 = roughness * roughness, r4 = r2 * r2;
  return ggx_d(NdotH, r4) / vndf_norm(V, r4, NdotV);
}

static float microfacet_evaluate(float roughness, float NdotH, float NdotL, float NdotV) { /* :227-237 */
  const float r2 = roughness * roughness, r4 = r2 * r2;
  return ggx_d(NdotH, r4) * smith_g2(r4, NdotL, NdotV) * NdotL;
}
static float microfacet_eval_sampled_microfacet(OrcVec3 V, float roughness, float NdotL, float NdotV) { /* :239-257 */
  const float r2 = roughness * roughness, r4 = r2 * r2;
  return vndf_norm(V, r4, NdotV) * smith_g2(r4, NdotL, NdotV) * NdotL;
}
static float microfacet_eval_sampled_diffuse(float roughness, float NdotH, float NdotL, float NdotV) { /* :259-269 */
  const float r2 = roughness * roughness, r4 = r2 * r2;
  return ggx_d(NdotH, r4) * smith_g2(r4, NdotL, NdotV) * ORC_PI;
}

static OrcVec3 refraction_sample_normal(OrcVec3 V, float roughness, OrcFloat2 rnd) { /* spherical caps, :276-288 */
  const float r2    = roughness * roughness;
  const OrcVec3 v   = v_normalize(v_get(r2 * V.x, r2 * V.y, V.z));
  const float phi   = 2.0f * ORC_PI * rnd.x;
  const float z     = (1.0f - rnd.y) * (1.0f + v.z) - v.z;
  const float st    = sqrtf(orc_saturate(1.0f - z * z));
  const OrcVec3 smp = v_add(v_get(st * cosf(phi), st * sinf(phi), z), v);
  return v_normalize(v_get(smp.x * r2, smp.y * r2, smp.z));
}

static float refraction_pdf(float roughness, float NdotH, float NdotV, float NdotL, float HdotV, float HdotL, float ior) { /* :290-305 */
  (void) NdotL;
  const float r2 = roughness * roughness, r4 = r2 * r2;
  float den = ior * HdotV + HdotL;
  den       = den * den;
  return ggx_d(NdotH, r4) * smith_g1(r4, NdotV) * (HdotV / NdotV) * (HdotL / den);
}

static float refraction_evaluate(float roughness, float HdotL, float HdotV, float NdotH, float NdotL, float NdotV, float ior) { /* :313-328 */
  const float r2 = roughness * roughness, r4 = r2 * r2;
  float den = ior * HdotV + HdotL;
  den       = den * den;
  return 4.0f * NdotL * HdotV * HdotL * ggx_d(NdotH, r4) * smith_g2(r4, NdotL, NdotV) / den;
}

static float diffuse_pdf(float NdotL) { return orc_saturate(NdotL) * (1.0f / ORC_PI); }

static float diffuse_eval_sampled_microfacet(OrcVec3 V, float roughness, float NdotL, float NdotH, float NdotV) { /* :362-377 */
  const float r2 = roughness * roughness, r4 = r2 * r2;
  return NdotL * vndf_norm(V, r4, NdotV) / (ORC_PI * ggx_d(NdotH, r4));
}

static float conductor_albedo(const OrcScene* s, float NdotV, float roughness) { return tex2d(s->lut_conductor, NdotV, roughness); }
static float glossy_albedo(const OrcScene* s, float NdotV, float roughness) { return tex2d(s->lut_glossy, NdotV, roughness); }
static float dielectric_albedo(const OrcScene* s, float NdotV, float roughness, float ior) { /* :495-505 */
  const bool use_inv = ior > 1.0f;
  const float coord  = use_inv ? (ior - 1.0f) * 0.5f : (1.0f / ior - 1.0f) * 0.5f;
  return tex3d(use_inv ? s->lut_dielectric_inv : s->lut_dielectric, NdotV, roughness, coord);
}

static float ss_term_for(Hint hint, const Params* p, const RayCtx* c, float one_over_pdf, float ior_quirk) {
  const float r = p->roughness;
  switch (hint) {
    case HINT_GENERAL:
      return microfacet_evaluate(r, c->NdotH, c->NdotL, c->NdotV) * one_over_pdf;
    case HINT_MICROFACET:
      return microfacet_eval_sampled_microfacet(c->V, r, c->NdotL, c->NdotV);
    case HINT_DIFFUSE:
      return microfacet_eval_sampled_diffuse(r, c->NdotH, c->NdotL, c->NdotV);
    default:
      return microfacet_evaluate(r, c->NdotH, c->NdotL, c->NdotV) / refraction_pdf(r, c->NdotH, c->NdotV, c->NdotL, c->HdotV, c->HdotL, ior_quirk);
  }
}

static OrcRGB bsdf_conductor(const OrcScene* s, const Params* p, const RayCtx* c, Hint hint, float one_over_pdf) { /* :383-428 */
  if (c->NdotL <= 0.0f || c->NdotV <= 0.0f || (p->flags & MF_TRANSLUCENT) || (p->flags & MF_METALLIC) == 0)
    return c_splat(0.0f);
  /* QUIRK: for the refraction hint the reference reads the roughness into `ior` */
  const float ior  = (hint == HINT_REFRACTION) ? p->roughness : 1.0f;
  const float ss   = ss_term_for(hint, p, c, one_over_pdf, ior);
  const float da   = conductor_albedo(s, c->NdotV, p->roughness);
  const OrcRGB f0  = p->albedo;
  const OrcRGB fr  = fresnel_schlick(f0, shadowed_f90(f0), c->HdotV);
  const OrcRGB ssf = c_scale(fr, ss);
  const OrcRGB msf = c_mul(f0, c_scale(fr, ((1.0f / da) - 1.0f) * ss));
  return c_add(ssf, msf);
}

static OrcRGB bsdf_glossy(const OrcScene* s, const Params* p, const RayCtx* c, Hint hint, float one_over_pdf) { /* :434-493 */
  if (c->NdotL <= 0.0f || c->NdotV <= 0.0f || (p->flags & MF_TRANSLUCENT) || (p->flags & MF_METALLIC) != 0)
    return c_splat(0.0f);
  const float ior = (hint == HINT_REFRACTION) ? p->roughness : 1.0f;
  const float r   = p->roughness;
  const float ss  = ss_term_for(hint, p, c, one_over_pdf, ior);
  float diff;
  switch (hint) {
    case HINT_GENERAL:
      diff = diffuse_pdf(c->NdotL) * one_over_pdf;
      break;
    case HINT_DIFFUSE:
      diff = 1.0f;
      break;
    case HINT_MICROFACET:
      diff = diffuse_eval_sampled_microfacet(c->V, r, c->NdotL, c->NdotH, c->NdotV);
      break;
    default:
      diff = diffuse_pdf(c->NdotL) / refraction_pdf(r, c->NdotH, c->NdotV, c->NdotL, c->HdotV, c->HdotL, ior);
      break;
  }
  const float cda  = conductor_albedo(s, c->NdotV, r);
  const float gda  = glossy_albedo(s, c->NdotV, r);
  const OrcRGB f0  = c_splat(0.04f);
  const OrcRGB fr  = fresnel_schlick(f0, shadowed_f90(f0), c->HdotV);
  const OrcRGB ssf = c_scale(fr, ss / cda);
  const OrcRGB dif = c_scale(p->albedo, diff * (1.0f - gda));
  return c_add(ssf, dif);
}

static OrcRGB bsdf_dielectric(const OrcScene* s, const Params* p, const RayCtx* c, Hint hint, float one_over_pdf) { /* :507-572 */
  if (c->NdotL <= 0.0f || c->NdotV <= 0.0f || (p->flags & MF_TRANSLUCENT) == 0)
    return c_splat(0.0f);
  /* QUIRK: `ior` is read from the ROUGHNESS parameter in the reference (bsdf_utils.cuh:516) */
  const float ior = p->roughness;
  const float r   = p->roughness;
  float term      = 0.0f;
  if (c->is_refraction) {
    switch (hint) {
      case HINT_GENERAL:
        term = refraction_evaluate(r, c->HdotL, c->HdotV, c->NdotH, c->NdotL, c->NdotV, ior) * one_over_pdf;
        break;
      case HINT_REFRACTION:
        term = smith_g2_over_g1(r * r * r * r, c->NdotL, c->NdotV);
        break;
      default:
        term = 0.0f;
        break;
    }
    term *= (1.0f - c->fresnel_dielectric);
  }
  else {
    switch (hint) {
      case HINT_GENERAL:
        term = microfacet_evaluate(r, c->NdotH, c->NdotL, c->NdotV) * one_over_pdf;
        break;
      case HINT_MICROFACET:
        term = microfacet_eval_sampled_microfacet(c->V, r, c->NdotL, c->NdotV);
        break;
      case HINT_DIFFUSE: /* QUIRK: missing break in the reference, falls through to the refraction case */
      default:
        term = microfacet_evaluate(r, c->NdotH, c->NdotL, c->NdotV) / refraction_pdf(r, c->NdotH, c->NdotV, c->NdotL, c->HdotV, c->HdotL, ior);
        break;
    }
    term *= c->fresnel_dielectric;
  }
  term /= dielectric_albedo(s, c->NdotV, r, ior);
  if (ior == 1.0f && c->is_refraction)
    term = (hint == HINT_REFRACTION) ? 1.0f : 0.0f;
  return c_scale(p->albedo, term);
}

static OrcRGB bsdf_multiscattering(const OrcScene* s, const Params* p, const RayCtx* c, Hint hint, float one_over_pdf) { /* :578-587 */
  if (c->is_refraction)
    return c_scale(bsdf_dielectric(s, p, c, hint, one_over_pdf), p->opacity);
  const OrcRGB a = bsdf_conductor(s, p, c, hint, one_over_pdf);
  const OrcRGB b = bsdf_glossy(s, p, c, hint, one_over_pdf);
  const OrcRGB d = bsdf_dielectric(s, p, c, hint, one_over_pdf);
  return c_scale(c_add(c_add(a, b), d), p->opacity);
}

/* bsdf_evaluate_analyze, cuda/bsdf.cuh:11-50 */
static RayCtx evaluate_analyze(const Params* p, OrcVec3 normal, OrcVec3 V, OrcVec3 L) {
  RayCtx c;
  c.NdotL         = v_dot(normal, L);
  c.NdotV         = orc_saturate(v_dot(normal, V));
  c.is_refraction = c.NdotL < 0.0f;
  c.NdotL         = c.is_refraction ? -c.NdotL : c.NdotL;
  const float ior = p->ior;
  OrcVec3 refraction_vector, H;
  bool total_reflection;
  if (c.is_refraction) {
    total_reflection  = false;
    H                 = normal_from_pair(L, V, ior);
    refraction_vector = L;
  }
  else {
    H                 = normal_from_pair(L, V, 1.0f);
    refraction_vector = refract_vector(V, H, ior, &total_reflection);
  }
  c.HdotV = fabsf(v_dot(H, V));
  c.HdotL = fabsf(v_dot(H, L));
  c.NdotH = v_dot(normal, H);
  if (c.NdotH < 0.0f) {
    H       = v_scale(H, -1.0f);
    c.NdotH = -c.NdotH;
  }
  c.fresnel_dielectric = total_reflection ? 1.0f : bsdf_fresnel(H, V, refraction_vector, ior);
  c.V                  = V;
  return c;
}

/* bsdf_evaluate_core, cuda/bsdf.cuh:52-65 */
static OrcRGB evaluate_core(const OrcScene* s, const Params* p, const RayCtx* c, Hint hint, OrcVec3 L, OrcVec3 face_normal, float one_over_pdf) {
  const float fndl = v_dot(face_normal, L);
  const float flip = c->is_refraction ? -1.0f : 1.0f;
  if (fndl * flip < ORC_EPS)
    return c_splat(0.0f);
  return bsdf_multiscattering(s, p, c, hint, one_over_pdf);
}

static OrcRGB bsdf_evaluate(const OrcScene* s, const Ctx* ctx, OrcVec3 L, Hint hint, bool* is_refraction, float one_over_pdf) { /* :73-85 */
  const RayCtx c = evaluate_analyze(&ctx->params, ctx->normal, ctx->V, L);
  *is_refraction = c.is_refraction;
  return evaluate_core(s, &ctx->params, &c, hint, L, orc_unpack_normal(ctx->face_normal), one_over_pdf);
}

/* bsdf_sample_context, cuda/bsdf.cuh:103-133 */
static RayCtx sample_context(const Params* p, OrcVec3 normal, OrcVec3 V, OrcVec3 H, OrcVec3 L, bool is_refraction) {
  RayCtx c;
  c.NdotL         = v_dot(normal, L);
  c.NdotV         = orc_saturate(v_dot(normal, V));
  c.is_refraction = is_refraction;
  c.NdotL         = is_refraction ? -c.NdotL : c.NdotL;
  const float ior = p->ior;
  bool total_reflection         = false;
  const OrcVec3 refraction_vec  = is_refraction ? L : refract_vector(V, H, ior, &total_reflection);
  c.HdotV                       = fabsf(v_dot(H, V));
  c.HdotL                       = fabsf(v_dot(H, L));
  c.NdotH                       = v_dot(normal, H);
  float flip                    = 1.0f;
  if (c.NdotH < 0.0f) {
    flip    = -1.0f;
    c.NdotH = -c.NdotH;
  }
  c.fresnel_dielectric = total_reflection ? 1.0f : bsdf_fresnel(v_scale(H, flip), V, refraction_vec, ior);
  c.V                  = V;
  return c;
}

typedef struct {
  OrcVec3 ray;
  OrcRGB weight;
  bool is_transparent_pass, is_microfacet_based;
} SampleInfo;

/* bsdf_sample<MATERIAL_GEOMETRY>, cuda/bsdf.cuh:135-301; set = RandomSet::BSDF<set> */
static SampleInfo bsdf_sample(const OrcScene* s, const Ctx* ctx, OrcPathID pid, uint32_t depth, uint32_t set) {
  const Params* p = &ctx->params;
  SampleInfo info;

  if (p->opacity < 1.0f) {
    const float tr = orc_random_1d(ORC_RT_BSDF_OPACITY + set, pid, depth);
    if (tr > p->opacity) {
      info.ray                 = v_scale(ctx->V, -1.0f);
      info.weight              = (p->flags & MF_COLORED_TRANSPARENCY) ? p->albedo : c_splat(1.0f);
      info.is_microfacet_based = false;
      info.is_transparent_pass = true;
      return info;
    }
  }

  const OrcQuat rot      = rotation_to_z(ctx->normal);
  const OrcVec3 V_local  = orc_quat_apply(rot, ctx->V);
  const OrcVec3 fn_local = orc_quat_apply(rot, orc_unpack_normal(ctx->face_normal));
  const OrcVec3 up       = v_get(0.0f, 0.0f, 1.0f);

  info.is_transparent_pass = false;
  info.is_microfacet_based = false;

  const bool translucent        = (p->flags & MF_TRANSLUCENT) != 0;
  const bool include_diffuse    = !translucent && ((p->flags & MF_METALLIC) == 0);
  const bool include_refraction = translucent;

  float sum_weights    = 0.0f;
  OrcRGB selected_eval = c_splat(0.0f);
  OrcVec3 ray_local    = up;
  float resampling     = orc_random_1d(ORC_RT_BSDF_RESAMPLING + set, pid, depth);
  const float ior      = p->ior;
  const float rough    = p->roughness;

  { /* microfacet reflection */
    const OrcVec3 m  = microfacet_sample_normal(V_local, rough, orc_random_2d(ORC_RT_BSDF_REFLECTION + set, pid, depth));
    const OrcVec3 r  = reflect_vector(V_local, m);
    const RayCtx c   = sample_context(p, up, V_local, m, r, false);
    const OrcRGB ev  = evaluate_core(s, p, &c, HINT_MICROFACET, r, fn_local, 1.0f);
    const float pdf  = microfacet_pdf(V_local, rough, c.NdotH, c.NdotV);
    const float dpdf = include_diffuse ? diffuse_pdf(c.NdotL) : 0.0f;
    const float rpdf = include_refraction ? refraction_pdf(rough, c.NdotH, c.NdotV, c.NdotL, c.HdotV, c.HdotL, ior) : 0.0f;
    const float sum  = pdf + dpdf + rpdf;
    const float mis  = (sum > 0.0f) ? pdf / sum : 0.0f;
    const float w    = c_importance(ev) * mis;
    ray_local                = r;
    sum_weights              = w;
    selected_eval            = ev;
    info.is_microfacet_based = true;
  }

  if (include_diffuse) {
    const OrcFloat2 rnd = orc_random_2d(ORC_RT_BSDF_DIFFUSE + set, pid, depth);
    const OrcVec3 r     = sample_ray_sphere(rnd.x, rnd.y);
    const OrcVec3 m     = v_normalize(v_add(V_local, r));
    const RayCtx c      = sample_context(p, up, V_local, m, r, false);
    const OrcRGB ev     = evaluate_core(s, p, &c, HINT_DIFFUSE, r, fn_local, 1.0f);
    const float pdf     = diffuse_pdf(c.NdotL);
    const float mpdf    = microfacet_pdf(V_local, rough, c.NdotH, c.NdotV);
    const float rpdf    = include_refraction ? refraction_pdf(rough, c.NdotH, c.NdotV, c.NdotL, c.HdotV, c.HdotL, ior) : 0.0f;
    const float sum     = pdf + mpdf + rpdf;
    const float mis     = (sum > 0.0f) ? pdf / sum : 0.0f;
    const float w       = c_importance(ev) * mis;
    sum_weights += w;
    const float prob = w / sum_weights;
    if (resampling < prob) {
      ray_local                = r;
      selected_eval            = ev;
      info.is_transparent_pass = false;
      info.is_microfacet_based = false;
      resampling               = orc_random_saturate(resampling / prob);
    }
    else {
      resampling = orc_random_saturate((resampling - prob) / (1.0f - prob));
    }
  }

  if (include_refraction) {
    bool total_reflection;
    const OrcVec3 m = refraction_sample_normal(V_local, rough, orc_random_2d(ORC_RT_BSDF_REFRACTION + set, pid, depth));
    const OrcVec3 r = refract_vector(V_local, m, ior, &total_reflection);
    const RayCtx c  = sample_context(p, up, V_local, m, r, !total_reflection);
    const OrcRGB ev = evaluate_core(s, p, &c, HINT_REFRACTION, r, fn_local, 1.0f);
    float mis       = 1.0f;
    if (total_reflection) {
      const float pdf  = refraction_pdf(rough, c.NdotH, c.NdotV, c.NdotL, c.HdotV, c.HdotL, ior);
      const float rpdf = microfacet_pdf(V_local, rough, c.NdotH, c.NdotV);
      const float dpdf = include_diffuse ? diffuse_pdf(c.NdotL) : 0.0f;
      const float sum  = pdf + rpdf + dpdf;
      mis              = (sum > 0.0f) ? pdf / sum : 0.0f;
    }
    const float w = c_importance(ev) * mis;
    sum_weights += w;
    const float prob = w / sum_weights;
    if (resampling < prob) {
      ray_local                = r;
      selected_eval            = ev;
      info.is_transparent_pass = !total_reflection;
      info.is_microfacet_based = true;
      resampling               = orc_random_saturate(resampling / prob);
    }
    else {
      resampling = orc_random_saturate((resampling - prob) / (1.0f - prob));
    }
  }

  info.weight = (sum_weights > 0.0f) ? c_scale(selected_eval, sum_weights / c_importance(selected_eval)) : c_splat(0.0f);
  info.ray    = v_normalize(orc_quat_apply(quat_inverse(rot), ray_local));
  return info;
}

/* ------------------------------------------------------------------ */
/* RIS, cuda/ris.cuh:22-157                                             */
/* ------------------------------------------------------------------ */
typedef struct {
  float sum_weight, selected_target, random;
} Reservoir;

static Reservoir reservoir_init(float random) {
  Reservoir r = {0.0f, 0.0f, random};
  return r;
}

static bool reservoir_add(Reservoir* r, float target, float sampling_weight) {
  const float weight = target * sampling_weight;
  r->sum_weight += weight;
  if (weight == 0.0f)
    return false;
  const float prob     = weight / r->sum_weight;
  const bool accepted  = r->random < prob;
  r->selected_target   = accepted ? target : r->selected_target;
  const float shift    = accepted ? 0.0f : prob;
  const float scale    = accepted ? prob : 1.0f - prob;
  r->random            = orc_random_saturate((r->random - shift) / scale);
  return accepted;
}

static float reservoir_weight(const Reservoir* r) { return (r->selected_target > 0.0f) ? r->sum_weight / r->selected_target : 0.0f; }

/* ------------------------------------------------------------------ */
/* light tree, cuda/light_tree.cuh                                      */
/* ------------------------------------------------------------------ */
#pragma pack(push, 1)
typedef struct { /* DeviceLightTreeRootHeader, device_utils.h:305-318 */
  uint16_t x, y, z, num_root_lights, power_normalization;
  uint8_t num_sections, padding1;
  int8_t exp_x, exp_y, exp_z, exp_std_dev;
} RootHeader;
typedef struct { /* DeviceLightTreeRootSection, :320-327 */
  uint8_t rel_mean_x[8], rel_mean_y[8], rel_mean_z[8], rel_std_dev[8];
  uint16_t rel_power[8];
} RootSection;
typedef struct { /* DeviceLightTreeNode, :283-303 */
  uint16_t x, y, z, padding;
  int8_t exp_x, exp_y, exp_z, exp_std_dev;
  uint8_t num_lights, padding1;
  uint16_t padding2;
  uint32_t child_ptr, light_ptr;
  uint8_t rel_mean_x[8], rel_mean_y[8], rel_mean_z[8], rel_std_dev[8], rel_power[8];
} TreeNode;
#pragma pack(pop)

static float bfloat_unpack(uint16_t v) { return orc_u2f(((uint32_t) v) << 16); }

static float tree_importance(const Ctx* ctx, float power, OrcVec3 mean, float std_dev) { /* :68-89 */
  const OrcVec3 PO  = v_sub(mean, ctx->position);
  const float d2    = v_dot(PO, PO);
  const float var   = std_dev * std_dev;
  const float inv   = 1.0f / (d2 + var);
  float result      = power * inv;
  if (ctx->params.flags & MF_TRANSLUCENT)
    return result;
  const float t     = var * inv;
  const float NdotL = orc_saturate(v_dot(PO, ctx->normal) * sqrtf(inv));
  return result * (NdotL * (1.0f - t) + t);
}

static float child_importance(const Ctx* ctx, float power, float rel_std, float mx, float my, float mz, OrcVec3 base, OrcVec3 ex, float exp_v) {
  if (power == 0.0f)
    return 0.0f;
  const float std_dev = rel_std * exp_v;
  const OrcVec3 mean  = v_add(v_mul(v_get(mx, my, mz), ex), base);
  return fmaxf(tree_importance(ctx, power, mean, std_dev), 0.0f);
}

typedef struct {
  uint32_t is_light, child_index, probability; /* LightTreeContinuation: 1 / 8 / 20 bits */
} Continuation;

typedef struct {
  Continuation data[LIGHT_TREE_NUM_OUTPUTS];
  float root_sum;
} TreeWork;

static TreeWork tree_prepass(const OrcScene* s, const Ctx* ctx, OrcPathID pid, uint32_t depth) { /* :191-262 */
  const RootHeader* header = (const RootHeader*) s->light_tree.root;
  const RootSection* sections = (const RootSection*) ((const uint8_t*) s->light_tree.root + 16);

  float agg_sum = 0.0f;
  float lane_target[LIGHT_TREE_NUM_OUTPUTS], lane_random[LIGHT_TREE_NUM_OUTPUTS];
  uint8_t selected[LIGHT_TREE_NUM_OUTPUTS];
  for (uint32_t l = 0; l < LIGHT_TREE_NUM_OUTPUTS; l++) {
    lane_random[l] = orc_random_1d(ORC_RT_LIGHT_GEO_TREE_PREPASS + l, pid, depth);
    lane_target[l] = 0.0f;
    selected[l]    = 0;
  }

  const OrcVec3 base = v_get(bfloat_unpack(header->x), bfloat_unpack(header->y), bfloat_unpack(header->z));
  const OrcVec3 ex   = v_get(exp2f(header->exp_x), exp2f(header->exp_y), exp2f(header->exp_z));
  const float exp_v  = exp2f(header->exp_std_dev);
  float sum          = 0.0f;

  for (uint32_t sec = 0; sec < header->num_sections; sec++) {
    const RootSection* S = &sections[sec];
    for (uint32_t c = 0; c < 8; c++) {
      const float target =
        child_importance(ctx, (float) S->rel_power[c], S->rel_std_dev[c], S->rel_mean_x[c], S->rel_mean_y[c], S->rel_mean_z[c], base, ex, exp_v);
      /* ris_aggregator_add_sample, ris.cuh:114-124 */
      agg_sum += target;
      const float prob = (target > 0.0f) ? target / agg_sum : 0.0f;
      if (prob == 0.0f)
        continue;
      sum += target;
      for (uint32_t l = 0; l < LIGHT_TREE_NUM_OUTPUTS; l++) { /* ris_lane_add_sample, ris.cuh:140-151 */
        const bool accepted = lane_random[l] < prob;
        lane_target[l]      = accepted ? target : lane_target[l];
        const float shift   = accepted ? 0.0f : prob;
        const float scale   = accepted ? prob : 1.0f - prob;
        lane_random[l]      = orc_random_saturate((lane_random[l] - shift) / scale);
        if (accepted)
          selected[l] = (uint8_t) (sec * 8 + c);
      }
    }
  }

  TreeWork work;
  work.root_sum = sum * (bfloat_unpack(header->power_normalization) / 0xFFFF);
  for (uint32_t l = 0; l < LIGHT_TREE_NUM_OUTPUTS; l++) {
    const bool is_light = selected[l] < header->num_root_lights;
    const uint8_t index = is_light ? selected[l] : (uint8_t) (selected[l] - header->num_root_lights);
    const float prob    = (agg_sum > 0.0f) ? lane_target[l] / agg_sum : 0.0f;
    work.data[l].is_light    = is_light;
    work.data[l].child_index = index;
    uint32_t q               = 0;
    if (prob > 0.0f) {
      q = (uint32_t) ((0xFFFFF * prob) + 0.5f);
      if (q < 1)
        q = 1;
    }
    work.data[l].probability = q & 0xFFFFF;
  }
  return work;
}

static void tree_postpass(const OrcScene* s, const Ctx* ctx, OrcPathID pid, uint32_t depth, uint32_t lane, const TreeWork* work, uint32_t* light_id,
                          float* weight) { /* :264-320 */
  const Continuation c = work->data[lane];
  const float prob     = c.probability * (1.0f / 0xFFFFF) * LIGHT_TREE_NUM_OUTPUTS;
  *light_id            = 0xFFFFFFFFu;
  *weight              = (prob > 0.0f) ? 1.0f / prob : 0.0f;
  if (prob == 0.0f)
    return;
  if (c.is_light) {
    *light_id = c.child_index;
    return;
  }
  const TreeNode* nodes = (const TreeNode*) s->light_tree.nodes;
  const TreeNode* node  = &nodes[c.child_index];
  Reservoir res         = reservoir_init(orc_random_1d(ORC_RT_LIGHT_GEO_TREE_POSTPASS + lane, pid, depth));

  while (*light_id == 0xFFFFFFFFu) {
    const OrcVec3 base = v_get(bfloat_unpack(node->x), bfloat_unpack(node->y), bfloat_unpack(node->z));
    const OrcVec3 ex   = v_get(exp2f(node->exp_x), exp2f(node->exp_y), exp2f(node->exp_z));
    const float exp_v  = exp2f(node->exp_std_dev);
    uint8_t sel        = 0xFF;
    for (uint32_t k = 0; k < 8; k++) {
      const float target = child_importance(ctx, (float) node->rel_power[k], node->rel_std_dev[k], node->rel_mean_x[k], node->rel_mean_y[k],
                                            node->rel_mean_z[k], base, ex, exp_v);
      if (reservoir_add(&res, target, 1.0f))
        sel = (uint8_t) k;
    }
    if (sel == 0xFF)
      break;
    *weight *= reservoir_weight(&res);
    if (sel < node->num_lights) {
      *light_id = node->light_ptr + sel;
      break;
    }
    node           = &nodes[node->child_ptr + (sel - node->num_lights)];
    res.sum_weight = 0.0f;
    res.selected_target = 0.0f;
  }
}

/* ------------------------------------------------------------------ */
/* triangle lights, cuda/light_triangle.cuh                             */
/* ------------------------------------------------------------------ */
typedef struct {
  OrcVec3 vertex, edge1, edge2;
  uint16_t material_id;
  bool bidirectional;
  uint32_t prim;        /* flattened primitive of the emitter (for its vertex uvs) */
  OrcFloat2 tex_coords; /* set by light_intersect (light_triangle_sample_finalize_dist_and_uvs, :74-90) */
} TriLight;

static TriLight light_init(const OrcScene* s, uint32_t light_id) { /* :33-72 with light_tree_get_light, light_tree.cuh:322-328 */
  const uint32_t inst   = s->light_tree.tri_handle_map[2 * light_id + 0];
  const uint32_t tri    = s->light_tree.tri_handle_map[2 * light_id + 1];
  const OrcInstance* in = &s->instances[inst];
  const OrcMesh* mesh   = &s->meshes[in->mesh_id];
  const float* vb       = mesh->vertex + 9 * (size_t) tri;
  const OrcVec3 v0      = v_get(vb[0], vb[1], vb[2]);
  const OrcVec3 e1      = v_sub(v_get(vb[3], vb[4], vb[5]), v0);
  const OrcVec3 e2      = v_sub(v_get(vb[6], vb[7], vb[8]), v0);
  TriLight L;
  L.vertex        = orc_transform_apply(&in->transform, v0);
  L.edge1         = orc_transform_apply_relative(&in->transform, e1);
  L.edge2         = orc_transform_apply_relative(&in->transform, e2);
  L.material_id   = mesh->material[tri];
  L.bidirectional = (s->materials[L.material_id].flags & DMF_BIDIRECTIONAL_EMISSION) != 0;
  L.prim          = s->instance_prim_offset[inst] + tri;
  L.tex_coords.x = L.tex_coords.y = 0.0f;
  return L;
}

static float light_intersect_uv(const TriLight* L, OrcVec3 origin, OrcVec3 ray, float* cu, float* cv) { /* light_triangle_intersection_uv, :10-31 */
  const float v9[9] = {L->vertex.x, L->vertex.y, L->vertex.z, L->vertex.x + L->edge1.x, L->vertex.y + L->edge1.y, L->vertex.z + L->edge1.z,
                       L->vertex.x + L->edge2.x, L->vertex.y + L->edge2.y, L->vertex.z + L->edge2.z};
  /* the reference feeds vertex/edge1/edge2 directly; orc_tri_mt recomputes the edges from v9, which would
   * round differently, so the arithmetic is repeated here on the edges */
  (void) v9;
  const OrcVec3 h = v_cross(ray, L->edge2);
  const float a   = v_dot(L->edge1, h);
  const float f   = 1.0f / a;
  const OrcVec3 sv = v_sub(origin, L->vertex);
  const float u   = f * v_dot(sv, h);
  const OrcVec3 q = v_cross(sv, L->edge1);
  const float v   = f * v_dot(ray, q);
  *cu = u, *cv = v;
  if (v < 0.0f || u < 0.0f || !(u + v <= 1.0f))
    return ORC_FLT_MAX;
  const float t = f * v_dot(L->edge2, q);
  return (t >= 0.0f) ? t : ORC_FLT_MAX;
}

/* light_triangle_sample_finalize_dist_and_uvs, light_triangle.cuh:74-90: distance + texture coordinates of the hit point */
static float light_intersect(const OrcScene* s, TriLight* L, OrcVec3 origin, OrcVec3 ray) {
  float cu, cv;
  const float dist = light_intersect_uv(L, origin, ray, &cu, &cv);
  if (dist != ORC_FLT_MAX && s->num_textures)
    L->tex_coords = orc_prim_tex_coords(s, L->prim, cu, cv);
  return dist;
}

static float light_solid_angle(const TriLight* L, OrcVec3 origin) { /* :92-106 */
  const OrcVec3 v0 = v_normalize(v_sub(L->vertex, origin));
  const OrcVec3 v1 = v_normalize(v_sub(v_add(L->vertex, L->edge1), origin));
  const OrcVec3 v2 = v_normalize(v_sub(v_add(L->vertex, L->edge2), origin));
  const float G0   = fabsf(v_dot(v_cross(v0, v1), v2));
  const float G1   = v_dot(v0, v2) + v_dot(v1, v2);
  const float G2   = 1.0f + v_dot(v0, v1);
  return 2.0f * atan2f(G0, G1 + G2);
}

static float light_area(const TriLight* L) { return v_len(v_cross(L->edge1, L->edge2)) * 0.5f; }

static bool is_non_finite(float a) { return isnan(a) || isinf(a); }

static bool light_sample_solid_angle(const TriLight* L, OrcVec3 origin, OrcFloat2 rnd, OrcVec3* ray, float* solid_angle) { /* :112-158 */
  const OrcVec3 v0 = v_normalize(v_sub(L->vertex, origin));
  const OrcVec3 v1 = v_normalize(v_sub(v_add(L->vertex, L->edge1), origin));
  const OrcVec3 v2 = v_normalize(v_sub(v_add(L->vertex, L->edge2), origin));
  const float G0s  = v_dot(v_cross(v0, v1), v2);
  if (!L->bidirectional && G0s >= 0.0f)
    return false;
  const float G0 = fabsf(G0s);
  const float G1 = v_dot(v0, v2) + v_dot(v1, v2);
  const float G2 = 1.0f + v_dot(v0, v1);
  *solid_angle   = 2.0f * atan2f(G0, G1 + G2);
  if (is_non_finite(*solid_angle) || *solid_angle < 1e-7f)
    return false;
  const float ssa = rnd.x * *solid_angle;
  const OrcVec3 r = v_add(v_scale(v0, G0 * cosf(0.5f * ssa) - G1 * sinf(0.5f * ssa)), v_scale(v2, G2 * sinf(0.5f * ssa)));
  const OrcVec3 v2t = v_sub(v_scale(r, 2.0f * v_dot(v0, r) / v_dot(r, r)), v0);
  const float s2 = v_dot(v1, v2t);
  const float sv = (1.0f - rnd.y) + rnd.y * s2;
  const float t  = sqrtf(fmaxf((1.0f - sv * sv) / (1.0f - s2 * s2), 0.0f));
  *ray           = v_normalize(v_add(v_scale(v1, sv - t * s2), v_scale(v2t, t)));
  if (is_non_finite(ray->x) || is_non_finite(ray->y) || is_non_finite(ray->z))
    return false;
  return true;
}

static OrcRGB light_color_of(const OrcScene* s, const TriLight* L) { /* light_get_color, :244-280 */
  const OrcMaterialPacked* mp = &s->materials[L->material_id];
  const Material m            = load_material(mp);
  OrcRGB c                    = m.emission;
  if (mp->luminance_tex != 0xFFFF) {
    const float def[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float e[4];
    orc_texture_load(s, mp->luminance_tex, L->tex_coords.x, L->tex_coords.y, true, true, def, e);
    c = c_scale(c_get(e[0], e[1], e[2]), m.emission_scale);
  }
  if (c_any(c)) {
    float alpha = m.albedo[3];
    if (mp->albedo_tex != 0xFFFF) {
      const float def[4] = {0.0f, 0.0f, 0.0f, 1.0f};
      float a[4];
      orc_texture_load(s, mp->albedo_tex, L->tex_coords.x, L->tex_coords.y, true, true, def, a);
      alpha = a[3];
    }
    c = c_scale(c, alpha);
  }
  return c;
}

/* ------------------------------------------------------------------ */
/* light BSDF sampling + MIS, cuda/light_bsdf.cuh, mis.cuh              */
/* ------------------------------------------------------------------ */
static float lerpf(float a, float b, float t) { return a + t * (b - a); }
static float remap01(float v, float lo, float hi) { return orc_saturate((v - lo) / (hi - lo) * (1.0f - 0.0f) + 0.0f); }
static float lbsdf_sampling_roughness(float r) { return lerpf(r, 1.0f, 0.04f); }
static float lbsdf_rr_probability(float r) { return remap01(r, 0.5f, 0.1f); }

typedef struct {
  OrcVec3 ray;
  OrcRGB weight;
  float sampling_probability;
} LightBsdfSample;

static LightBsdfSample light_bsdf_get_sample(const OrcScene* s, const Ctx* ctx, OrcPathID pid, uint32_t depth) { /* light_bsdf.cuh:24-104 */
  const Params* p        = &ctx->params;
  const OrcQuat rot      = rotation_to_z(ctx->normal);
  const OrcVec3 V_local  = orc_quat_apply(rot, ctx->V);
  const OrcVec3 fn_local = orc_quat_apply(rot, orc_unpack_normal(ctx->face_normal));
  const OrcVec3 up       = v_get(0.0f, 0.0f, 1.0f);
  const bool include_refraction = (p->flags & MF_TRANSLUCENT) != 0;
  const uint32_t num_techniques = 1 + (include_refraction ? 1 : 0);
  const float refraction_prob   = include_refraction ? 1.0f / num_techniques : 0.0f;
  const float choice            = orc_random_1d(ORC_RT_LIGHT_BSDF_CHOICE, pid, depth);
  const uint32_t technique_id   = (uint32_t) (choice * num_techniques);
  const bool use_refraction     = (technique_id == 1 && include_refraction);
  const float roughness         = p->roughness;
  const float rr_random         = orc_random_1d(ORC_RT_LIGHT_BSDF_RR, pid, depth);
  const float rr_prob           = lbsdf_rr_probability(roughness);

  LightBsdfSample out;
  out.ray    = up;
  out.weight = c_splat(0.0f);
  if (rr_random >= rr_prob) {
    out.sampling_probability = 0.0f;
    return out;
  }
  const float sr = lbsdf_sampling_roughness(roughness);
  if (!use_refraction) {
    const OrcVec3 m = microfacet_sample_normal(V_local, sr, orc_random_2d(ORC_RT_LIGHT_BSDF_DIRECTION, pid, depth));
    const OrcVec3 r = reflect_vector(V_local, m);
    const RayCtx c  = sample_context(p, up, V_local, m, r, false);
    const float pdf = microfacet_pdf(V_local, sr, c.NdotH, c.NdotV);
    out.weight      = evaluate_core(s, p, &c, HINT_GENERAL, r, fn_local, 1.0f / pdf);
    out.ray         = r;
    out.sampling_probability = (1.0f - refraction_prob) * pdf;
  }
  else {
    bool total_reflection;
    const OrcVec3 m = refraction_sample_normal(V_local, sr, orc_random_2d(ORC_RT_LIGHT_BSDF_DIRECTION, pid, depth));
    const OrcVec3 r = refract_vector(V_local, m, p->ior, &total_reflection);
    const RayCtx c  = sample_context(p, up, V_local, m, r, !total_reflection);
    const float pdf = refraction_pdf(sr, c.NdotH, c.NdotV, c.NdotL, c.HdotV, c.HdotL, p->ior);
    out.weight      = evaluate_core(s, p, &c, HINT_GENERAL, r, fn_local, 1.0f / pdf);
    out.ray         = r;
    out.sampling_probability = refraction_prob * pdf;
  }
  out.weight = c_scale(out.weight, 1.0f / rr_prob);
  out.sampling_probability *= rr_prob;
  out.ray = v_normalize(orc_quat_apply(quat_inverse(rot), out.ray));
  return out;
}

static float light_bsdf_get_probability(const Ctx* ctx, OrcVec3 L) { /* light_bsdf.cuh:106-146 */
  const Params* p       = &ctx->params;
  const OrcQuat rot     = rotation_to_z(ctx->normal);
  const OrcVec3 V_local = v_normalize(orc_quat_apply(rot, ctx->V));
  const OrcVec3 L_local = v_normalize(orc_quat_apply(rot, L));
  const bool include_refraction = (p->flags & MF_TRANSLUCENT) != 0;
  const uint32_t num_techniques = 1 + (include_refraction ? 1 : 0);
  const float refraction_prob   = include_refraction ? 1.0f / num_techniques : 0.0f;
  const RayCtx c  = evaluate_analyze(p, v_get(0.0f, 0.0f, 1.0f), V_local, L_local);
  const float sr  = lbsdf_sampling_roughness(p->roughness);
  float prob;
  if (c.is_refraction)
    prob = refraction_prob * refraction_pdf(sr, c.NdotH, c.NdotV, c.NdotL, c.HdotV, c.HdotL, p->ior);
  else
    prob = (1.0f - refraction_prob) * microfacet_pdf(V_local, sr, c.NdotH, c.NdotV);
  return prob * lbsdf_rr_probability(p->roughness);
}

/* ------------------------------------------------------------------ */
/* sun NEE: direct_lighting_sun_create_task / direct_lighting_sun_direct, direct_lighting.cuh:21-120, 353-383              */
/* ------------------------------------------------------------------ */
static OrcVec3 bsdf_sample_for_sun(const Ctx* ctx, OrcPathID pid, uint32_t depth) { /* bsdf.cuh:379-403 */
  const Params* p       = &ctx->params;
  const OrcQuat rot     = rotation_to_z(ctx->normal);
  const OrcVec3 V_local = orc_quat_apply(rot, ctx->V);
  /* bsdf_sample_for_light_probabilities, bsdf.cuh:355-374 */
  const float reflection_weight = 1.0f, refraction_weight = (p->flags & MF_TRANSLUCENT) ? 1.0f : 0.0f;
  const float reflection_prob   = reflection_weight / (reflection_weight + refraction_weight);
  const float random_method     = orc_random_1d(ORC_RT_LIGHT_SUN_BSDF_METHOD, pid, depth);
  OrcVec3 ray_local;
  if (random_method < reflection_prob) {
    const OrcVec3 m = microfacet_sample_normal(V_local, p->roughness, orc_random_2d(ORC_RT_LIGHT_SUN_BSDF, pid, depth));
    ray_local       = reflect_vector(V_local, m);
  }
  else {
    bool total_reflection;
    const OrcVec3 m = refraction_sample_normal(V_local, p->roughness, orc_random_2d(ORC_RT_LIGHT_SUN_BSDF, pid, depth));
    ray_local       = refract_vector(V_local, m, p->ior, &total_reflection);
  }
  return v_normalize(orc_quat_apply(quat_inverse(rot), ray_local));
}

static float bsdf_sample_for_sun_pdf(const Ctx* ctx, OrcVec3 L) { /* bsdf.cuh:438-458; NOTE: the world-space V enters the reflection pdf */
  const Params* p = &ctx->params;
  const float reflection_weight = 1.0f, refraction_weight = (p->flags & MF_TRANSLUCENT) ? 1.0f : 0.0f;
  const float reflection_prob   = reflection_weight / (reflection_weight + refraction_weight);
  const float refraction_prob   = refraction_weight / (reflection_weight + refraction_weight);
  const RayCtx c = evaluate_analyze(p, ctx->normal, ctx->V, L);
  if (c.is_refraction)
    return refraction_prob * refraction_pdf(p->roughness, c.NdotH, c.NdotV, c.NdotL, c.HdotV, c.HdotL, p->ior);
  return reflection_prob * microfacet_pdf(ctx->V, p->roughness, c.NdotH, c.NdotV);
}

static void sun_create_task(const OrcScene* s, const Ctx* ctx, OrcPathID pid, uint32_t depth, OrcUint2* color_out, OrcUint2* ray_out) {
  const OrcSky* sky = s->sky;
  color_out->x = color_out->y = 0; /* PACKED_RECORD_BLACK */
  ray_out->x = ray_out->y = 0;
  const OrcVec3 sky_pos        = orc_world_to_sky(sky, ctx->position);
  const bool sun_below_horizon = orc_sph_ray_hit_p0(v_normalize(v_sub(sky->sun_pos, sky_pos)), sky_pos, ORC_SKY_EARTH_RADIUS);
  const bool inside_earth      = v_len(sky_pos) < ORC_SKY_EARTH_RADIUS;
  if (sun_below_horizon || inside_earth)
    return;

  /* BSDF importance sample */
  const OrcVec3 dir_bsdf = bsdf_sample_for_sun(ctx, pid, depth);
  OrcRGB light_bsdf      = c_splat(0.0f);
  bool is_refr;
  if (orc_sphere_ray_hit(dir_bsdf, sky_pos, sky->sun_pos, ORC_SKY_SUN_RADIUS)) {
    light_bsdf = orc_sky_sun_color(sky, sky_pos, dir_bsdf);
    light_bsdf = c_mul(light_bsdf, bsdf_evaluate(s, ctx, dir_bsdf, HINT_GENERAL, &is_refr, 1.0f));
  }
  /* solid angle sample */
  const OrcFloat2 random = orc_random_2d(ORC_RT_LIGHT_SUN_RAY, pid, depth);
  float solid_angle;
  const OrcVec3 dir_solid_angle = orc_sample_sphere(sky->sun_pos, ORC_SKY_SUN_RADIUS, sky_pos, random, &solid_angle);
  OrcRGB light_solid_angle      = orc_sky_sun_color(sky, sky_pos, dir_solid_angle);
  light_solid_angle             = c_mul(light_solid_angle, bsdf_evaluate(s, ctx, dir_solid_angle, HINT_GENERAL, &is_refr, 1.0f));

  /* resampled importance sampling */
  const float target_pdf_bsdf        = c_importance(light_bsdf);
  const float target_pdf_solid_angle = c_importance(light_solid_angle);
  const float mis_weight_bsdf        = solid_angle / (bsdf_sample_for_sun_pdf(ctx, dir_bsdf) * solid_angle + 1.0f);
  const float mis_weight_solid_angle = solid_angle / (bsdf_sample_for_sun_pdf(ctx, dir_solid_angle) * solid_angle + 1.0f);
  const float weight_bsdf            = target_pdf_bsdf * mis_weight_bsdf;
  const float weight_solid_angle     = target_pdf_solid_angle * mis_weight_solid_angle;
  const float sum_weights            = weight_bsdf + weight_solid_angle;
  if (sum_weights == 0.0f)
    return;
  float target_pdf;
  OrcVec3 dir;
  OrcRGB light_color;
  if (orc_random_1d(ORC_RT_LIGHT_SUN_RESAMPLING, pid, depth) * sum_weights < weight_bsdf)
    dir = dir_bsdf, target_pdf = target_pdf_bsdf, light_color = light_bsdf;
  else
    dir = dir_solid_angle, target_pdf = target_pdf_solid_angle, light_color = light_solid_angle;
  light_color = c_scale(light_color, sum_weights / target_pdf);
  if (target_pdf == 0.0f)
    return;
  if (c_importance(light_color) == 0.0f)
    return;
  /* volume_integrate_transmittance: VOLUME_TYPE_NONE on this path */
  *color_out = orc_record_pack(light_color);
  *ray_out   = orc_ray_pack(dir);
}

static float mis_weight_base(float gi_pdf, float solid_angle, float power, float dist_sq, float root_sum) { /* mis.cuh:19-24 */
  const float dl_pdf = LIGHT_TREE_NUM_OUTPUTS * (1.0f / solid_angle) * (power / dist_sq) * (1.0f / root_sum);
  return (dl_pdf > 0.0f) ? gi_pdf / (gi_pdf + dl_pdf) : 1.0f;
}

/* ------------------------------------------------------------------ */
/* shadow / enumeration rays                                            */
/* ------------------------------------------------------------------ */
static void shadow_response(const OrcScene* s, uint32_t prim, float bu, float bv, const OrcMaterialPacked* m, float* rgb, bool* opaque) { /* optix_anyhit.cuh:49-93 */
  float al[4];
  orc_shadow_albedo(s, prim, bu, bv, al);
  const float a      = al[3];
  const bool colored = (m->flags & DMF_COLORED_TRANSPARENCY) != 0;
  *opaque            = false;
  if (a == 1.0f) {
    *opaque = true;
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
  }
  else if (a == 0.0f && !colored) {
    rgb[0] = rgb[1] = rgb[2] = 1.0f;
  }
  else {
    const float tr = 1.0f - a;
    rgb[0]         = colored ? al[0] * tr : tr;
    rgb[1]         = colored ? al[1] * tr : tr;
    rgb[2]         = colored ? al[2] * tr : tr;
  }
}

typedef struct {
  const OrcScene* s;
  uint32_t ignore_prim, target_prim;
  float limit;
  float vis[3];
  bool blocked;
} ShadowState;

/* generic "visit every triangle hit in [tmin, tmax)" traversal over the scene BVH2 */
bool orc_tri_watertight(const float* v9, OrcVec3 origin, OrcVec3 ray, float* t, float* u, float* v);

static OrcRGB shadow_visibility(const OrcScene* s, OrcVec3 origin, OrcVec3 ray, float tmin, float limit, uint32_t ignore_prim, uint32_t target_prim,
                                uint64_t* counter) {
  if (counter)
    (*counter)++;
  float vis[3]       = {1.0f, 1.0f, 1.0f};
  const float o[3]   = {origin.x, origin.y, origin.z};
  const float inv[3] = {1.0f / ray.x, 1.0f / ray.y, 1.0f / ray.z};
  uint32_t stack[128];
  int sp      = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    const OrcBvhNode* n = &s->nodes[stack[--sp]];
    float tn = tmin, tf = limit;
    for (int k = 0; k < 3; k++) {
      const float t0 = (n->lo[k] - o[k]) * inv[k];
      const float t1 = (n->hi[k] - o[k]) * inv[k];
      tn             = fmaxf(tn, fminf(t0, t1));
      tf             = fminf(tf, fmaxf(t0, t1));
    }
    if (!(tn <= tf * 1.0000004f))
      continue;
    if (n->count) {
      for (uint32_t i = 0; i < n->count; i++) {
        const uint32_t p = s->prim_order[n->left + i];
        if (p == ignore_prim || p == target_prim)
          continue;
        float t, u, v;
        if (!orc_tri_watertight(s->world + 9 * (size_t) p, origin, ray, &t, &u, &v))
          continue;
        if (!(t >= tmin) || !(t < limit))
          continue;
        const uint32_t inst = s->prim_instance[p];
        const uint16_t mid  = s->meshes[s->instances[inst].mesh_id].material[s->prim_tri[p]];
        float rgb[3];
        bool opaque;
        shadow_response(s, p, u, v, &s->materials[mid], rgb, &opaque);
        if (opaque)
          return c_splat(0.0f);
        vis[0] *= rgb[0], vis[1] *= rgb[1], vis[2] *= rgb[2];
      }
    }
    else {
      stack[sp++] = n->left;
      stack[sp++] = n->left + 1;
    }
  }
  return c_get(vis[0], vis[1], vis[2]);
}

/* light_bsdf_trace any-hit (optix_anyhit.cuh:145-205). OptiX calls any-hit programs in traversal order, which is
 * unspecified; this oracle (and the product) fix the order to ascending hit distance, ties by light id. */
typedef struct {
  float t, u, v;
  uint32_t light;
} LightHit;

static int cmp_light_hit(const void* a, const void* b) {
  const LightHit* x = (const LightHit*) a;
  const LightHit* y = (const LightHit*) b;
  if (x->t < y->t)
    return -1;
  if (x->t > y->t)
    return 1;
  return (x->light < y->light) ? -1 : (x->light > y->light);
}

static uint32_t enumerate_lights(const OrcScene* s, OrcVec3 origin, OrcVec3 ray, uint32_t ignore_prim, float random, uint32_t* num_hits_out) {
  LightHit hits[64];
  int nh             = 0;
  const float o[3]   = {origin.x, origin.y, origin.z};
  const float inv[3] = {1.0f / ray.x, 1.0f / ray.y, 1.0f / ray.z};
  uint32_t stack[128];
  int sp      = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    const OrcBvhNode* n = &s->light_nodes[stack[--sp]];
    float tn = ORC_EPS, tf = ORC_FLT_MAX;
    for (int k = 0; k < 3; k++) {
      const float t0 = (n->lo[k] - o[k]) * inv[k];
      const float t1 = (n->hi[k] - o[k]) * inv[k];
      tn             = fmaxf(tn, fminf(t0, t1));
      tf             = fminf(tf, fmaxf(t0, t1));
    }
    if (!(tn <= tf * 1.0000004f))
      continue;
    if (n->count) {
      for (uint32_t i = 0; i < n->count; i++) {
        const uint32_t l = s->light_order[n->left + i];
        float t, u, v;
        if (!orc_tri_watertight(s->light_world + 9 * (size_t) l, origin, ray, &t, &u, &v))
          continue;
        if (!(t >= ORC_EPS))
          continue;
        if (nh < 64) {
          hits[nh].t     = t;
          hits[nh].u     = u;
          hits[nh].v     = v;
          hits[nh].light = l;
          nh++;
        }
      }
    }
    else {
      stack[sp++] = n->left;
      stack[sp++] = n->left + 1;
    }
  }
  qsort(hits, nh, sizeof(LightHit), cmp_light_hit);

  uint32_t num_hits = 0, selected = ORC_LIGHT_ID_INVALID;
  for (int i = 0; i < nh; i++) {
    const uint32_t l    = hits[i].light;
    const uint32_t inst = s->light_tree.tri_handle_map[2 * l + 0];
    const uint32_t tri  = s->light_tree.tri_handle_map[2 * l + 1];
    const uint32_t prim = s->instance_prim_offset[inst] + tri;
    if (prim == ignore_prim)
      continue;
    const OrcMaterialPacked* m = &s->materials[s->meshes[s->instances[inst].mesh_id].material[tri]];
    float al[4];
    orc_shadow_albedo(s, prim, hits[i].u, hits[i].v, al); /* optix_get_albedo_for_shadowing with the light-GAS barycentrics */
    const float a              = al[3];
    const bool colored         = (m->flags & DMF_COLORED_TRANSPARENCY) != 0;
    if (a == 0.0f && !colored)
      continue;
    num_hits++;
    bool accepted = true;
    if (num_hits > 1) {
      const float prob  = 1.0f / num_hits;
      accepted          = random < prob;
      const float shift = accepted ? 0.0f : prob;
      const float scale = accepted ? prob : 1.0f - prob;
      random            = orc_random_saturate((random - shift) / scale);
    }
    if (accepted)
      selected = l;
    if (a == 1.0f)
      break; /* opaque emitter culls everything behind it */
  }
  *num_hits_out = num_hits;
  return selected;
}

/* ------------------------------------------------------------------ */
/* scene setters                                                        */
/* ------------------------------------------------------------------ */
void orc_scene_set_light_tree(OrcScene* s, const OrcLightTree* tree) {
  free(s->light_nodes);
  free(s->light_order);
  free(s->light_world);
  s->light_nodes = NULL, s->light_order = NULL, s->light_world = NULL;
  s->has_lights = 0;
  if (!tree || tree->num_lights == 0)
    return;
  s->light_tree  = *tree;
  s->has_lights  = 1;
  s->light_world = (float*) malloc(sizeof(float) * 9 * tree->num_lights);
  for (uint32_t l = 0; l < tree->num_lights; l++) {
    const uint32_t prim = s->instance_prim_offset[tree->tri_handle_map[2 * l]] + tree->tri_handle_map[2 * l + 1];
    memcpy(s->light_world + 9 * (size_t) l, s->world + 9 * (size_t) prim, sizeof(float) * 9);
  }
  orc_build_bvh_public(s->light_world, tree->num_lights, &s->light_nodes, &s->num_light_nodes, &s->light_order);
}

void orc_scene_set_bsdf_luts(OrcScene* s, const uint16_t* conductor, const uint16_t* glossy, const uint16_t* dielectric, const uint16_t* dielectric_inv) {
  s->lut_conductor      = conductor;
  s->lut_glossy         = glossy;
  s->lut_dielectric     = dielectric;
  s->lut_dielectric_inv = dielectric_inv;
}

/* ------------------------------------------------------------------ */
/* BSDF LUT generation, cuda/bsdf_lut.cuh:20-209                        */
/* ------------------------------------------------------------------ */
static uint16_t lut_quantise(float sum) { return (uint16_t) (1 + (uint16_t) (ceilf(orc_saturate(sum) * 0xFFFE))); }

void orc_bsdf_lut_generate(uint16_t* conductor, uint16_t* glossy, uint16_t* dielectric, uint16_t* dielectric_inv, uint32_t iterations, int num_threads,
                           int with_dielectric) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#endif
#pragma omp parallel for schedule(dynamic, 8)
  for (int id = 0; id < 32 * 32; id++) {
    const uint32_t y      = id / 32;
    const uint32_t x      = id - y * 32;
    const float NdotV     = fmaxf(32.0f * ORC_EPS, x * (1.0f / 31));
    const float roughness = y * (1.0f / 31);
    const OrcVec3 V       = v_normalize(v_get(0.0f, sqrtf(1.0f - NdotV * NdotV), NdotV));
    float sum = 0.0f, sum_g = 0.0f;
    const OrcRGB f0 = c_splat(0.04f);
    for (uint32_t i = 0; i < iterations; i++) {
      const OrcPathID pid = orc_path_id_get(0, 0, i);
      const OrcVec3 H     = microfacet_sample_normal(V, roughness, orc_random_2d(ORC_RT_BSDF_REFLECTION, pid, 0));
      const OrcVec3 R     = reflect_vector(V, H);
      const float NdotL   = R.z;
      if (NdotL > 0.0f) {
        /* The reference is built with nvcc's default -fmad=true: every `sum += a * b` of these loops is ONE fused multiply-add
         * (seen in the SASS of oracle/_ref/librefdev.so), and the serial chain of 65 536 of them has a rounding bias that depends on
         * exactly this: written out here with fmaf(). */
        const float r2 = roughness * roughness, r4 = r2 * r2;
        const float pre = vndf_norm(V, r4, NdotV) * smith_g2(r4, NdotL, NdotV);
        sum             = fmaf(pre, NdotL, sum);
        const OrcRGB fr = fresnel_schlick(f0, shadowed_f90(f0), fabsf(v_dot(H, V)));
        sum_g           = fmaf(pre * NdotL, c_luminance(fr), sum_g);
      }
    }
    sum /= iterations;
    sum_g /= iterations;
    conductor[id]    = lut_quantise(sum);
    const float ss   = conductor[id] * (1.0f / 0xFFFF);
    glossy[id]       = lut_quantise(sum_g / ss);
  }
  if (!with_dielectric)
    return;
#pragma omp parallel for schedule(dynamic, 8)
  for (int id = 0; id < 32 * 32 * 32; id++) {
    const uint32_t z      = id / (32 * 32);
    const uint32_t y      = (id - z * 32 * 32) / 32;
    const uint32_t x      = id - y * 32 - z * 32 * 32;
    const float NdotV     = fmaxf(32.0f * ORC_EPS, x * (1.0f / 31));
    const float roughness = y * (1.0f / 31);
    const float ior       = 1.0f + z * (1.0f / 31) * 2.0f;
    const OrcVec3 V       = v_normalize(v_get(0.0f, sqrtf(1.0f - NdotV * NdotV), NdotV));
    for (int pass = 0; pass < 2; pass++) {
      /* pass 0: ratio 1/ior -> dst; pass 1: ratio ior -> dst_inv. The loops use the unquantised ratio. */
      const float ratio = (pass == 0) ? 1.0f / ior : ior;
      float sum         = 0.0f;
      for (uint32_t i = 0; i < iterations; i++) {
        const OrcPathID pid = orc_path_id_get(0, 0, i);
        bool tot;
        OrcVec3 H        = microfacet_sample_normal(V, roughness, orc_random_2d(ORC_RT_BSDF_REFLECTION, pid, 0));
        OrcVec3 refl     = reflect_vector(V, H);
        OrcVec3 refr     = refract_vector(V, H, ratio, &tot);
        float fresnel    = tot ? 1.0f : bsdf_fresnel(H, V, refr, ratio);
        const float NdotL = refl.z;
        if (NdotL > 0.0f)
          sum = fmaf(microfacet_eval_sampled_microfacet(V, roughness, NdotL, NdotV), fresnel, sum);

        H        = refraction_sample_normal(V, roughness, orc_random_2d(ORC_RT_BSDF_REFRACTION, pid, 0));
        refr     = refract_vector(V, H, ratio, &tot);
        /* QUIRK: total reflection counts as fresnel 1 in the first table and 0 in the second (bsdf_lut.cuh:146,186) */
        fresnel  = tot ? ((pass == 0) ? 1.0f : 0.0f) : bsdf_fresnel(H, V, refr, ratio);
        const float HdotV = fabsf(v_dot(H, V));
        (void) HdotV;
        const float NdotR = -refr.z;
        if (NdotR > 0.0f) {
          const float r4 = roughness * roughness * roughness * roughness;
          sum            = fmaf(smith_g2_over_g1(r4, NdotR, NdotV), 1.0f - fresnel, sum);
        }
      }
      sum /= iterations;
      if (pass == 0)
        dielectric[id] = lut_quantise(sum);
      else
        dielectric_inv[id] = lut_quantise(sum);
    }
  }
}

/* single dielectric texel (both tables) for spot checks; same arithmetic as the loop above */
void orc_bsdf_lut_dielectric_texel(uint32_t id, uint32_t iterations, uint16_t* out, uint16_t* out_inv) {
  const uint32_t z      = id / (32 * 32);
  const uint32_t y      = (id - z * 32 * 32) / 32;
  const uint32_t x      = id - y * 32 - z * 32 * 32;
  const float NdotV     = fmaxf(32.0f * ORC_EPS, x * (1.0f / 31));
  const float roughness = y * (1.0f / 31);
  const float ior       = 1.0f + z * (1.0f / 31) * 2.0f;
  const OrcVec3 V       = v_normalize(v_get(0.0f, sqrtf(1.0f - NdotV * NdotV), NdotV));
  for (int pass = 0; pass < 2; pass++) {
    const float ratio = (pass == 0) ? 1.0f / ior : ior;
    float sum         = 0.0f;
    for (uint32_t i = 0; i < iterations; i++) {
      const OrcPathID pid = orc_path_id_get(0, 0, i);
      bool tot;
      OrcVec3 H         = microfacet_sample_normal(V, roughness, orc_random_2d(ORC_RT_BSDF_REFLECTION, pid, 0));
      OrcVec3 refl      = reflect_vector(V, H);
      OrcVec3 refr      = refract_vector(V, H, ratio, &tot);
      float fresnel     = tot ? 1.0f : bsdf_fresnel(H, V, refr, ratio);
      if (refl.z > 0.0f)
        sum = fmaf(microfacet_eval_sampled_microfacet(V, roughness, refl.z, NdotV), fresnel, sum);
      H       = refraction_sample_normal(V, roughness, orc_random_2d(ORC_RT_BSDF_REFRACTION, pid, 0));
      refr    = refract_vector(V, H, ratio, &tot);
      fresnel = tot ? ((pass == 0) ? 1.0f : 0.0f) : bsdf_fresnel(H, V, refr, ratio);
      const float NdotR = -refr.z;
      if (NdotR > 0.0f) {
        const float r4 = roughness * roughness * roughness * roughness;
        sum            = fmaf(smith_g2_over_g1(r4, NdotR, NdotV), 1.0f - fresnel, sum);
      }
    }
    sum /= iterations;
    if (pass == 0)
      *out = lut_quantise(sum);
    else
      *out_inv = lut_quantise(sum);
  }
}

/* ------------------------------------------------------------------ */
/* the path loop                                                        */
/* ------------------------------------------------------------------ */
static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int g_dbg_x = -1, g_dbg_y = -1;
void orc_set_debug_pixel(int x, int y) { g_dbg_x = x, g_dbg_y = y; }
#include <stdio.h>
#define DBG(...) do { if ((int) x == g_dbg_x && (int) y == g_dbg_y) { printf(__VA_ARGS__); } } while (0)

/* One call of geometry_process_tasks' loop body (cuda/geometry.cuh:20-177) for a single task: everything the reference kernel
 * writes for this path vertex (NEE tasks, emission, bounce task), BEFORE any shadow / enumeration ray is traced. */
static void shade_vertex(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, OrcPathID pid, uint32_t depth, uint16_t state,
                         OrcVec3 origin, OrcVec3 ray, uint32_t prim, float t, OrcUint2 record, uint32_t medium_ior, OrcVertexOut* out) {
  const bool sky_on = set->sky_mode != 0; /* direct_lighting_ambient_is_allowed: sky.mode != DEFAULT */
  const OrcRGB sky  = (set->sky_mode == 2) ? set->sky_constant_color : c_splat(0.0f);
  memset(out, 0, sizeof(*out));
  out->geo_light_id = ORC_LIGHT_ID_INVALID;

  const OrcVec3 hit_point = v_add(origin, v_scale(ray, t));
  const Ctx ctx           = get_context(s, prim, hit_point, ray, state, medium_ior);
  const OrcRGB rec_in     = orc_record_unpack(record);
  out->hit_point          = hit_point;

  float root_sum = 0.0f;
  if (s->has_lights) {
    /* direct_lighting_geometry_create_task -> light_sample, light.cuh:144-159 */
    const TreeWork work = tree_prepass(s, &ctx, pid, depth);
    root_sum            = work.root_sum;
    Reservoir res       = reservoir_init(orc_random_1d(ORC_RT_LIGHT_GEO_RESAMPLING, pid, depth));
    uint32_t sel_light  = ORC_LIGHT_ID_INVALID;
    OrcVec3 sel_ray     = v_get(0, 0, 1);
    OrcRGB sel_color    = c_splat(0.0f);
    float sel_dist      = 0.0f;
    for (uint32_t lane = 0; lane < LIGHT_TREE_NUM_OUTPUTS; lane++) {
      uint32_t light_id;
      float tree_weight;
      tree_postpass(s, &ctx, pid, depth, lane, &work, &light_id, &tree_weight);
      if (light_id == ORC_LIGHT_ID_INVALID)
        continue;
      const uint32_t linst = s->light_tree.tri_handle_map[2 * light_id], ltri = s->light_tree.tri_handle_map[2 * light_id + 1];
      if (linst == ctx.instance_id && ltri == ctx.tri_id)
        continue;
      TriLight L = light_init(s, light_id);
      /* light_evaluate_candidate, light.cuh:49-83 */
      const OrcFloat2 rr = orc_random_2d(ORC_RT_LIGHT_GEO_RAY + lane, pid, depth);
      OrcVec3 lray;
      float solid_angle;
      if (!light_sample_solid_angle(&L, ctx.position, rr, &lray, &solid_angle))
        continue;
      const float dist = light_intersect(s, &L, ctx.position, lray);
      if (dist == ORC_FLT_MAX)
        continue;
      OrcRGB lcol = light_color_of(s, &L);
      bool is_refr;
      const OrcRGB bw = bsdf_evaluate(s, &ctx, lray, HINT_GENERAL, &is_refr, 1.0f);
      /* mis_compute_weight_dl, mis.cuh:46-57 */
      const float power  = c_importance(lcol) * light_area(&L);
      const float gi_pdf = light_bsdf_get_probability(&ctx, lray);
      const float mis    = 1.0f - mis_weight_base(gi_pdf, solid_angle, power, dist * dist, root_sum);
      lcol               = c_scale(c_mul(lcol, bw), mis);
      if (reservoir_add(&res, c_importance(lcol), tree_weight * solid_angle)) {
        sel_light = light_id;
        sel_ray   = lray;
        sel_color = lcol;
        sel_dist  = dist;
      }
    }
    out->geo_light_id = sel_light;
    out->geo_color    = c_scale(sel_color, reservoir_weight(&res));
    out->geo_ray      = sel_ray;
    out->geo_dist     = sel_dist;

    /* direct_lighting_bsdf_create_task, direct_lighting.cuh:425-443 */
    const LightBsdfSample bs = light_bsdf_get_sample(s, &ctx, pid, depth);
    out->bsdf_weight         = bs.weight;
    out->bsdf_ray            = bs.ray;
    out->bsdf_prob           = bs.sampling_probability;
    out->bsdf_root_sum       = root_sum;
  }

  /* direct_lighting_sun_is_allowed: sky.mode != CONSTANT_COLOR (geometry.cuh:57-65); a scene without a sky renders mode 0 black */
  if (set->sky_mode != 2 && s->sky)
    sun_create_task(s, &ctx, pid, depth, &out->sun_color, &out->sun_ray);

  /* bounce sampling */
  const SampleInfo bounce = bsdf_sample(s, &ctx, pid, depth, 0);

  /* ambient NEE task, direct_lighting.cuh:382-401: allowed whenever the sky is not the procedural one; its colour is
   * sky_color_no_compute(position, ray, state = 0), sky.cuh:534-565: the constant colour, or the HDRI table without the sun's disc */
  if (sky_on) {
    OrcRGB amb_sky = sky;
    if (set->sky_mode == 1 && s->sky && s->sky->hdri_color)
      amb_sky = orc_sky_color_mode(s->sky, 1, ctx.position, bounce.ray, false, 0.0f);
    out->amb_color = orc_record_pack(c_mul(amb_sky, bounce.weight));
    out->amb_ray   = orc_ray_pack(bounce.ray);
    out->amb_valid = 1;
  }

  /* delta-path bookkeeping, geometry.cuh:80-97 */
  bool is_delta;
  if (bounce.is_transparent_pass) {
    const float ior   = ctx.params.ior;
    const float scale = (ior >= 1.0f) ? ior : 1.0f / ior;
    is_delta          = ctx.params.roughness * fminf(scale - 1.0f, 1.0f) <= GEOMETRY_DELTA_PATH_CUTOFF;
  }
  else {
    is_delta = bounce.is_microfacet_based && (ctx.params.roughness <= GEOMETRY_DELTA_PATH_CUTOFF);
  }
  /* bsdf_is_pass_through_ray, bsdf_utils.cuh:69-73 */
  const bool pass_through = bounce.is_transparent_pass && ((ctx.params.ior == 1.0f) || !bounce.is_microfacet_based);

  if (c_any(ctx.params.emission))
    out->emission = c_mul(ctx.params.emission, rec_in);

  OrcRGB rec = c_mul(rec_in, bounce.weight);

  uint16_t new_state = state | ORC_STATE_USE_IGNORE_HANDLE;
  if (sky_on && !pass_through)
    new_state &= ~ORC_STATE_ALLOW_AMBIENT;
  else
    new_state |= ORC_STATE_ALLOW_AMBIENT;
  if (!is_delta)
    new_state &= ~ORC_STATE_DELTA_PATH;
  if (!pass_through) {
    new_state &= ~ORC_STATE_CAMERA_DIRECTION;
    new_state &= ~ORC_STATE_ALLOW_EMISSION;
  }

  out->bounce_alive = 1;
  /* task_russian_roulette, directives.cuh:11-32: tested on the state BEFORE the update */
  if (!(state & ORC_STATE_DELTA_PATH)) {
    const float value = c_importance(rec);
    if (value < cam->russian_roulette_threshold) {
      const float p = (value > 0.0f) ? fmaxf(value / cam->russian_roulette_threshold, RUSSIAN_ROULETTE_CLAMP) : 0.0f;
      if (orc_random_1d(ORC_RT_RUSSIAN_ROULETTE, pid, depth) > p)
        out->bounce_alive = 0;
      else
        rec = c_scale(rec, 1.0f / p);
    }
  }

  if (bounce.is_transparent_pass) { /* medium transition, geometry.cuh:160-175 */
    const bool inside = (ctx.params.flags & MF_REFRACTION_IS_INSIDE) != 0;
    if (!inside) {
      const float ray_ior = orc_ior_decompress(medium_ior & 0xFF);
      const float new_ior = ray_ior / ctx.params.ior;
      medium_ior          = (medium_ior << 8) | orc_ior_compress(new_ior);
    }
    else {
      medium_ior >>= 8;
    }
  }

  out->bounce_origin     = ctx.position;
  out->bounce_ray        = bounce.ray;
  out->bounce_record     = orc_record_pack(rec);
  out->bounce_state      = new_state;
  out->bounce_medium_ior = medium_ior;
  out->bounce_weight     = bounce.weight;
  out->normal            = ctx.normal;
  out->is_transparent_pass = bounce.is_transparent_pass;
}

/* sky_process_inscattering_events (kernels.cuh:356-389): runs between the trace and geometry_process_tasks when the scene has aerial
 * perspective (device_manager.c:475): adds in-scattering x throughput to the result and attenuates the throughput. Returns the
 * contribution; *record is updated in place (packed like the reference stores it). */
static bool aerial_on(const OrcScene* s, const OrcSettings* set) { return s->sky && set->sky_mode != 2 && s->sky->p.aerial_perspective; }
static OrcRGB vertex_inscatter(const OrcScene* s, OrcPathID pid, uint32_t depth, OrcVec3 origin, OrcVec3 ray, float t, OrcUint2* record) {
  OrcRGB tr;
  const OrcRGB ins = orc_sky_inscattering(s->sky, origin, ray, t, depth, orc_random_1d(ORC_RT_SKY_INSCATTERING_STEP, pid, depth),
                                          orc_random_1d(ORC_RT_SKY_STEP_OFFSET, pid, depth), &tr);
  const OrcRGB rec = orc_record_unpack(*record);
  *record          = orc_record_pack(c_mul(rec, tr));
  const OrcRGB add = c_mul(ins, rec);
  return c_any(add) ? add : c_splat(0.0f); /* write_beauty_buffer: color_any */
}

/* vertices come as the trace leaves them: with aerial perspective the in-scattering step is applied first, its contribution is
 * reported in `emission` (what the vertex adds to the result record before any shadow ray) */
void orc_shade_vertices(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t n, uint32_t depth, const OrcVertexIn* in,
                        OrcVertexOut* out, int num_threads) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int64_t i = 0; i < (int64_t) n; i++) {
    const OrcVertexIn* v = in + i;
    OrcUint2 record      = v->record;
    OrcRGB ins           = c_splat(0.0f);
    if (aerial_on(s, set))
      ins = vertex_inscatter(s, v->path_id, depth, v->origin, v->ray, v->t, &record);
    shade_vertex(s, cam, set, v->path_id, depth, (uint16_t) v->state, v->origin, v->ray, v->prim, v->t, record, v->medium_ior, out + i);
    out[i].emission = c_add(out[i].emission, ins);
  }
}

static OrcRGB trace_path(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t x, uint32_t y, uint32_t sample_id,
                         OrcRayCounts* counts, uint32_t capture_iter, OrcVertexIn* capture) {
  const OrcPathID pid = orc_path_id_get(x, y, sample_id);
  OrcVec3 origin, ray;
  orc_camera_sample(cam, set, pid, &origin, &ray);

  uint16_t state      = ORC_STATE_DELTA_PATH | ORC_STATE_CAMERA_DIRECTION | ORC_STATE_ALLOW_EMISSION | ORC_STATE_ALLOW_AMBIENT;
  OrcUint2 record     = orc_record_pack(c_splat(1.0f));
  uint32_t medium_ior = 0; /* medium_stack_ior_modify({}, 1.0f, true) */
  uint32_t ignore     = 0xFFFFFFFFu;
  OrcRGB result       = c_splat(0.0f);
  const bool sky_on   = set->sky_mode != 0; /* ambient NEE: direct_lighting_ambient_is_allowed */
  const OrcRGB sky    = (set->sky_mode == 2) ? set->sky_constant_color : c_splat(0.0f);

  for (uint32_t iter = 0; iter <= set->max_ray_depth; iter++) {
    /* device.state.depth as the kernels see it: UPDATE_DEPTH is skipped when depth + 1 == max_depth
     * (device_renderer.c:126-130), so the last iteration re-uses the previous value. */
    uint32_t depth = iter;
    if (iter == set->max_ray_depth && iter > 0)
      depth = iter - 1;

    counts->closest_rays++;
    const OrcHit hit = orc_closest_hit(s, origin, ray, 0.0f, ORC_FLT_MAX, (state & ORC_STATE_USE_IGNORE_HANDLE) ? ignore : 0xFFFFFFFFu, NULL, NULL);

    if (hit.prim == ORC_HIT_SKY) { /* sky_process_tasks, sky.cuh:609-633 */
      if (state & ORC_STATE_ALLOW_AMBIENT) {
        OrcRGB sky_color = sky;
        if (set->sky_mode != 2 && s->sky) {
          const bool include_sun = (state & (ORC_STATE_CAMERA_DIRECTION | ORC_STATE_ALLOW_EMISSION)) != 0;
          sky_color = orc_sky_color_mode(s->sky, set->sky_mode, origin, ray, include_sun, orc_random_1d(ORC_RT_SKY_STEP_OFFSET, pid, depth));
        }
        result = c_add(result, c_mul(sky_color, orc_record_unpack(record)));
      }
      break;
    }

    if (capture && iter == capture_iter) {
      capture->path_id    = pid;
      capture->state      = state;
      capture->origin     = origin;
      capture->ray        = ray;
      capture->prim       = hit.prim;
      capture->t          = hit.t;
      capture->record     = record;
      capture->medium_ior = medium_ior;
      return result;
    }

    if (aerial_on(s, set))
      result = c_add(result, vertex_inscatter(s, pid, depth, origin, ray, hit.t, &record));

    /* geometry_process_tasks */
    OrcVertexOut vo;
    shade_vertex(s, cam, set, pid, depth, state, origin, ray, hit.prim, hit.t, record, medium_ior, &vo);
    const OrcVec3 hit_point = vo.hit_point;
    const OrcRGB rec_in     = orc_record_unpack(record);
    OrcRGB nee              = c_splat(0.0f);

    if (s->has_lights) {
      /* direct_lighting_geometry_evaluate_task, direct_lighting.cuh:445-463 */
      if (vo.geo_light_id != ORC_LIGHT_ID_INVALID) {
        const uint32_t tprim =
          s->instance_prim_offset[s->light_tree.tri_handle_map[2 * vo.geo_light_id]] + s->light_tree.tri_handle_map[2 * vo.geo_light_id + 1];
        const OrcRGB vis = shadow_visibility(s, hit_point, vo.geo_ray, ORC_EPS, vo.geo_dist, hit.prim, tprim, &counts->shadow_rays);
        nee              = c_add(nee, c_mul(vo.geo_color, vis));
      }

      /* direct_lighting_bsdf_evaluate_task, direct_lighting.cuh:601-669 */
      if (vo.bsdf_prob != 0.0f) {
        counts->light_enum_rays++;
        uint32_t num_hits    = 0;
        const float trnd     = orc_random_1d(ORC_RT_LIGHT_BSDF_TRACE, pid, depth);
        const uint32_t light = enumerate_lights(s, hit_point, vo.bsdf_ray, hit.prim, trnd, &num_hits);
        if (light != ORC_LIGHT_ID_INVALID) {
          TriLight L       = light_init(s, light);
          const float dist = light_intersect(s, &L, hit_point, vo.bsdf_ray);
          if (dist != ORC_FLT_MAX) {
            OrcRGB lcol = light_color_of(s, &L);
            float mis   = 1.0f; /* mis_compute_weight_gi, mis.cuh:26-39 */
            if (vo.bsdf_root_sum != 0.0f) {
              const float power = c_importance(lcol) * light_area(&L);
              mis               = mis_weight_base(vo.bsdf_prob, light_solid_angle(&L, hit_point), power, dist * dist, vo.bsdf_root_sum);
            }
            lcol = c_scale(lcol, mis * num_hits);
            lcol = c_mul(lcol, vo.bsdf_weight);
            const uint32_t tprim = s->instance_prim_offset[s->light_tree.tri_handle_map[2 * light]] + s->light_tree.tri_handle_map[2 * light + 1];
            const OrcRGB vis     = shadow_visibility(s, hit_point, vo.bsdf_ray, ORC_EPS, dist, hit.prim, tprim, &counts->shadow_rays);
            nee                  = c_add(nee, c_mul(lcol, vis));
          }
        }
      }
    }

    DBG("iter %u depth %u prim %u t %f bounce ray (%f %f %f) w (%f %f %f) nee (%f %f %f) root_sum %f\n", iter, depth, hit.prim, hit.t, vo.bounce_ray.x,
        vo.bounce_ray.y, vo.bounce_ray.z, vo.bounce_weight.r, vo.bounce_weight.g, vo.bounce_weight.b, nee.r, nee.g, nee.b, vo.bsdf_root_sum);

    /* ambient NEE evaluation, direct_lighting.cuh:531-599 */
    if (sky_on && (vo.amb_color.x != 0 || vo.amb_color.y != 0)) {
      const OrcVec3 aray = orc_ray_unpack(vo.amb_ray);
      const OrcRGB vis   = shadow_visibility(s, hit_point, aray, ORC_EPS, ORC_FLT_MAX, hit.prim, 0xFFFFFFFFu, &counts->shadow_rays);
      nee                = c_add(nee, c_mul(orc_record_unpack(vo.amb_color), vis));
    }

    /* sun NEE evaluation, direct_lighting.cuh:465-529 (no ocean: one unbounded shadow ray, optix_anyhit.cuh:100-139) */
    if (vo.sun_color.x != 0 || vo.sun_color.y != 0) {
      const OrcVec3 sray = orc_ray_unpack(vo.sun_ray);
      const OrcRGB vis   = shadow_visibility(s, hit_point, sray, ORC_EPS, ORC_FLT_MAX, hit.prim, 0xFFFFFFFFu, &counts->shadow_rays);
      nee                = c_add(nee, c_mul(orc_record_unpack(vo.sun_color), vis));
    }

    /* emission + NEE into the result record */
    if (c_any(vo.emission))
      result = c_add(result, vo.emission);
    {
      const OrcRGB acc = c_mul(nee, rec_in);
      if (c_any(acc))
        result = c_add(result, acc);
    }

    if (!vo.bounce_alive)
      break;

    origin     = vo.bounce_origin;
    ray        = vo.bounce_ray;
    record     = vo.bounce_record;
    medium_ior = vo.bounce_medium_ior;
    ignore     = hit.prim;
    state      = (uint16_t) vo.bounce_state;
  }
  return result;
}

/* Per-vertex view of the NEE evaluation that follows geometry_process_tasks: the up to three shadow segments of one path vertex
 * (direct_lighting_geometry / bsdf / ambient _evaluate_task, direct_lighting.cuh:445-669) with their unshadowed contribution
 * ALREADY multiplied by the path throughput, the emitter they aim at and the transmittance the any-hit programs
 * (optix_anyhit.cuh:49-139) leave along them. Test infrastructure for the product's k_shade / k_trace_shadow pair. */
void orc_nee_segments(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t n, uint32_t depth, const OrcVertexIn* in,
                      OrcNeeSegment* out, int num_threads) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int64_t i = 0; i < (int64_t) n; i++) {
    const OrcVertexIn* v = in + i;
    OrcNeeSegment* seg   = out + ORC_NEE_SLOTS * i;
    memset(seg, 0, ORC_NEE_SLOTS * sizeof(OrcNeeSegment));
    for (int k = 0; k < ORC_NEE_SLOTS; k++)
      seg[k].target_prim = 0xFFFFFFFFu;
    OrcVertexOut vo;
    OrcUint2 record = v->record;
    if (aerial_on(s, set))
      vertex_inscatter(s, v->path_id, depth, v->origin, v->ray, v->t, &record);
    shade_vertex(s, cam, set, v->path_id, depth, (uint16_t) v->state, v->origin, v->ray, v->prim, v->t, record, v->medium_ior, &vo);
    const OrcVec3 hit_point = vo.hit_point;
    const OrcRGB rec_in     = orc_record_unpack(record);
    const bool sky_on       = set->sky_mode != 0;
    if (s->has_lights) {
      if (vo.geo_light_id != ORC_LIGHT_ID_INVALID) {
        const uint32_t tprim =
          s->instance_prim_offset[s->light_tree.tri_handle_map[2 * vo.geo_light_id]] + s->light_tree.tri_handle_map[2 * vo.geo_light_id + 1];
        const OrcRGB col = c_mul(vo.geo_color, rec_in);
        if (c_any(col)) {
          seg[0].valid = 1, seg[0].ray = vo.geo_ray, seg[0].dist = vo.geo_dist, seg[0].color = col, seg[0].target_prim = tprim;
          seg[0].visibility = shadow_visibility(s, hit_point, vo.geo_ray, ORC_EPS, vo.geo_dist, v->prim, tprim, NULL);
        }
      }
      if (vo.bsdf_prob != 0.0f) {
        uint32_t num_hits    = 0;
        const float trnd     = orc_random_1d(ORC_RT_LIGHT_BSDF_TRACE, v->path_id, depth);
        const uint32_t light = enumerate_lights(s, hit_point, vo.bsdf_ray, v->prim, trnd, &num_hits);
        seg[1].enum_hits     = num_hits;
        if (light != ORC_LIGHT_ID_INVALID) {
          TriLight L       = light_init(s, light);
          const float dist = light_intersect(s, &L, hit_point, vo.bsdf_ray);
          if (dist != ORC_FLT_MAX) {
            OrcRGB lcol = light_color_of(s, &L);
            float mis   = 1.0f;
            if (vo.bsdf_root_sum != 0.0f) {
              const float power = c_importance(lcol) * light_area(&L);
              mis               = mis_weight_base(vo.bsdf_prob, light_solid_angle(&L, hit_point), power, dist * dist, vo.bsdf_root_sum);
            }
            lcol = c_scale(lcol, mis * num_hits);
            lcol = c_mul(c_mul(lcol, vo.bsdf_weight), rec_in);
            const uint32_t tprim = s->instance_prim_offset[s->light_tree.tri_handle_map[2 * light]] + s->light_tree.tri_handle_map[2 * light + 1];
            if (c_any(lcol)) {
              seg[1].valid = 1, seg[1].ray = vo.bsdf_ray, seg[1].dist = dist, seg[1].color = lcol, seg[1].target_prim = tprim;
              seg[1].visibility = shadow_visibility(s, hit_point, vo.bsdf_ray, ORC_EPS, dist, v->prim, tprim, NULL);
            }
          }
        }
      }
    }
    if (sky_on && (vo.amb_color.x != 0 || vo.amb_color.y != 0)) {
      const OrcVec3 aray = orc_ray_unpack(vo.amb_ray);
      const OrcRGB col   = c_mul(orc_record_unpack(vo.amb_color), rec_in);
      if (c_any(col)) {
        seg[2].valid = 1, seg[2].ray = aray, seg[2].dist = ORC_FLT_MAX, seg[2].color = col;
        seg[2].visibility = shadow_visibility(s, hit_point, aray, ORC_EPS, ORC_FLT_MAX, v->prim, 0xFFFFFFFFu, NULL);
      }
    }
    if (vo.sun_color.x != 0 || vo.sun_color.y != 0) {
      const OrcVec3 sray = orc_ray_unpack(vo.sun_ray);
      const OrcRGB col   = c_mul(orc_record_unpack(vo.sun_color), rec_in);
      if (c_any(col)) {
        seg[3].valid = 1, seg[3].ray = sray, seg[3].dist = ORC_FLT_MAX, seg[3].color = col;
        seg[3].visibility = shadow_visibility(s, hit_point, sray, ORC_EPS, ORC_FLT_MAX, v->prim, 0xFFFFFFFFu, NULL);
      }
    }
  }
}

/* Transmittance of explicit shadow rays (shadow_trace any-hit, optix_anyhit.cuh:49-139): tmin = eps, hits at t < limit count, the
 * ignore and target primitives are skipped, an opaque hit gives 0, transparent hits multiply. visibility = 3 floats per ray. */
void orc_shadow_rays(const OrcScene* s, uint32_t n, const float* origins, const float* dirs, const float* limits, const uint32_t* ignore_prims,
                     const uint32_t* target_prims, float* visibility, int num_threads) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int64_t i = 0; i < (int64_t) n; i++) {
    const OrcRGB v = shadow_visibility(s, v_get(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]), v_get(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]),
                                       ORC_EPS, limits[i], ignore_prims[i], target_prims[i], NULL);
    visibility[3 * i + 0] = v.r, visibility[3 * i + 1] = v.g, visibility[3 * i + 2] = v.b;
  }
}

double orc_render_region(
  const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t first_sample, uint32_t num_samples, uint32_t x0, uint32_t y0,
  uint32_t x1, uint32_t y1, float* planes, int num_threads, OrcRayCounts* counts) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#endif
  const size_t npix = (size_t) set->width * set->height;
  uint64_t cr = 0, sr = 0, lr = 0;
  const uint32_t rw = x1 - x0, rh = y1 - y0;
  const double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : cr, sr, lr)
  for (int64_t i = 0; i < (int64_t) rw * rh; i++) {
    const uint32_t y = y0 + (uint32_t) (i / rw);
    const uint32_t x = x0 + (uint32_t) (i % rw);
    OrcRayCounts c   = {0, 0, 0};
    const size_t idx = x + (size_t) y * set->width;
    for (uint32_t k = 0; k < num_samples; k++) {
      const OrcRGB v = trace_path(s, cam, set, x, y, first_sample + k, &c, 0, NULL);
      /* accumulation_collect_results, accumulation.cuh:36-60 */
      planes[0 * npix + idx] += v.r;
      planes[1 * npix + idx] += v.g;
      planes[2 * npix + idx] += v.b;
      planes[3 * npix + idx] += c_luminance(c_mul(v, v));
    }
    cr += c.closest_rays;
    sr += c.shadow_rays;
    lr += c.light_enum_rays;
  }
  if (counts) {
    counts->closest_rays += cr;
    counts->shadow_rays += sr;
    counts->light_enum_rays += lr;
  }
  return now_s() - t0;
}

/* For every pixel: follows the path of sample `sample_id` to wavefront iteration `iter` and, if it is still alive there and hit
 * geometry, stores what geometry_process_tasks would load for it (valid[i] = 1). Inputs for orc_shade_vertices / the reference kernel. */
void orc_path_vertices(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t sample_id, uint32_t iter, OrcVertexIn* out,
                       uint8_t* valid, int num_threads) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int64_t i = 0; i < (int64_t) set->width * set->height; i++) {
    const uint32_t y = (uint32_t) (i / set->width), x = (uint32_t) (i % set->width);
    OrcRayCounts c   = {0, 0, 0};
    OrcVertexIn v;
    memset(&v, 0, sizeof(v));
    v.prim = ORC_HIT_INVALID;
    trace_path(s, cam, set, x, y, sample_id, &c, iter, &v);
    valid[i] = v.prim != ORC_HIT_INVALID;
    out[i]   = v;
  }
}

/* ------------------------------------------------------------------ */
/* debug shading modes: the one-bounce queue of _device_renderer_build_debug_kernel_queue (device/device_renderer.c:136-182):         */
/* raytrace, [inscattering], sort, geometry_process_tasks_debug (cuda/geometry.cuh:182-246), sky_process_tasks_debug (cuda/sky.cuh:635-668) */
/* ------------------------------------------------------------------ */
static OrcRGB debug_path(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t mode, uint32_t x, uint32_t y, uint32_t sample_id) {
  const OrcPathID pid = orc_path_id_get(x, y, sample_id);
  OrcVec3 origin, ray;
  orc_camera_sample(cam, set, pid, &origin, &ray);
  const uint16_t state = ORC_STATE_DELTA_PATH | ORC_STATE_CAMERA_DIRECTION | ORC_STATE_ALLOW_EMISSION | ORC_STATE_ALLOW_AMBIENT;
  OrcUint2 record      = orc_record_pack(c_splat(1.0f));
  OrcRGB result        = c_splat(0.0f);
  const OrcHit hit     = orc_closest_hit(s, origin, ray, 0.0f, ORC_FLT_MAX, 0xFFFFFFFFu, NULL, NULL);

  if (hit.prim == ORC_HIT_SKY) { /* sky_process_tasks_debug */
    if (mode == ORC_SHADING_MODE_ALBEDO) {
      /* sky_color_main(task.origin, task.ray, STATE_FLAG_CAMERA_DIRECTION, task.path_id) */
      OrcRGB sky_color = (set->sky_mode == 2) ? set->sky_constant_color : c_splat(0.0f);
      if (set->sky_mode != 2 && s->sky)
        sky_color = orc_sky_color_mode(s->sky, set->sky_mode, origin, ray, true, orc_random_1d(ORC_RT_SKY_STEP_OFFSET, pid, 0));
      return sky_color;
    }
    if (mode == ORC_SHADING_MODE_IDENTIFICATION)
      return c_get(0.0f, 0.63f, 1.0f);
    return result;
  }

  if (aerial_on(s, set)) /* sky_process_inscattering_events stays in the debug queue (device_renderer.c:151-155) */
    result = c_add(result, vertex_inscatter(s, pid, 0, origin, ray, hit.t, &record));

  const OrcVec3 hit_point = v_add(origin, v_scale(ray, hit.t)); /* task.origin + task.ray * trace.depth */
  OrcRGB dbg              = c_splat(0.0f);
  switch (mode) {
    case ORC_SHADING_MODE_ALBEDO: {
      const Ctx ctx = get_context(s, hit.prim, hit_point, ray, state, 0);
      dbg           = c_add(ctx.params.albedo, ctx.params.emission);
    } break;
    case ORC_SHADING_MODE_DEPTH: {
      const float v = orc_saturate((1.0f / hit.t) * 2.0f);
      dbg           = c_splat(v);
    } break;
    case ORC_SHADING_MODE_NORMAL: {
      const Ctx ctx = get_context(s, hit.prim, hit_point, ray, state, 0);
      dbg           = c_get(orc_saturate(ctx.normal.x), orc_saturate(ctx.normal.y), orc_saturate(ctx.normal.z));
    } break;
    case ORC_SHADING_MODE_IDENTIFICATION: {
      const uint32_t v = orc_squares32(0x55555555u, (s->prim_instance[hit.prim] << 16) | s->prim_tri[hit.prim]);
      dbg = c_get((float) (uint16_t) (v & 0x7ff) / 0x7ff, (float) (uint16_t) ((v >> 10) & 0x7ff) / 0x7ff, (float) (uint16_t) ((v >> 20) & 0x7ff) / 0x7ff);
    } break;
    case ORC_SHADING_MODE_LIGHTS: {
      const Ctx ctx = get_context(s, hit.prim, hit_point, ray, state, 0);
      dbg           = c_add(c_scale(ctx.params.albedo, 0.025f), ctx.params.emission);
    } break;
    default:
      break;
  }
  return c_add(result, dbg); /* write_beauty_buffer adds to the task's result */
}

double orc_render_debug(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t shading_mode, uint32_t first_sample,
                        uint32_t num_samples, float* planes, int num_threads) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#endif
  const size_t npix = (size_t) set->width * set->height;
  const double t0   = now_s();
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < (int64_t) npix; i++) {
    const uint32_t y = (uint32_t) (i / set->width), x = (uint32_t) (i % set->width);
    for (uint32_t k = 0; k < num_samples; k++) {
      const OrcRGB v = debug_path(s, cam, set, shading_mode, x, y, first_sample + k);
      planes[0 * npix + i] += v.r;
      planes[1 * npix + i] += v.g;
      planes[2 * npix + i] += v.b;
      planes[3 * npix + i] += c_luminance(c_mul(v, v));
    }
  }
  return now_s() - t0;
}

double orc_render(
  const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t first_sample, uint32_t num_samples, float* planes, int num_threads,
  OrcRayCounts* counts) {
  return orc_render_region(s, cam, set, first_sample, num_samples, 0, 0, set->width, set->height, planes, num_threads, counts);
}

size_t orc_sizeof_vertex_in(void) { return sizeof(OrcVertexIn); }
size_t orc_sizeof_vertex_out(void) { return sizeof(OrcVertexOut); }

/* ------------------------------------------------------------------ */
/* adaptive sampling: the render loop (orc_adaptive.c holds the estimators)                                              */
/* tasks_create_adaptive_sampling (cuda/kernels.cuh:195-356): in stage s >= 1 every pixel of a block gets count_{s-1}(block)     */
/* samples per execution, with sample ids offset(block) + k; offset = samples the pixel has already received.            */
/* ------------------------------------------------------------------ */
uint64_t orc_render_adaptive(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, const OrcAdaptiveParams* p, uint32_t num_executions,
                             float* planes, uint32_t* words, uint32_t* executions, uint32_t* stage, int num_threads, OrcRayCounts* counts) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#endif
  const uint32_t W = set->width, H = set->height;
  const uint32_t bw = (W + 3) >> 2, bh = (H + 3) >> 2;
  const size_t npix = (size_t) W * H;
  uint64_t paths    = 0;
  float* block_var  = (float*) malloc(sizeof(float) * bw * bh);
  for (uint32_t e = 0; e < num_executions; e++) {
    const uint32_t st = *stage;
    uint64_t cr = 0, sr = 0, lr = 0, np = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : cr, sr, lr, np)
    for (int64_t i = 0; i < (int64_t) npix; i++) {
      const uint32_t y = (uint32_t) (i / W), x = (uint32_t) (i % W);
      const uint32_t word = words[(x >> 2) + (y >> 2) * bw];
      const uint32_t n    = (st == 0) ? 1u : orc_adaptive_stage_count(word, st - 1);
      const uint32_t base = orc_adaptive_block_samples(word, executions); /* adapative_sampling_get_sample_offset */
      OrcRayCounts c      = {0, 0, 0};
      for (uint32_t k = 0; k < n; k++) {
        const uint32_t sample_id = base + k;
        if (sample_id >= (1u << 20))
          continue;
        const OrcRGB v = trace_path(s, cam, set, x, y, sample_id, &c, 0, NULL);
        planes[0 * npix + i] += v.r;
        planes[1 * npix + i] += v.g;
        planes[2 * npix + i] += v.b;
        planes[3 * npix + i] += c_luminance(c_mul(v, v));
        np++;
      }
      cr += c.closest_rays, sr += c.shadow_rays, lr += c.light_enum_rays;
    }
    paths += np;
    if (counts)
      counts->closest_rays += cr, counts->shadow_rays += sr, counts->light_enum_rays += lr;
    executions[st]++;
    /* _device_renderer_queue_adaptive_sampling_update, device_renderer.c:350-376 */
    if (st < ORC_ADAPTIVE_STAGES && executions[st] >= (p->update_interval << st)) {
      const float sum = orc_adaptive_block_variance(planes, W, H, words, executions, p, block_var);
      orc_adaptive_stage_counts(block_var, sum, bw * bh, st, p, words);
      *stage = st + 1;
    }
  }
  free(block_var);
  return paths;
}

void orc_adaptive_resolve(const float* planes, uint32_t width, uint32_t height, const uint32_t* words, const uint32_t* executions, float* rgb) {
  const uint32_t bw = (width + 3) >> 2;
  const size_t n    = (size_t) width * height;
  for (uint32_t y = 0; y < height; y++)
    for (uint32_t x = 0; x < width; x++) {
      const size_t i  = x + (size_t) y * width;
      const float inv = 1.0f / (float) orc_adaptive_block_samples(words[(x >> 2) + (y >> 2) * bw], executions);
      rgb[i] = planes[i] * inv, rgb[n + i] = planes[n + i] * inv, rgb[2 * n + i] = planes[2 * n + i] * inv;
    }
}
