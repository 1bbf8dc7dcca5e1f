/* orc_texture.c - CPU restatement of the reference's material-texture fetch. TEST INFRASTRUCTURE ONLY.
 *
 * The reference samples material textures with tex2DLod<float4>(handle, u, v, 0) on CUDA texture objects created by
 * device_texture_create (device/device_texture.c:1-330): normalised coordinates, address mode per axis, point or
 * linear filter, unorm read mode for u8 / u16 data. texture_load (cuda/texture_utils.cuh:28-45) flips v, and applies
 * powf(rgb, gamma) but never to alpha. Every load on the per-bounce path uses mip level 0, so mip chains do not
 * enter. The filter arithmetic below is the one the CUDA programming guide publishes ("Texture Fetching": xB = x - 0.5,
 * i = floor(xB), weights in 1.8 fixed point); the unit's internal precision is not published, so this restatement is
 * pinned against the hardware by tests/test_texture_gpu.py with a stated tolerance. */
#include <math.h>
#include <stdlib.h>

#include "orc_internal.h"

static int address(int i, int n, uint32_t mode, bool* border) {
  *border = false;
  switch (mode) {
    case 0: { /* wrap */
      int m = i % n;
      return (m < 0) ? m + n : m;
    }
    case 2: { /* mirror */
      int m = i % (2 * n);
      if (m < 0)
        m += 2 * n;
      return (m >= n) ? 2 * n - 1 - m : m;
    }
    case 3: /* border */
      if (i < 0 || i >= n) {
        *border = true;
        return 0;
      }
      return i;
    default: /* clamp */
      return (i < 0) ? 0 : ((i > n - 1) ? n - 1 : i);
  }
}

static void texel(const OrcTexture* t, int x, int y, float out[4]) {
  bool bx, by;
  const int ix = address(x, (int) t->width, t->wrap_u, &bx);
  const int iy = address(y, (int) t->height, t->wrap_v, &by);
  out[0] = out[1] = out[2] = 0.0f;
  out[3]                   = 1.0f; /* missing components read (0, 0, 0, 1) */
  if (bx || by) {
    out[3] = (t->num_components == 4) ? 0.0f : 1.0f;
    return;
  }
  const uint8_t* row = (const uint8_t*) t->data + (size_t) t->pitch * iy;
  for (uint32_t c = 0; c < t->num_components && c < 4; c++) {
    const size_t k = (size_t) ix * t->num_components + c;
    switch (t->type) {
      case ORC_TEX_U8:
        out[c] = row[k] / 255.0f;
        break;
      case ORC_TEX_U16:
        out[c] = ((const uint16_t*) row)[k] / 65535.0f;
        break;
      default:
        out[c] = ((const float*) row)[k];
        break;
    }
  }
}

void orc_texture_fetch(const OrcTexture* t, float u, float v, float out[4]) {
  if (t->filter == 0) { /* point */
    texel(t, (int) floorf(u * t->width), (int) floorf(v * t->height), out);
    return;
  }
  const float xb = u * t->width - 0.5f, yb = v * t->height - 0.5f;
  const float fx = floorf(xb), fy = floorf(yb);
  const float a = floorf((xb - fx) * 256.0f + 0.5f) * (1.0f / 256.0f);
  const float b = floorf((yb - fy) * 256.0f + 0.5f) * (1.0f / 256.0f);
  float t00[4], t10[4], t01[4], t11[4];
  texel(t, (int) fx, (int) fy, t00);
  texel(t, (int) fx + 1, (int) fy, t10);
  texel(t, (int) fx, (int) fy + 1, t01);
  texel(t, (int) fx + 1, (int) fy + 1, t11);
  for (int c = 0; c < 4; c++)
    out[c] = (1.0f - a) * (1.0f - b) * t00[c] + a * (1.0f - b) * t10[c] + (1.0f - a) * b * t01[c] + a * b * t11[c];
}

void orc_scene_set_textures(OrcScene* s, const OrcTexture* textures, uint32_t count) {
  free(s->textures);
  s->textures     = NULL;
  s->num_textures = 0;
  if (!count)
    return;
  s->textures = (OrcTexture*) malloc(sizeof(OrcTexture) * count);
  memcpy(s->textures, textures, sizeof(OrcTexture) * count);
  s->num_textures = count;
}

bool orc_texture_valid(const OrcScene* s, uint16_t tex) { return tex < s->num_textures && s->textures[tex].data != NULL; }

/* texture_load, cuda/texture_utils.cuh:28-45 */
void orc_texture_load(const OrcScene* s, uint16_t tex, float u, float v, bool flip_v, bool apply_gamma, const float def[4], float out[4]) {
  if (!orc_texture_valid(s, tex)) {
    memcpy(out, def, sizeof(float) * 4);
    return;
  }
  const OrcTexture* t = &s->textures[tex];
  orc_texture_fetch(t, u, flip_v ? 1.0f - v : v, out);
  if (apply_gamma) {
    out[0] = powf(out[0], t->gamma);
    out[1] = powf(out[1], t->gamma);
    out[2] = powf(out[2], t->gamma);
  }
}

/* load_triangle_tex_coords, cuda/memory.cuh:414-425: the device keeps bfloat16-truncated uv (device_packing.c:36-43) */
OrcFloat2 orc_prim_tex_coords(const OrcScene* s, uint32_t prim, float cu, float cv) {
  const uint32_t inst = s->prim_instance[prim], tri = s->prim_tri[prim];
  const float* uv     = s->meshes[s->instances[inst].mesh_id].uv + 6 * (size_t) tri;
  const OrcFloat2 t0 = orc_unpack_uv(orc_pack_uv(uv[0], uv[1])), t1 = orc_unpack_uv(orc_pack_uv(uv[2], uv[3])),
                  t2 = orc_unpack_uv(orc_pack_uv(uv[4], uv[5]));
  OrcFloat2 r;
  r.x = t0.x + cu * (t1.x - t0.x) + cv * (t2.x - t0.x); /* lerp_uv, math.cuh:246-253 */
  r.y = t0.y + cu * (t1.y - t0.y) + cv * (t2.y - t0.y);
  return r;
}

static const OrcMaterialPacked* prim_material(const OrcScene* s, uint32_t prim) {
  const uint32_t inst = s->prim_instance[prim];
  return &s->materials[s->meshes[s->instances[inst].mesh_id].material[s->prim_tri[prim]]];
}

/* optix_alpha_test, cuda/optix_common.cuh:20-46: true when the hit is a cut-out (alpha == 0) and must be ignored */
bool orc_alpha_cutout(const OrcScene* s, uint32_t prim, float bu, float bv) {
  const OrcMaterialPacked* m = prim_material(s, prim);
  if (m->albedo_tex == 0xFFFF || !orc_texture_valid(s, m->albedo_tex))
    return false;
  const OrcFloat2 uv = orc_prim_tex_coords(s, prim, bu, bv);
  const float def[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  float c[4];
  orc_texture_load(s, m->albedo_tex, uv.x, uv.y, true, true, def, c);
  return c[3] == 0.0f;
}

/* optix_get_albedo_for_shadowing, cuda/optix_common.cuh:48-66 */
void orc_shadow_albedo(const OrcScene* s, uint32_t prim, float bu, float bv, float rgba[4]) {
  const OrcMaterialPacked* m = prim_material(s, prim);
  const float inv            = 1.0f / 0xFFFF;
  rgba[0] = m->albedo_r * inv, rgba[1] = m->albedo_g * inv, rgba[2] = m->albedo_b * inv, rgba[3] = m->albedo_a * inv;
  if (m->albedo_tex == 0xFFFF)
    return;
  if (!orc_texture_valid(s, m->albedo_tex)) {
    rgba[0] = rgba[1] = rgba[2] = 0.9f;
    rgba[3]                     = 1.0f;
    return;
  }
  const OrcFloat2 uv = orc_prim_tex_coords(s, prim, bu, bv);
  const float def[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  orc_texture_load(s, m->albedo_tex, uv.x, uv.y, true, true, def, rgba);
}
