// texture.cuh - material textures on the device: the fetch, the any-hit alpha test and the shadow albedo.
//
// Replaces, on the reference's per-bounce path:
//   texture_load / texture_is_valid        cuda/texture_utils.cuh:24-45   (tex2DLod at mip level 0, v flip, gamma on rgb only)
//   load_texture_object                    cuda/memory.cuh:505-517        (16-byte DeviceTextureObject {handle, gamma})
//   load_triangle_tex_coords / lerp_uv     cuda/memory.cuh:414-425, cuda/math.cuh:246-253
//   optix_alpha_test                       cuda/optix_common.cuh:20-46    (closest-hit any-hit: alpha == 0 texels are cut out)
//   optix_get_albedo_for_shadowing         cuda/optix_common.cuh:48-66    (shadow / light-enumeration any-hit)
// Texture objects are created by lumb200_device_update_textures (device_api.cu) the way device_texture_create does
// (device/device_texture.c): normalised coordinates, per-axis address mode, point / linear filter, unorm reads of u8 / u16
// texels. Every load of the path uses mip level 0 (texture_get_default_args), so no mip chain is built.
#pragma once

#include "lumb200_internal.cuh"

#define LB_TEXTURE_NONE 0xFFFFu

struct LbTexture {  // DeviceTextureObject
  cudaTextureObject_t handle;  // 0 = invalid texture (TEXTURE_OBJECT_INVALID)
  float gamma;
  uint32_t size;  // width | height << 16 (DeviceTextureObject.width / .height)
};

// what the traversal kernels need to evaluate a textured any-hit; passed by value
struct LbTexScene {
  const LbTexture* textures;
  uint32_t num_textures;
  const uint4* materials;  // DeviceMaterialCompressed as 2 x uint4
  const uint2* prim_handle;
  const uint32_t* instance_mesh;
  const uint4* const* mesh_textris;
  const uint16_t* prim_material;
  const float4* shadow_tab;  // per material any-hit response; w == 2 marks the materials whose albedo texture has to be fetched per hit
};

__device__ __forceinline__ bool lb_texture_valid(const LbTexture* __restrict__ textures, uint32_t num_textures, uint32_t tex, LbTexture& out) {
  if (tex >= num_textures)
    return false;
  const uint4 raw = __ldg((const uint4*) (textures + tex));
  out.handle      = ((unsigned long long) raw.y << 32) | raw.x;
  out.gamma       = __uint_as_float(raw.z);
  out.size        = raw.w;
  return out.handle != 0ull;
}

__device__ __forceinline__ float2 lb_uv_unpack(uint32_t p) {  // uv_unpack, cuda/math.cuh:1706-1713
  return make_float2(__uint_as_float(p & 0xFFFF0000u), __uint_as_float(p << 16));
}

__device__ __forceinline__ float2 lb_lerp_uv(uint4 textri, float cu, float cv) {  // lerp_uv, cuda/math.cuh:246-253
  const float2 t0 = lb_uv_unpack(textri.x), t1 = lb_uv_unpack(textri.y), t2 = lb_uv_unpack(textri.z);
  return make_float2(t0.x + cu * (t1.x - t0.x) + cv * (t2.x - t0.x), t0.y + cu * (t1.y - t0.y) + cv * (t2.y - t0.y));
}

// texture_load with an already validated texture object
__device__ __forceinline__ float4 lb_texture_fetch(const LbTexture& t, float2 uv, bool flip_v, bool apply_gamma) {
  float4 r = tex2D<float4>(t.handle, uv.x, flip_v ? 1.0f - uv.y : uv.y);
  if (apply_gamma) {
    r.x = powf(r.x, t.gamma);
    r.y = powf(r.y, t.gamma);
    r.z = powf(r.z, t.gamma);
  }
  return r;
}

__device__ __forceinline__ float4 lb_texture_load(const LbTexture* __restrict__ textures, uint32_t num_textures, uint32_t tex, float2 uv, bool flip_v,
                                                   bool apply_gamma, float4 def) {
  LbTexture t;
  if (!lb_texture_valid(textures, num_textures, tex, t))
    return def;
  return lb_texture_fetch(t, uv, flip_v, apply_gamma);
}

__device__ __forceinline__ uint4 lb_prim_textri(const LbTexScene& T, uint32_t prim) {
  const uint2 handle = __ldg(T.prim_handle + prim);
  return __ldg(T.mesh_textris[__ldg(T.instance_mesh + handle.x)] + handle.y);
}

// optix_alpha_test: true when the hit lies on a texel with alpha == 0 and has to be ignored
__device__ __forceinline__ bool lb_alpha_cutout(const LbTexScene& T, uint32_t prim, float bu, float bv) {
  const uint32_t mid = __ldg(T.prim_material + prim);
  if (__ldg(&T.shadow_tab[mid].w) != 2.0f)  // untextured, or a texture whose alpha is 1 everywhere
    return false;
  const uint32_t tex = __ldg(&T.materials[2 * mid + 1].z) & 0xFFFFu;  // albedo_tex
  LbTexture t;
  if (!lb_texture_valid(T.textures, T.num_textures, tex, t))
    return false;
  const float2 uv = lb_lerp_uv(lb_prim_textri(T, prim), bu, bv);
  return tex2D<float4>(t.handle, uv.x, 1.0f - uv.y).w == 0.0f;
}

// optix_get_albedo_for_shadowing for a material that has an albedo texture
__device__ __forceinline__ float4 lb_shadow_albedo(const LbTexScene& T, uint32_t prim, uint32_t tex, float bu, float bv) {
  LbTexture t;
  if (!lb_texture_valid(T.textures, T.num_textures, tex, t))
    return make_float4(0.9f, 0.9f, 0.9f, 1.0f);
  return lb_texture_fetch(t, lb_lerp_uv(lb_prim_textri(T, prim), bu, bv), true, true);
}

// shadow any-hit response (cuda/optix_anyhit.cuh:49-93) from an albedo: (r, g, b multiplier, w = 1 if opaque)
__device__ __forceinline__ float4 lb_shadow_response(float4 albedo, bool colored) {
  if (albedo.w == 1.0f)
    return make_float4(0.0f, 0.0f, 0.0f, 1.0f);
  if (albedo.w == 0.0f && !colored)
    return make_float4(1.0f, 1.0f, 1.0f, 0.0f);
  const float tr = 1.0f - albedo.w;
  return colored ? make_float4(albedo.x * tr, albedo.y * tr, albedo.z * tr, 0.0f) : make_float4(tr, tr, tr, 0.0f);
}
