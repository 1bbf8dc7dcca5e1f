// shade_api.cuh - host-visible interface of the shading translation unit (shade.cu).
#pragma once

#include "lumb200_internal.cuh"
#include "sky.cuh"
#include "texture.cuh"
#include "wavefront.cuh"

#define LB_RNG_TARGET_COUNT 577  // == lbrng::T_COUNT
#define LB_RNG_TABLE_DEPTHS 64    // max_ray_depth < 64

#define LB_LUT_SIZE 32  // BSDF_LUT_SIZE, reference device_utils.h:42

struct LbLutTexObjects {
  cudaTextureObject_t conductor;       // 32 x 32, R16 unorm, linear, clamp (device_bsdf.c:7-54)
  cudaTextureObject_t glossy;          // 32 x 32
  cudaTextureObject_t dielectric;      // 32 x 32 x 32
  cudaTextureObject_t dielectric_inv;  // 32 x 32 x 32
};

struct LbLutTextures {
  bool valid = false;
  LbLutTexObjects tex = {0, 0, 0, 0};
  cudaArray_t arrays[4] = {nullptr, nullptr, nullptr, nullptr};
  uint16_t* d_data[4]   = {nullptr, nullptr, nullptr, nullptr};
};

struct LbShadeParams {
  LbPaths paths;
  LbFrame frame;
  LbCameraDev camera;
  const uint32_t* bluenoise;
  const uint4* rng_table;  // [depth][target], see lbrng::TabSampler
  uint32_t sample_id;
  uint32_t rng_depth;
  uint32_t is_last;
  LbCounters* counters;
  const uint32_t* queue_in;  // sorted: [0, n_hits) surface hits, [n_hits, n_active) misses
  uint32_t* queue_out;       // survivors of this bounce
  // scene
  const uint2* prim_handle;
  const float4* const* mesh_vertices;
  const uint4* const* mesh_textris;
  const uint32_t* instance_mesh;
  const LbTransform* instance_xform;
  const uint32_t* instance_offset;
  const uint4* materials;
  const uint16_t* prim_material;
  const LbTexture* textures;  // material textures (texture.cuh); textured = some material references one
  uint32_t num_textures;
  uint32_t textured;
  uint32_t adaptive;  // paths carry their own sample ids (P.paths.sample_id): k_shade<*, *, true>
  uint32_t count;     // instrumented pass (lumb200_device_measure_traversal): k_shade<*, *, false, true>
  uint32_t class_materials[LB_NUM_CLASSES];  // materials per class (host side: classes without materials are not launched)
  LbLutTexObjects luts;
  LbSkyDev sky;  // procedural atmosphere (frame.sky_mode 0 / 1): miss shading and the sun's NEE
  // lights
  const uint4* light_root;
  const float4* light_root_children;  // decoded root children, 2 x float4 each (k_unpack_light_root)
  const uint4* light_nodes;
  const uint2* light_handles;
  const uint32_t* light_prims;
  const float4* light_records;  // 4 x float4 per light, see k_build_light_records
  uint32_t num_lights;
  Bvh8 light_bvh;
};

// returns the number of kernels launched. aux != nullptr: the metal class and the misses run on `aux` beside the other classes (fork / join
// events around them), so that the small launches fill the tail of the large one instead of following it
int lb_launch_shade(const LbShadeParams& sp, int grid, cudaStream_t s, cudaStream_t aux = nullptr, cudaEvent_t fork = nullptr,
                    cudaEvent_t join = nullptr);
int lb_launch_shade_debug(const LbShadeParams& sp, uint32_t mode, int grid, cudaStream_t s);  // debug shading modes 1..5, same return
// sky.cu
void lb_launch_sky_transmittance_lut(const LbSkyDev& sky, float4* dst_low, float4* dst_high, cudaStream_t s);
void lb_launch_sky_multiscattering_lut(const LbSkyDev& sky, float4* dst_low, float4* dst_high, cudaStream_t s);
void lb_launch_shade_miss_sky(const LbShadeParams& sp, int grid, cudaStream_t s);
void lb_launch_sky_inscattering(const LbShadeParams& sp, int grid, cudaStream_t s);  // queue_in = the unsorted queue of the bounce
void lb_launch_sky_hdri(const LbSkyDev& sky, const uint32_t* bluenoise, const float origin[3], uint32_t dim, uint32_t sample_count, float4* dst,
                        cudaStream_t s);
void lb_launch_enum_finish(const LbShadeParams& sp, int grid, cudaStream_t s);
void lb_launch_sample_texture(const LbTexture* textures, uint32_t num_textures, uint32_t tex, const float2* uv, uint32_t n, float lod, float4* out,
                              cudaStream_t s);
void lb_launch_mipmap_level(cudaTextureObject_t src, cudaSurfaceObject_t dst, uint32_t width, uint32_t height, uint32_t type, cudaStream_t s);
void lb_launch_light_compute_intensity(const LbShadeParams& sp, const uint32_t* mesh_ids, const uint32_t* tri_ids, uint32_t count, float* out,
                                       cudaStream_t s);
void lb_launch_build_light_records(const LbShadeParams& sp, float4* records, cudaStream_t s);
void lb_launch_unpack_light_root(const void* root, float4* out, uint32_t num_sections, cudaStream_t s);
void lb_launch_rng_table(uint4* table, uint32_t sample_id, uint32_t depths, cudaStream_t s);
void lb_launch_accumulate(const LbPaths& P, const LbFrame& F, float* planes, LbCounters* counters, int grid, cudaStream_t s);
void lb_launch_output_argb8(const float* planes, uint32_t width, uint32_t height, uint32_t sample_count, const Lumb200OutputParams& op,
                            const uint16_t* bluenoise_1d, void* dst, int grid, cudaStream_t s, bool raw = false);
uint32_t lb_bloom_mip_count(uint32_t width, uint32_t height);
void lb_launch_bloom(float* result, uint32_t width, uint32_t height, float* const* mips, uint32_t mip_count, float blend, int grid, cudaStream_t s);
void lb_launch_resolve(const float* planes, float* result, uint32_t width, uint32_t height, const LbAdaptive& A, uint32_t uniform_count,
                       uint32_t mode, uint32_t local_error_minimization, uint32_t stage, const Lumb200OutputParams& tm, int grid, cudaStream_t s);
void lb_launch_adaptive_build_stage(const float* planes, uint32_t width, uint32_t height, const LbAdaptive& A, const Lumb200OutputParams& tm,
                                    uint32_t stage, uint32_t max_rate, uint32_t avg_rate, uint32_t* words, float* block_variance, float* sum,
                                    uint32_t* task_prefix, uint32_t* total_tasks, cudaStream_t s);
void lb_launch_raygen_adaptive(const LbPaths& P, const LbFrame& F, const LbCameraDev& cam, const uint32_t* bluenoise, const LbAdaptive& A,
                               uint32_t stage, const uint32_t* task_prefix, uint32_t num_blocks, uint32_t task_begin, uint32_t n_tasks,
                               uint32_t* queue, LbCounters* C, int grid, cudaStream_t s);
void lb_launch_accumulate_adaptive(const LbPaths& P, uint32_t n_slots, uint32_t n_pixels, float* planes, LbCounters* counters, int grid,
                                   cudaStream_t s);
void lb_launch_generate_result(const float* planes, float* result, uint32_t num_pixels, uint32_t sample_count, int grid, cudaStream_t s);

Lumb200Result lb_lut_generate(LbLutTextures* luts, const uint32_t* bluenoise, cudaStream_t s);
Lumb200Result lb_lut_upload(LbLutTextures* luts, const uint16_t* conductor, const uint16_t* glossy, const uint16_t* dielectric,
                            const uint16_t* dielectric_inv, cudaStream_t s);
Lumb200Result lb_lut_download(LbLutTextures* luts, uint16_t* conductor, uint16_t* glossy, uint16_t* dielectric, uint16_t* dielectric_inv,
                              cudaStream_t s);
void lb_lut_destroy(LbLutTextures* luts);
