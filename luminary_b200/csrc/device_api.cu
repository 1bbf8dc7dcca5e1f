// device_api.cu - implementation of the C ABI declared in include/lumb200.h.
//
// One Lumb200Device owns one CUDA device: scene tables, the two BVH8s (all geometry / emitters only), the
// wavefront state, the accumulation planes and a stream. It plays the role of the reference's `Device`
// object (device/device.c) for the path-tracing hot path; the per-bounce launch schedule of
// device/device_renderer.c:53-134 is re-stated in render_pass() below.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "lumb200_internal.cuh"
#include "shade_api.cuh"
#include "wavefront.cuh"

// launchers implemented in trace.cu
void lb_launch_raygen(const LbPaths& P, const LbFrame& F, const LbCameraDev& cam, const uint32_t* bluenoise, uint32_t sample_id, uint32_t* queue,
                      LbCounters* C, int grid, cudaStream_t s, bool tile_order = false);
void lb_launch_trace_closest(const Bvh8& bvh, const LbPaths& P, const uint32_t* queue, LbCounters* C, float2* uv, int grid, cudaStream_t s,
                             bool count, const LbTexScene* tex);
void lb_launch_trace_shadow(const Bvh8& bvh, const LbPaths& P, LbCounters* C, const uint16_t* prim_material,
                            const float4* shadow_tab, int grid, cudaStream_t s, bool count, const LbTexScene* tex);
void lb_launch_raygen_adaptive(const LbPaths& P, const LbFrame& F, const LbCameraDev& cam, const uint32_t* bluenoise, const LbAdaptive& A,
                               uint32_t stage, const uint32_t* task_prefix, uint32_t num_blocks, uint32_t task_begin, uint32_t n_tasks,
                               uint32_t* queue, LbCounters* C, int grid, cudaStream_t s);
void lb_launch_sort(const LbPaths& P, const uint32_t* queue_in, uint32_t* queue_out, LbCounters* C, const uint16_t* prim_material,
                    const uint16_t* sort_rank, const LbSortClasses& classes, uint32_t* bins, int grid, cudaStream_t s);
void lb_launch_trace_enum(const Bvh8& light_bvh, const LbPaths& P, LbCounters* C, const uint32_t* light_prims, const LbTexScene& T, bool textured,
                          int grid, cudaStream_t s);
void lb_launch_next_bounce(LbCounters* C, cudaStream_t s);
void lb_launch_load_rays(const LbPaths& P, const float* origins, const float* dirs, uint32_t n, uint32_t* queue, LbCounters* C, int grid,
                         cudaStream_t s);
void lb_launch_raygen_pixel(const LbPaths& P, const LbFrame& F, const LbCameraDev& cam, const uint32_t* bluenoise, uint32_t x, uint32_t y,
                            uint32_t sample_id, uint32_t* queue, LbCounters* C, cudaStream_t s);
void lb_launch_load_vertices(const LbPaths& P, const Lumb200VertexIn* in, uint32_t n, uint32_t width, uint32_t sample_id, uint32_t* queue,
                             LbCounters* C, int grid, cudaStream_t s);
void lb_launch_extract_segments(const LbPaths& P, const LbCounters* C, Lumb200VertexOut* out, int grid, cudaStream_t s);
void lb_launch_extract_vertices(const LbPaths& P, uint32_t n, const uint32_t* queue_out, const LbCounters* C, Lumb200VertexOut* out, int grid,
                                cudaStream_t s);
void lb_launch_load_shadow_rays(const LbPaths& P, const float* origins, const float* dirs, const float* max_dist, const uint32_t* ignore_prims,
                                const uint32_t* target_prims, uint32_t n, LbCounters* C, int grid, cudaStream_t s);
void lb_launch_extract_visibility(const LbPaths& P, uint32_t n, float* out, int grid, cudaStream_t s);
void lb_launch_extract_hits(const LbPaths& P, const uint2* prim_handle, const float2* uv, uint32_t n, uint32_t* inst, uint32_t* tri, float* t,
                            float* u, float* v, int grid, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// error string
// ---------------------------------------------------------------------------------------------
static thread_local char g_last_error[1024] = "";

extern "C" void lumb200_set_last_error(const char* fmt, ...) {
  va_list args;
  va_start(args, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, args);
  va_end(args);
}

extern "C" const char* lumb200_last_error(void) { return g_last_error; }

#define LB_REQUIRE(cond, code, ...)      \
  do {                                   \
    if (!(cond)) {                       \
      lumb200_set_last_error(__VA_ARGS__); \
      return (code);                     \
    }                                    \
  } while (0)

#define LB_TRAVERSAL_MAX_LEVELS 16  // == LB_LOOP_STACK / 2 (trace_loop.cuh), <= LB_STACK_SIZE (traverse.cuh)

#define LB_TRY(expr)                  \
  do {                                \
    Lumb200Result _r = (expr);        \
    if (_r != LUMB200_SUCCESS)        \
      return _r;                      \
  } while (0)

// ---------------------------------------------------------------------------------------------
// device object
// ---------------------------------------------------------------------------------------------
struct MeshDev {
  uint32_t num_tris = 0;
  float4* vertices  = nullptr;  // 3 per triangle: position + packed normal
  uint4* textris    = nullptr;  // packed uv x3 + material id
  std::vector<uint16_t> host_material;  // material id per triangle (host copy, used for the per-prim table)
  uint16_t max_material = 0;
};

struct TextureDev {  // DeviceTexture, device/device_texture.h
  cudaArray_t array       = nullptr;
  cudaMipmappedArray_t mips = nullptr;  // instead of `array` when the texture carries a generated mip chain
  uint32_t num_levels     = 1;
  cudaTextureObject_t obj = 0;
  float gamma             = 1.0f;
  uint32_t width = 0, height = 0;
  bool fully_opaque       = false;  // every fetch returns alpha == 1: four components, all texels at full alpha, no border addressing
};

struct Lumb200Device {
  int cuda_index      = 0;
  cudaStream_t stream = nullptr;
  int num_sms         = 0;
  int trace_grid      = 0;
  int stream_grid     = 0;
  int shade_grid      = 0;  // k_shade: many more blocks than are resident (see lumb200_device_create)

  std::vector<MeshDev> meshes;
  std::vector<Lumb200Instance> instances;
  std::vector<uint8_t> materials_packed;  // 32 bytes each
  uint32_t num_materials = 0;
  std::vector<TextureDev> textures;
  LbTexture* d_textures  = nullptr;  // DeviceTextureObject[] (16 bytes each)
  bool any_albedo_tex    = false;    // some material has an albedo texture: textured any-hit kernel variants
  bool any_material_tex  = false;    // some material references any texture: textured shading variant

  // scene tables on the device
  float4** d_mesh_vertices       = nullptr;
  uint4** d_mesh_textris         = nullptr;
  uint32_t* d_instance_mesh      = nullptr;
  LbTransform* d_instance_xform  = nullptr;
  uint32_t* d_instance_offset    = nullptr;
  uint2* d_prim_handle           = nullptr;
  uint16_t* d_prim_material      = nullptr;
  uint4* d_materials             = nullptr;
  float4* d_shadow_tab           = nullptr;
  uint16_t* d_sort_rank          = nullptr;  // material -> sort bin: rank in (class, material id) order (wavefront.cuh: LB_CLASS_*)
  LbSortClasses sort_classes     = {{0, LB_SORT_KEY_SKY, LB_SORT_KEY_SKY, LB_SORT_KEY_SKY}};
  uint32_t class_materials[LB_NUM_CLASSES] = {0, 0, 0};
  float4* d_world_tris           = nullptr;  // flattened order (kept: emitters / shading re-use)
  uint32_t num_prims             = 0;
  bool accel_dirty               = true;

  LbBvhBuffers bvh;
  LbBvhBuffers light_bvh;
  float4* d_light_world = nullptr;  // 3 float4 per light id (w of v0 = light id)

  // light tree blobs
  void* d_light_root          = nullptr;
  float4* d_light_root_children = nullptr;  // decoded root children (2 x float4 each)
  float4* d_light_records       = nullptr;  // per-light world-space triangle + colour (4 x float4 each)
  bool light_records_dirty      = true;
  void* d_light_nodes         = nullptr;
  uint2* d_light_handles      = nullptr;
  uint32_t* d_light_prims     = nullptr;  // light id -> flattened prim
  std::vector<uint32_t> light_handles_host;
  uint32_t num_lights         = 0;
  uint32_t light_root_bytes   = 0;
  uint32_t light_root_sections = 0;

  uint32_t* d_bluenoise = nullptr;
  uint16_t* d_bluenoise_1d = nullptr;  // dither mask of the output chain
  uchar4* d_output      = nullptr;     // ARGB8 staging of the output chain (width * height)
  std::vector<float*> bloom_mips;      // DevicePost.bloom_mips (device_post.c:43-60), allocated on first use
  uint32_t bloom_w = 0, bloom_h = 0;
  float* d_peer_planes  = nullptr;     // landing buffer of lumb200_device_add_planes_from
  size_t peer_floats    = 0;
  uint4* d_rng_table    = nullptr;  // per pass: Sobol pair + blue-noise offset of every (depth, target) dimension

  // BSDF LUTs
  LbLutTextures luts;

  Lumb200Settings settings = {0, 0, 0, 1};
  bool tile_order = true;  // render passes enumerate the frame in 8 x 4 tiles (k_raygen); LUMB200_TILE_ORDER=0 restores row order (A/B)
  uint32_t shading_mode = 0;  // LuminaryShadingMode: 0 the path tracer, 1..5 the one-bounce debug queue (device_renderer.c:136-182)
  LbCameraDev camera;
  Lumb200Sky sky;  // lumb200_sky_default() at create: constant colour (1, 1, 1)
  // procedural atmosphere (SkyLUT / DeviceSkyLUT / SkyStars, device_sky.h): LUT tables + texture objects, star catalogue
  LbSkyDev sky_dev      = {};
  float4* d_sky_lut[4]  = {nullptr, nullptr, nullptr, nullptr};  // transmittance low / high (256 x 64), multiscattering low / high (32 x 32)
  cudaArray_t sky_arrays[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaTextureObject_t sky_tex[4] = {0, 0, 0, 0};
  bool sky_lut_valid    = false;
  Lumb200Sky sky_lut_params;   // the medium the tables were built for
  float4* d_stars       = nullptr;
  uint32_t* d_stars_offsets = nullptr;
  std::vector<float> stars_host;
  std::vector<uint32_t> stars_offsets_host;
  uint32_t stars_count = 0xFFFFFFFFu, stars_seed = 0;
  // HDRI mode (SkyHDRI / DeviceSkyHDRI, device_sky.h): the baked table, its texture and what it was baked for
  float4* d_hdri          = nullptr;
  cudaArray_t hdri_array  = nullptr;
  cudaTextureObject_t hdri_tex = 0;
  uint32_t hdri_dim       = 0;
  bool hdri_valid         = false;
  float hdri_origin[3]    = {0.0f, 0.0f, 0.0f};
  // moon surface textures (DeviceEmbeddedData.moon_albedo_tex / moon_normal_tex)
  cudaArray_t moon_array[2]        = {nullptr, nullptr};
  cudaTextureObject_t moon_tex[2]  = {0, 0};

  // wavefront state
  LbPaths paths      = {};
  uint32_t* queue[2] = {nullptr, nullptr};
  uint32_t* sort_bins = nullptr;
  LbCounters* counters = nullptr;
  float2* d_uv        = nullptr;
  uint32_t paths_capacity = 0;

  // adaptive sampler (AdaptiveSampler + DeviceAdaptiveSampler, device/device_adaptive_sampler.h)
  Lumb200AdaptiveSampling as_params = {0, 256, 2, 64, 1, 1.0f, 4, 1.0f, 1.0f, 1.0f, 0};
  bool as_active            = false;  // latched at start_render
  uint32_t as_stage         = 0;
  uint32_t as_executions[LB_ADAPTIVE_STAGES + 1] = {0, 0, 0, 0, 0};
  uint32_t as_total_tasks   = 0;
  uint32_t as_bw = 0, as_bh = 0;
  uint32_t* d_as_words      = nullptr;
  uint32_t* d_as_prefix     = nullptr;
  uint32_t* d_as_total      = nullptr;
  float* d_as_block_var     = nullptr;
  float* d_as_var_sum       = nullptr;
  uint64_t as_paths         = 0;

  // accumulation
  float* planes          = nullptr;
  size_t planes_floats   = 0;
  bool planes_external   = false;
  float* d_result        = nullptr;
  // asynchronous result download: two staging buffers, a copy stream, one event pair per slot
  float* d_result_async[2]     = {nullptr, nullptr};
  cudaStream_t copy_stream     = nullptr;
  cudaStream_t aux_stream      = nullptr;  // the small material classes of a bounce are shaded beside the large one (lb_launch_shade)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool shade_overlap           = true;   // LUMB200_SHADE_OVERLAP=0: one stream (A/B, profiles/r2at_shade_overlap.txt)
  cudaEvent_t ev_resolved[2]   = {nullptr, nullptr};
  cudaEvent_t ev_copied[2]     = {nullptr, nullptr};
  bool slot_pending[2]         = {false, false};

  // per-kernel-class profiling
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;  // pairs
  std::vector<int> prof_class;
  size_t prof_used = 0;
  Lumb200Profile profile = {};

  // timing
  std::vector<cudaEvent_t> ev_start, ev_end;
  size_t events_pending = 0;
  double render_seconds = 0.0;
  double accel_seconds  = 0.0;
  uint64_t launches     = 0;
  uint32_t samples_done = 0;
  uint64_t device_bytes = 0;
  std::unordered_map<void*, size_t> alloc_bytes;  // live dev_alloc allocations, so that dev_free keeps device_bytes exact
};

template <typename T>
static Lumb200Result dev_alloc(Lumb200Device* d, T** ptr, size_t count) {
  *ptr             = nullptr;
  const size_t bytes = sizeof(T) * (count ? count : 1);
  cudaError_t e    = cudaMalloc((void**) ptr, bytes);
  if (e != cudaSuccess) {
    lumb200_set_last_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return LUMB200_ERROR_OUT_OF_MEMORY;
  }
  d->device_bytes += bytes;
  d->alloc_bytes[(void*) *ptr] = bytes;
  return LUMB200_SUCCESS;
}

template <typename T>
static void dev_free(Lumb200Device* d, T*& ptr) {
  if (ptr) {
    auto it = d->alloc_bytes.find((void*) ptr);
    if (it != d->alloc_bytes.end()) {
      d->device_bytes -= it->second;
      d->alloc_bytes.erase(it);
    }
    cudaFree(ptr);
  }
  ptr = nullptr;
}

// scratch device allocation of the parity / measurement hooks: released on every exit path
struct DevTmp {
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <typename T>
  T* as() const {
    return (T*) p;
  }
  ~DevTmp() {
    if (p)
      cudaFree(p);
  }
};

static Lumb200Result make_current(Lumb200Device* d) {
  LB_CHECK(cudaSetDevice(d->cuda_index));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_get_device_count(uint32_t* count) {
  LB_REQUIRE(count, LUMB200_ERROR_ARGUMENT_NULL, "count is NULL");
  int n         = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    lumb200_set_last_error("cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    return LUMB200_ERROR_CUDA;
  }
  *count = (uint32_t) n;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_get_device_properties(uint32_t cuda_index, char* name, size_t name_capacity, size_t* memory_bytes) {
  LB_REQUIRE(name && memory_bytes && name_capacity > 0, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  name[0]       = '\0';
  *memory_bytes = 0;
  cudaDeviceProp prop;
  LB_CHECK(cudaGetDeviceProperties(&prop, (int) cuda_index));
  snprintf(name, name_capacity, "%s", prop.name);
  *memory_bytes = prop.totalGlobalMem;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_create(Lumb200Device** device, uint32_t cuda_index) {
  LB_REQUIRE(device, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  *device = nullptr;
  uint32_t count;
  LB_TRY(lumb200_get_device_count(&count));
  LB_REQUIRE(cuda_index < count, LUMB200_ERROR_INVALID_DEVICE, "CUDA device %u does not exist (%u devices)", cuda_index, count);

  Lumb200Device* d = new Lumb200Device();
  d->cuda_index    = (int) cuda_index;
  lumb200_sky_default(&d->sky);
  d->sky_lut_params = d->sky;
  if (make_current(d) != LUMB200_SUCCESS) {
    delete d;
    return LUMB200_ERROR_CUDA;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, d->cuda_index) != cudaSuccess || cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) {
    lumb200_set_last_error("failed to initialise CUDA device %u", cuda_index);
    delete d;
    return LUMB200_ERROR_CUDA;
  }
  d->num_sms = prop.multiProcessorCount;
  // persistent grids: a multiple of the SM count (148 on B200)
  d->trace_grid  = d->num_sms * 8;
  d->stream_grid = d->num_sms * 8;
  // k_shade divides its class range statically over the grid, vertices differ a lot in cost and only 5 (opaque classes) or 4 (generic)
  // blocks are resident per SM: with 8 blocks per SM the second, partial wave ran at 3 / 5 of the occupancy. Many small blocks let the
  // hardware block scheduler balance instead (measured, profiles/r2am_shade_grid.txt: 8 -> 2.99 ms, 5 -> 2.90, 40 -> 2.74, 80 -> 2.72,
  // 160 -> 2.76, 640 -> 3.33 ms of k_shade per atrium-1M pass; blocks past the end of a range exit before the staging).
  d->shade_grid  = d->num_sms * 64;
  if (const char* e = getenv("LUMB200_SHADE_OVERLAP"))
    d->shade_overlap = atoi(e) != 0;
  if (d->shade_overlap && (cudaStreamCreateWithFlags(&d->aux_stream, cudaStreamNonBlocking) != cudaSuccess ||
                           cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                           cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming) != cudaSuccess))
    d->shade_overlap = false;
  if (const char* e = getenv("LUMB200_TILE_ORDER"))  // tuning experiments only
    d->tile_order = atoi(e) != 0;
  if (const char* e = getenv("LUMB200_SHADE_BLOCKS_PER_SM"))  // tuning experiments only
    d->shade_grid = d->num_sms * (atoi(e) > 0 ? atoi(e) : 64);

  // default camera (reference camera.c:10-64)
  memset(&d->camera, 0, sizeof(d->camera));
  d->camera.qw              = 1.0f;
  d->camera.fov             = 1.0f;
  d->camera.object_distance = 1.0f;
  d->camera.camera_scale    = 1.0f;
  d->camera.rr_threshold    = 0.1f;
  d->camera.aperture_blade_count = 7;

  Lumb200Result r = dev_alloc(d, &d->counters, 1);
  if (r == LUMB200_SUCCESS)
    r = dev_alloc(d, &d->sort_bins, 2 * LB_SORT_BINS);
  if (r == LUMB200_SUCCESS && cudaMemset(d->sort_bins, 0, 2 * LB_SORT_BINS * sizeof(uint32_t)) != cudaSuccess)  // k_sort_scan keeps the count bins zero
    r = LUMB200_ERROR_CUDA;
  if (r == LUMB200_SUCCESS)
    r = dev_alloc(d, &d->d_rng_table, (size_t) LB_RNG_TABLE_DEPTHS * LB_RNG_TARGET_COUNT);
  if (r != LUMB200_SUCCESS) {
    delete d;
    return r;
  }
  cudaMemsetAsync(d->counters, 0, sizeof(LbCounters), d->stream);
  *device = d;
  return LUMB200_SUCCESS;
}

static void free_paths(Lumb200Device* d) {
  dev_free(d, d->paths.org);
  dev_free(d, d->paths.dir);
  dev_free(d, d->paths.prim);
  dev_free(d, d->paths.record);
  dev_free(d, d->paths.pixel);
  dev_free(d, d->paths.state);
  dev_free(d, d->paths.medium);
  dev_free(d, d->paths.result);
  dev_free(d, d->paths.sample_id);
  dev_free(d, d->paths.nee);
  dev_free(d, d->paths.sq_org);
  dev_free(d, d->paths.sq_dir);
  dev_free(d, d->paths.sq_col);
  dev_free(d, d->paths.eq_org);
  dev_free(d, d->paths.eq_dir);
  dev_free(d, d->paths.eq_weight);
  dev_free(d, d->paths.eq_rec);
  dev_free(d, d->paths.eq_hits);
  dev_free(d, d->queue[0]);
  dev_free(d, d->queue[1]);
  dev_free(d, d->d_uv);
  d->paths_capacity = 0;
}

static void free_scene_tables(Lumb200Device* d) {
  dev_free(d, d->d_mesh_vertices);
  dev_free(d, d->d_mesh_textris);
  dev_free(d, d->d_instance_mesh);
  dev_free(d, d->d_instance_xform);
  dev_free(d, d->d_instance_offset);
  dev_free(d, d->d_prim_handle);
  dev_free(d, d->d_prim_material);
  dev_free(d, d->d_world_tris);
  dev_free(d, d->d_light_world);
  dev_free(d, d->d_light_prims);
  d->device_bytes -= d->bvh.bytes < d->device_bytes ? d->bvh.bytes : d->device_bytes;
  lb_bvh8_free(&d->bvh);
  d->device_bytes -= d->light_bvh.bytes < d->device_bytes ? d->light_bvh.bytes : d->device_bytes;
  lb_bvh8_free(&d->light_bvh);
}

extern "C" Lumb200Result lumb200_device_destroy(Lumb200Device** device) {
  LB_REQUIRE(device, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  Lumb200Device* d = *device;
  if (!d)
    return LUMB200_SUCCESS;
  make_current(d);
  cudaStreamSynchronize(d->stream);
  if (d->copy_stream)
    cudaStreamSynchronize(d->copy_stream);
  free_paths(d);
  free_scene_tables(d);
  for (MeshDev& m : d->meshes) {
    dev_free(d, m.vertices);
    dev_free(d, m.textris);
  }
  dev_free(d, d->d_materials);
  dev_free(d, d->d_shadow_tab);
  dev_free(d, d->d_sort_rank);
  for (TextureDev& t : d->textures) {
    if (t.obj)
      cudaDestroyTextureObject(t.obj);
    if (t.array)
      cudaFreeArray(t.array);
    if (t.mips)
      cudaFreeMipmappedArray(t.mips);
  }
  dev_free(d, d->d_textures);
  for (int k = 0; k < 4; k++) {
    if (d->sky_tex[k])
      cudaDestroyTextureObject(d->sky_tex[k]);
    if (d->sky_arrays[k])
      cudaFreeArray(d->sky_arrays[k]);
    dev_free(d, d->d_sky_lut[k]);
  }
  dev_free(d, d->d_stars);
  dev_free(d, d->d_stars_offsets);
  if (d->hdri_tex)
    cudaDestroyTextureObject(d->hdri_tex);
  if (d->hdri_array)
    cudaFreeArray(d->hdri_array);
  dev_free(d, d->d_hdri);
  for (int k = 0; k < 2; k++) {
    if (d->moon_tex[k])
      cudaDestroyTextureObject(d->moon_tex[k]);
    if (d->moon_array[k])
      cudaFreeArray(d->moon_array[k]);
  }
  dev_free(d, d->d_light_root);
  dev_free(d, d->d_light_root_children);
  dev_free(d, d->d_light_records);
  dev_free(d, d->d_light_nodes);
  dev_free(d, d->d_light_handles);
  dev_free(d, d->d_bluenoise);
  dev_free(d, d->d_rng_table);
  dev_free(d, d->d_bluenoise_1d);
  dev_free(d, d->d_output);
  for (float*& m : d->bloom_mips)
    dev_free(d, m);
  dev_free(d, d->d_peer_planes);
  dev_free(d, d->d_as_words);
  dev_free(d, d->d_as_prefix);
  dev_free(d, d->d_as_total);
  dev_free(d, d->d_as_block_var);
  dev_free(d, d->d_as_var_sum);
  dev_free(d, d->counters);
  dev_free(d, d->sort_bins);
  dev_free(d, d->d_result);
  for (int k = 0; k < 2; k++) {
    dev_free(d, d->d_result_async[k]);
    if (d->ev_resolved[k])
      cudaEventDestroy(d->ev_resolved[k]);
    if (d->ev_copied[k])
      cudaEventDestroy(d->ev_copied[k]);
  }
  if (d->copy_stream)
    cudaStreamDestroy(d->copy_stream);
  if (d->aux_stream)
    cudaStreamDestroy(d->aux_stream);
  if (d->ev_fork)
    cudaEventDestroy(d->ev_fork);
  if (d->ev_join)
    cudaEventDestroy(d->ev_join);
  if (!d->planes_external)
    dev_free(d, d->planes);
  lb_lut_destroy(&d->luts);
  for (cudaEvent_t e : d->prof_events)
    cudaEventDestroy(e);
  for (cudaEvent_t e : d->ev_start)
    cudaEventDestroy(e);
  for (cudaEvent_t e : d->ev_end)
    cudaEventDestroy(e);
  cudaStreamDestroy(d->stream);
  delete d;
  *device = nullptr;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_load_bluenoise(Lumb200Device* d, const uint32_t* bluenoise_2d, size_t count) {
  LB_REQUIRE(d && bluenoise_2d, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(count == 256 * 256, LUMB200_ERROR_INVALID_API_ARGUMENT, "blue-noise mask must be 256x256 uint32 (got %zu entries)", count);
  LB_TRY(make_current(d));
  if (!d->d_bluenoise)
    LB_TRY(dev_alloc(d, &d->d_bluenoise, count));
  LB_CHECK(cudaMemcpyAsync(d->d_bluenoise, bluenoise_2d, count * sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  return LUMB200_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// packing helpers (host side of the reference: device_packing.c, device_structs.c)
// ---------------------------------------------------------------------------------------------
static uint32_t pack_normal_host(float nx, float ny, float nz) {  // device_packing.c:6-31
  double x = nx, y = ny, z = nz;
  const double rn = 1.0 / (fabs(x) + fabs(y) + fabs(z));
  x *= rn, y *= rn, z *= rn;
  const double t = fmax(fmin(-z, 1.0), 0.0);
  x += (x >= 0.0) ? t : -t;
  y += (y >= 0.0) ? t : -t;
  x = fmax(fmin(x, 1.0), -1.0);
  y = fmax(fmin(y, 1.0), -1.0);
  x = (x + 1.0) * 0.5;
  y = (y + 1.0) * 0.5;
  return (((uint32_t) (y * 0xFFFF + 0.5)) << 16) | ((uint32_t) (x * 0xFFFF + 0.5));
}

static uint32_t float_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

static uint32_t pack_uv_host(float u, float v) { return (float_bits(u) & 0xFFFF0000u) | (float_bits(v) >> 16); }  // device_packing.c:36-43

static uint16_t f01_u16(float f) { return (uint16_t) (f * 65535.0f + 0.5f); }  // device_structs.c:250-252

struct MaterialPacked {  // DeviceMaterialCompressed, device_structs.h:232-254
  uint8_t flags;
  uint8_t roughness_clamp;
  uint16_t metallic_tex;
  uint16_t roughness;
  uint16_t refraction_index;
  uint16_t albedo_r, albedo_g, albedo_b, albedo_a;
  uint16_t emission_r, emission_g, emission_b, emission_scale;
  uint16_t albedo_tex, luminance_tex, roughness_tex, normal_tex;
};
static_assert(sizeof(MaterialPacked) == 32, "DeviceMaterialCompressed is 32 bytes");

static void pack_material(const Lumb200Material& m, MaterialPacked& d) {  // device_structs.c:257-330
  memset(&d, 0, sizeof(d));
  d.flags |= m.emission_active ? 0x02 : 0;
  d.flags |= m.thin_walled ? 0x04 : 0;
  d.flags |= m.metallic ? 0x08 : 0;
  d.flags |= m.colored_transparency ? 0x10 : 0;
  d.flags |= m.roughness_as_smoothness ? 0x20 : 0;
  d.flags |= m.normal_map_is_compressed ? 0x40 : 0;
  d.flags |= m.bidirectional_emission ? 0x80 : 0;
  d.flags |= (m.base_substrate == 1) ? 0x01 : 0;
  d.roughness_clamp  = (uint8_t) (f01_u16(m.roughness_clamp) >> 8);
  d.roughness        = f01_u16(m.roughness);
  d.refraction_index = f01_u16(0.5f * (m.refraction_index - 1.0f));
  float er = m.emission[0], eg = m.emission[1], eb = m.emission[2];
  const float en = 1.0f / fminf(fmaxf(fmaxf(er, eg), eb) + 1.0f, 65535.0f);
  er *= en, eg *= en, eb *= en;
  d.albedo_r       = f01_u16(m.albedo[0]);
  d.albedo_g       = f01_u16(m.albedo[1]);
  d.albedo_b       = f01_u16(m.albedo[2]);
  d.albedo_a       = f01_u16(m.albedo[3]);
  d.emission_r     = f01_u16(er);
  d.emission_g     = f01_u16(eg);
  d.emission_b     = f01_u16(eb);
  d.emission_scale = (uint16_t) ((float_bits(m.emission_scale / en) >> 15) & 0xFFFF);
  d.albedo_tex    = m.albedo_tex;
  d.luminance_tex = m.luminance_tex;
  d.roughness_tex = m.roughness_tex;
  d.metallic_tex  = m.metallic_tex;
  d.normal_tex    = m.normal_tex;
}

static void euler_to_quat(const float rot[3], float q[4]) {  // host_math.c:6-21 -> (x, y, z, w)
  const float cr = cosf(rot[0] * 0.5f), sr = sinf(rot[0] * 0.5f);
  const float cp = cosf(rot[1] * 0.5f), sp = sinf(rot[1] * 0.5f);
  const float cy = cosf(rot[2] * 0.5f), sy = sinf(rot[2] * 0.5f);
  q[3] = cr * cp * cy + sr * sp * sy;
  q[0] = sr * cp * cy - cr * sp * sy;
  q[1] = cr * sp * cy + sr * cp * sy;
  q[2] = cr * cp * sy - sr * sp * cy;
}

// ---------------------------------------------------------------------------------------------
// scene upload
// ---------------------------------------------------------------------------------------------
static void pack_triangles(const Lumb200Mesh* mesh, float4* hv, uint4* ht) {  // device_struct_triangles_convert, device_structs.c:332-386
  for (size_t t = 0; t < mesh->triangle_count; t++) {
    for (int v = 0; v < 3; v++) {
      const float* p  = mesh->vertex_buffer + 9 * t + 3 * v;
      const float* nn = mesh->normal_buffer + 9 * t + 3 * v;
      float4 o;
      o.x = p[0], o.y = p[1], o.z = p[2];
      const uint32_t pn = pack_normal_host(nn[0], nn[1], nn[2]);
      memcpy(&o.w, &pn, 4);
      hv[3 * t + v] = o;
    }
    const float* uv = mesh->uv_buffer + 6 * t;
    ht[t]           = make_uint4(pack_uv_host(uv[0], uv[1]), pack_uv_host(uv[2], uv[3]), pack_uv_host(uv[4], uv[5]), mesh->material_id_buffer[t]);
  }
}

static LbTransform pack_transform(const Lumb200Instance& in) {  // device_struct_instance_transform_convert, device_structs.c:402-413 (+ quaternion16 :388-399)
  float q[4];
  euler_to_quat(in.rotation, q);
  LbTransform t;
  t.tx = in.translation[0], t.ty = in.translation[1], t.tz = in.translation[2];
  t.sx = in.scale[0], t.sy = in.scale[1], t.sz = in.scale[2];
  t.qx = (uint16_t) (((1.0f - q[0]) * 0x7FFF) + 0.5f);
  t.qy = (uint16_t) (((1.0f - q[1]) * 0x7FFF) + 0.5f);
  t.qz = (uint16_t) (((1.0f - q[2]) * 0x7FFF) + 0.5f);
  t.qw = (uint16_t) (((1.0f + q[3]) * 0x7FFF) + 0.5f);
  return t;
}

extern "C" Lumb200Result lumb200_host_pack_material(const Lumb200Material* material, void* dst32) {
  LB_REQUIRE(material && dst32, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  MaterialPacked m;
  pack_material(*material, m);
  memcpy(dst32, &m, sizeof(m));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_host_pack_triangles(const Lumb200Mesh* mesh, void* vertices48, void* textris16) {
  LB_REQUIRE(mesh && vertices48 && textris16, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(mesh->triangle_count == 0 || (mesh->vertex_buffer && mesh->normal_buffer && mesh->uv_buffer && mesh->material_id_buffer),
             LUMB200_ERROR_ARGUMENT_NULL, "mesh buffers are NULL");
  pack_triangles(mesh, (float4*) vertices48, (uint4*) textris16);
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_host_pack_transform(const Lumb200Instance* instance, void* dst32) {
  LB_REQUIRE(instance && dst32, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  const LbTransform t = pack_transform(*instance);
  memcpy(dst32, &t, sizeof(t));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_add_mesh(Lumb200Device* d, const Lumb200Mesh* mesh, uint32_t* mesh_id) {
  LB_REQUIRE(d && mesh, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  const uint32_t n = mesh->triangle_count;
  LB_REQUIRE(n == 0 || (mesh->vertex_buffer && mesh->normal_buffer && mesh->uv_buffer && mesh->material_id_buffer), LUMB200_ERROR_ARGUMENT_NULL,
             "mesh buffers are NULL");
  LB_REQUIRE(n < 0x7FFFFFFFu, LUMB200_ERROR_INVALID_API_ARGUMENT, "mesh has too many triangles");  // HIT_TYPE_TRIANGLE_ID_LIMIT
  LB_TRY(make_current(d));

  MeshDev md;
  md.num_tris = n;
  LB_TRY(dev_alloc(d, &md.vertices, 3 * (size_t) n));
  if (Lumb200Result r = dev_alloc(d, &md.textris, (size_t) n)) {
    dev_free(d, md.vertices);
    return r;
  }

  // device_mesh_set (device_mesh.c:19-51): 3 x {pos, packed normal} + {3 packed uv, material}
  std::vector<float4> hv(3 * (size_t) n);
  std::vector<uint4> ht(n);
  md.host_material.resize(n);
  pack_triangles(mesh, hv.data(), ht.data());
  for (size_t t = 0; t < n; t++) {
    md.host_material[t] = mesh->material_id_buffer[t];
    if (md.host_material[t] > md.max_material)
      md.max_material = md.host_material[t];
  }
  if (n) {
    LB_CHECK(cudaMemcpyAsync(md.vertices, hv.data(), sizeof(float4) * hv.size(), cudaMemcpyHostToDevice, d->stream));
    LB_CHECK(cudaMemcpyAsync(md.textris, ht.data(), sizeof(uint4) * ht.size(), cudaMemcpyHostToDevice, d->stream));
    LB_CHECK(cudaStreamSynchronize(d->stream));
  }
  if (mesh_id)
    *mesh_id = (uint32_t) d->meshes.size();
  d->meshes.push_back(std::move(md));
  d->accel_dirty = true;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_update_instances(Lumb200Device* d, const Lumb200Instance* instances, uint32_t count) {
  LB_REQUIRE(d && (instances || count == 0), LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  for (uint32_t i = 0; i < count; i++)
    LB_REQUIRE(instances[i].mesh_id < d->meshes.size(), LUMB200_ERROR_INVALID_API_ARGUMENT, "instance %u references mesh %u which does not exist",
               i, instances[i].mesh_id);
  d->instances.assign(instances, instances + count);
  d->accel_dirty = true;
  return LUMB200_SUCCESS;
}

static Lumb200Result upload_materials(Lumb200Device* d) {
  LB_TRY(make_current(d));
  d->light_records_dirty = true;  // the records cache the emitters' material colour
  dev_free(d, d->d_materials);
  dev_free(d, d->d_shadow_tab);
  dev_free(d, d->d_sort_rank);
  const uint32_t n = d->num_materials;
  LB_TRY(dev_alloc(d, &d->d_materials, 2 * (size_t) n));
  LB_TRY(dev_alloc(d, &d->d_shadow_tab, (size_t) n));
  LB_TRY(dev_alloc(d, &d->d_sort_rank, (size_t) n));
  std::vector<float4> tab(n ? n : 1);
  const MaterialPacked* mp = (const MaterialPacked*) d->materials_packed.data();
  d->any_albedo_tex        = false;
  d->any_material_tex      = false;
  for (uint32_t i = 0; i < n; i++) {
    if (mp[i].albedo_tex != 0xFFFF || mp[i].luminance_tex != 0xFFFF || mp[i].roughness_tex != 0xFFFF || mp[i].normal_tex != 0xFFFF ||
        mp[i].metallic_tex != 0xFFFF)
      d->any_material_tex = true;
    // shadow any-hit response of a material (cuda/optix_anyhit.cuh:49-93): albedo decoded like load_material
    const float inv = 1.0f / 0xFFFF;
    const float r = mp[i].albedo_r * inv, g = mp[i].albedo_g * inv, b = mp[i].albedo_b * inv, a = mp[i].albedo_a * inv;
    const bool colored = (mp[i].flags & 0x10) != 0;
    float4 t;
    const bool tex_opaque = mp[i].albedo_tex < d->textures.size() && d->textures[mp[i].albedo_tex].fully_opaque;
    if (mp[i].albedo_tex != 0xFFFF && tex_opaque)
      t = make_float4(0.0f, 0.0f, 0.0f, 3.0f);  // opaque whatever the texel (alpha == 1 everywhere): blocks, and is never cut out
    else if (mp[i].albedo_tex != 0xFFFF) {
      t                 = make_float4(0.0f, 0.0f, 0.0f, 2.0f);  // evaluated per hit from the albedo texture (k_trace_shadow<*, true>)
      d->any_albedo_tex = true;
    }
    else if (a == 1.0f)
      t = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    else if (a == 0.0f && !colored)
      t = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
    else {
      const float tr = 1.0f - a;
      t              = colored ? make_float4(r * tr, g * tr, b * tr, 0.0f) : make_float4(tr, tr, tr, 0.0f);
    }
    tab[i] = t;
  }
  // Material classes and sort bins (wavefront.cuh). A material is shaded by an opaque-class kernel only when nothing it can evaluate
  // to needs the generic code: opaque substrate, stored opacity exactly 1 and no albedo texture (a texture supplies its own alpha).
  // "Metallic" follows geometry_get_context: a material WITH a metallic map is not metallic (geometry_utils.cuh:160-162).
  std::vector<uint8_t> cls(n ? n : 1, LB_CLASS_GENERIC);
  const bool classify = n <= LB_SORT_KEY_SKY;  // one bin per material below the sky bin; larger scenes shade everything as GENERIC
  d->class_materials[0] = d->class_materials[1] = d->class_materials[2] = 0;
  for (uint32_t i = 0; i < n; i++) {
    const bool translucent = (mp[i].flags & 0x01) != 0;
    const bool opaque      = mp[i].albedo_a == 0xFFFF && mp[i].albedo_tex == 0xFFFF;
    const bool metallic    = (mp[i].flags & 0x08) != 0 && mp[i].metallic_tex == 0xFFFF;
    if (classify && !translucent && opaque)
      cls[i] = metallic ? LB_CLASS_METAL : LB_CLASS_DIELECTRIC;
    d->class_materials[cls[i]]++;
  }
  std::vector<uint16_t> rank(n ? n : 1, 0);
  {
    uint32_t next = 0;
    for (uint32_t c = 0; c < LB_NUM_CLASSES; c++) {
      d->sort_classes.first_rank[c] = classify ? next : (c == 0 ? 0u : LB_SORT_KEY_SKY);
      for (uint32_t i = 0; i < n; i++)
        if (cls[i] == c) {
          rank[i] = (uint16_t) (next < LB_SORT_KEY_SKY - 1u ? next : LB_SORT_KEY_SKY - 1u);
          next++;
        }
    }
    d->sort_classes.first_rank[LB_NUM_CLASSES] = LB_SORT_KEY_SKY;
    // an empty trailing class starts where the sky bin starts; empty leading / middle classes share their successor's first bin
    for (uint32_t c = 0; c < LB_NUM_CLASSES; c++)
      if (classify && d->sort_classes.first_rank[c] >= n)
        d->sort_classes.first_rank[c] = LB_SORT_KEY_SKY;
  }
  if (n) {
    LB_CHECK(cudaMemcpyAsync(d->d_materials, d->materials_packed.data(), 32 * (size_t) n, cudaMemcpyHostToDevice, d->stream));
    LB_CHECK(cudaMemcpyAsync(d->d_shadow_tab, tab.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, d->stream));
    LB_CHECK(cudaMemcpyAsync(d->d_sort_rank, rank.data(), sizeof(uint16_t) * n, cudaMemcpyHostToDevice, d->stream));
    LB_CHECK(cudaStreamSynchronize(d->stream));
  }
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_update_materials_packed(Lumb200Device* d, const void* materials, uint32_t count) {
  LB_REQUIRE(d && (materials || count == 0), LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(count <= 0xFFFF, LUMB200_ERROR_INVALID_API_ARGUMENT, "too many materials");
  d->materials_packed.assign((const uint8_t*) materials, (const uint8_t*) materials + 32 * (size_t) count);
  d->num_materials = count;
  return upload_materials(d);
}

extern "C" Lumb200Result lumb200_device_update_materials(Lumb200Device* d, const Lumb200Material* materials, uint32_t count) {
  LB_REQUIRE(d && (materials || count == 0), LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(count <= 0xFFFF, LUMB200_ERROR_INVALID_API_ARGUMENT, "too many materials");
  std::vector<MaterialPacked> packed(count);
  for (uint32_t i = 0; i < count; i++) {
    LB_REQUIRE(materials[i].base_substrate <= 1, LUMB200_ERROR_API_EXCEPTION, "Invalid base substrate.");
    pack_material(materials[i], packed[i]);
  }
  return lumb200_device_update_materials_packed(d, packed.data(), count);
}

// device_add_textures (device/device.h:160) -> device_texture_create (device/device_texture.c): CUDA array + texture object with
// normalised coordinates, the texture's address / filter modes and unorm reads of integer texels.
extern "C" Lumb200Result lumb200_device_add_textures(Lumb200Device* d, const Lumb200Texture* textures, uint32_t count) {
  LB_REQUIRE(d && (textures || count == 0), LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d->textures.size() + count <= 0xFFFF, LUMB200_ERROR_INVALID_API_ARGUMENT, "Exceeded limit of 65535 textures.");
  LB_TRY(make_current(d));
  // a texture that fails half-way must not leak its array / object: `guard` releases whatever `pending` still holds on any early return
  TextureDev td;
  TextureDev* pending = nullptr;
  struct Guard {
    TextureDev*& p;
    ~Guard() {
      if (!p)
        return;
      if (p->obj)
        cudaDestroyTextureObject(p->obj);
      if (p->array)
        cudaFreeArray(p->array);
      if (p->mips)
        cudaFreeMipmappedArray(p->mips);
    }
  } guard{pending};
  for (uint32_t i = 0; i < count; i++) {
    const Lumb200Texture& t = textures[i];
    td      = TextureDev();
    pending = &td;
    td.gamma  = t.gamma;
    td.width  = t.width;
    td.height = t.height;
    if (t.data) {
      LB_REQUIRE(t.width <= 0xFFFF && t.height <= 0xFFFF, LUMB200_ERROR_INVALID_API_ARGUMENT, "texture %u is larger than 65535 texels", i);
      LB_REQUIRE(t.width > 0 && t.height > 0, LUMB200_ERROR_INVALID_API_ARGUMENT, "texture %u has no extent", i);
      LB_REQUIRE(t.num_components == 1 || t.num_components == 2 || t.num_components == 4, LUMB200_ERROR_API_EXCEPTION,
                 "texture %u: %u components are not supported (1, 2 or 4)", i, t.num_components);
      LB_REQUIRE(t.type <= LUMB200_TEXTURE_U16, LUMB200_ERROR_API_EXCEPTION, "Texture data type is invalid.");
      LB_REQUIRE(t.wrap_mode_u <= LUMB200_WRAP_BORDER && t.wrap_mode_v <= LUMB200_WRAP_BORDER, LUMB200_ERROR_API_EXCEPTION,
                 "Texture wrapping mode is invalid.");
      LB_REQUIRE(t.filter <= LUMB200_FILTER_LINEAR, LUMB200_ERROR_API_EXCEPTION, "Texture filter mode is invalid.");
      const int bits         = (t.type == LUMB200_TEXTURE_U8) ? 8 : (t.type == LUMB200_TEXTURE_U16) ? 16 : 32;
      const size_t row_bytes = (size_t) t.width * t.num_components * (bits / 8);
      LB_REQUIRE(t.pitch >= row_bytes, LUMB200_ERROR_INVALID_API_ARGUMENT, "texture %u: pitch %u is smaller than a row (%zu bytes)", i, t.pitch,
                 row_bytes);
      if (t.num_components == 4 && t.wrap_mode_u != LUMB200_WRAP_BORDER && t.wrap_mode_v != LUMB200_WRAP_BORDER) {
        // any-hit shortcut: a texture whose alpha is 1 everywhere can neither cut a hit out nor let a shadow ray through, so
        // materials that use it keep their precomputed opaque response and never fetch in the traversal kernels
        bool opaque = true;
        for (uint32_t y = 0; y < t.height && opaque; y++) {
          const uint8_t* row = (const uint8_t*) t.data + (size_t) t.pitch * y;
          for (uint32_t x = 0; x < t.width && opaque; x++) {
            if (t.type == LUMB200_TEXTURE_U8)
              opaque = row[4 * x + 3] == 0xFF;
            else if (t.type == LUMB200_TEXTURE_U16)
              opaque = ((const uint16_t*) row)[4 * x + 3] == 0xFFFF;
            else
              opaque = ((const float*) row)[4 * x + 3] == 1.0f;
          }
        }
        td.fully_opaque = opaque;
      }
      const cudaChannelFormatKind kind = (t.type == LUMB200_TEXTURE_FP32) ? cudaChannelFormatKindFloat : cudaChannelFormatKindUnsigned;
      const int nc                     = (int) t.num_components;
      const cudaChannelFormatDesc desc = cudaCreateChannelDesc(bits, nc >= 2 ? bits : 0, nc == 4 ? bits : 0, nc == 4 ? bits : 0, kind);
      // _device_texture_get_num_mip_levels, device_texture.c:93-126: floor(log2(min(w, h))) levels, at least one
      uint32_t levels = 1;
      LB_REQUIRE(t.mipmap <= 1, LUMB200_ERROR_API_EXCEPTION, "Texture mipmap mode is invalid.");
      if (t.mipmap == 1 && nc == 4) {
        uint32_t min_dim = t.width < t.height ? t.width : t.height, l = 0;
        while (min_dim > 1) {
          l++;
          min_dim >>= 1;
        }
        levels = l ? l : 1;
      }
      td.num_levels = levels;
      cudaArray_t level0 = nullptr;
      if (levels > 1) {
        LB_CHECK(cudaMallocMipmappedArray(&td.mips, &desc, make_cudaExtent(t.width, t.height, 0), levels, cudaArraySurfaceLoadStore));
        LB_CHECK(cudaGetMipmappedArrayLevel(&level0, td.mips, 0));
      }
      else {
        LB_CHECK(cudaMallocArray(&td.array, &desc, t.width, t.height));
        level0 = td.array;
      }
      d->device_bytes += row_bytes * t.height;
      LB_CHECK(cudaMemcpy2DToArrayAsync(level0, 0, 0, t.data, t.pitch, row_bytes, t.height, cudaMemcpyHostToDevice, d->stream));
      LB_CHECK(cudaStreamSynchronize(d->stream));  // the caller keeps ownership of t.data
      cudaResourceDesc res;
      memset(&res, 0, sizeof(res));
      cudaTextureDesc tex;
      memset(&tex, 0, sizeof(tex));
      const cudaTextureAddressMode modes[4] = {cudaAddressModeWrap, cudaAddressModeClamp, cudaAddressModeMirror, cudaAddressModeBorder};
      tex.addressMode[0]      = modes[t.wrap_mode_u];
      tex.addressMode[1]      = modes[t.wrap_mode_v];
      tex.addressMode[2]      = cudaAddressModeClamp;
      tex.filterMode          = (t.filter == LUMB200_FILTER_LINEAR) ? cudaFilterModeLinear : cudaFilterModePoint;
      tex.readMode            = (t.type == LUMB200_TEXTURE_FP32) ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
      tex.normalizedCoords    = 1;
      tex.maxAnisotropy       = 1;
      tex.mipmapFilterMode    = cudaFilterModePoint;
      tex.minMipmapLevelClamp = 0.0f;
      tex.maxMipmapLevelClamp = (float) (levels - 1);
      // _device_texture_generate_mipmaps: level l + 1 from a texture object over level l (same sampler state) through a surface
      for (uint32_t l = 0; l + 1 < levels; l++) {
        cudaArray_t src_level = nullptr, dst_level = nullptr;
        LB_CHECK(cudaGetMipmappedArrayLevel(&src_level, td.mips, l));
        LB_CHECK(cudaGetMipmappedArrayLevel(&dst_level, td.mips, l + 1));
        cudaResourceDesc lr;
        memset(&lr, 0, sizeof(lr));
        lr.resType         = cudaResourceTypeArray;
        lr.res.array.array = src_level;
        cudaTextureDesc lt = tex;
        lt.maxMipmapLevelClamp = 0.0f;
        cudaTextureObject_t src_tex = 0;
        LB_CHECK(cudaCreateTextureObject(&src_tex, &lr, &lt, nullptr));
        lr.res.array.array = dst_level;
        cudaSurfaceObject_t dst_surf = 0;
        LB_CHECK(cudaCreateSurfaceObject(&dst_surf, &lr));
        lb_launch_mipmap_level(src_tex, dst_surf, t.width >> (l + 1), t.height >> (l + 1), t.type, d->stream);
        LB_CHECK(cudaStreamSynchronize(d->stream));
        cudaDestroyTextureObject(src_tex);
        cudaDestroySurfaceObject(dst_surf);
        d->launches++;
        d->device_bytes += (row_bytes * t.height) >> (2 * (l + 1));
      }
      if (levels > 1) {
        res.resType           = cudaResourceTypeMipmappedArray;
        res.res.mipmap.mipmap = td.mips;
      }
      else {
        res.resType         = cudaResourceTypeArray;
        res.res.array.array = td.array;
      }
      LB_CHECK(cudaCreateTextureObject(&td.obj, &res, &tex, nullptr));
    }
    d->textures.push_back(td);
    pending = nullptr;
  }
  std::vector<LbTexture> table(d->textures.size() ? d->textures.size() : 1);
  for (size_t i = 0; i < d->textures.size(); i++) {
    table[i].handle = d->textures[i].obj;
    table[i].gamma  = d->textures[i].gamma;
    table[i].size   = d->textures[i].width | (d->textures[i].height << 16);
  }
  dev_free(d, d->d_textures);
  LB_TRY(dev_alloc(d, &d->d_textures, table.size()));
  LB_CHECK(cudaMemcpyAsync(d->d_textures, table.data(), sizeof(LbTexture) * table.size(), cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  d->light_records_dirty = true;
  if (d->num_materials)
    LB_TRY(upload_materials(d));  // the per-material any-hit responses depend on the textures
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_compute_light_intensities(Lumb200Device* d, const uint32_t* mesh_ids, const uint32_t* triangle_ids,
                                                                  uint32_t count, float* intensities) {
  LB_REQUIRE(d && (count == 0 || (mesh_ids && triangle_ids && intensities)), LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  if (count == 0)
    return LUMB200_SUCCESS;
  LB_REQUIRE(d->d_materials, LUMB200_ERROR_API_EXCEPTION, "no materials uploaded");
  for (uint32_t i = 0; i < count; i++)
    LB_REQUIRE(mesh_ids[i] < d->meshes.size() && triangle_ids[i] < d->meshes[mesh_ids[i]].num_tris, LUMB200_ERROR_INVALID_API_ARGUMENT,
               "entry %u references triangle %u of mesh %u which does not exist", i, triangle_ids[i], mesh_ids[i]);
  LB_TRY(make_current(d));
  std::vector<uint4*> mt(d->meshes.size());
  for (size_t m = 0; m < d->meshes.size(); m++)
    mt[m] = d->meshes[m].textris;
  uint4** d_mt   = nullptr;
  uint32_t* d_m  = nullptr;
  uint32_t* d_t  = nullptr;
  float* d_out   = nullptr;
  Lumb200Result r = dev_alloc(d, &d_mt, mt.size());
  if (r == LUMB200_SUCCESS)
    r = dev_alloc(d, &d_m, count);
  if (r == LUMB200_SUCCESS)
    r = dev_alloc(d, &d_t, count);
  if (r == LUMB200_SUCCESS)
    r = dev_alloc(d, &d_out, count);
  cudaError_t e = cudaSuccess;
  if (r == LUMB200_SUCCESS) {
    cudaMemcpyAsync(d_mt, mt.data(), sizeof(uint4*) * mt.size(), cudaMemcpyHostToDevice, d->stream);
    cudaMemcpyAsync(d_m, mesh_ids, sizeof(uint32_t) * count, cudaMemcpyHostToDevice, d->stream);
    cudaMemcpyAsync(d_t, triangle_ids, sizeof(uint32_t) * count, cudaMemcpyHostToDevice, d->stream);
    LbShadeParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.mesh_textris = (const uint4* const*) d_mt;
    sp.materials    = d->d_materials;
    sp.textures     = d->d_textures;
    sp.num_textures = (uint32_t) d->textures.size();
    lb_launch_light_compute_intensity(sp, d_m, d_t, count, d_out, d->stream);
    cudaMemcpyAsync(intensities, d_out, sizeof(float) * count, cudaMemcpyDeviceToHost, d->stream);
    e = cudaStreamSynchronize(d->stream);
    d->launches++;
  }
  dev_free(d, d_mt);
  dev_free(d, d_m);
  dev_free(d, d_t);
  dev_free(d, d_out);
  LB_TRY(r);
  LB_CHECK(e);
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_sample_texture(Lumb200Device* d, uint32_t texture_id, const float* uv, uint32_t count, float* rgba_out) {
  return lumb200_device_sample_texture_lod(d, texture_id, uv, count, 0.0f, rgba_out);
}

extern "C" Lumb200Result lumb200_device_sample_texture_lod(Lumb200Device* d, uint32_t texture_id, const float* uv, uint32_t count, float lod,
                                                           float* rgba_out) {
  LB_REQUIRE(d && uv && rgba_out, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(texture_id < d->textures.size(), LUMB200_ERROR_INVALID_API_ARGUMENT, "texture %u does not exist", texture_id);
  LB_TRY(make_current(d));
  float2* d_uv  = nullptr;
  float4* d_out = nullptr;
  LB_TRY(dev_alloc(d, &d_uv, count));
  LB_TRY(dev_alloc(d, &d_out, count));
  cudaMemcpyAsync(d_uv, uv, sizeof(float2) * count, cudaMemcpyHostToDevice, d->stream);
  lb_launch_sample_texture(d->d_textures, (uint32_t) d->textures.size(), texture_id, d_uv, count, lod, d_out, d->stream);
  cudaMemcpyAsync(rgba_out, d_out, sizeof(float4) * count, cudaMemcpyDeviceToHost, d->stream);
  const cudaError_t e = cudaStreamSynchronize(d->stream);
  dev_free(d, d_uv);
  dev_free(d, d_out);
  LB_CHECK(e);
  d->launches++;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_update_light_tree(Lumb200Device* d, const Lumb200LightTree* tree) {
  LB_REQUIRE(d && tree, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_TRY(make_current(d));
  dev_free(d, d->d_light_root);
  dev_free(d, d->d_light_root_children);
  dev_free(d, d->d_light_records);
  d->light_records_dirty = true;
  dev_free(d, d->d_light_nodes);
  dev_free(d, d->d_light_handles);
  d->num_lights       = 0;
  d->light_root_bytes = 0;
  d->light_root_sections = 0;
  d->light_handles_host.clear();
  d->accel_dirty = true;
  if (tree->num_lights == 0)
    return LUMB200_SUCCESS;
  LB_REQUIRE(tree->root_data && tree->root_size >= 16 && tree->tri_handle_map, LUMB200_ERROR_INVALID_API_ARGUMENT, "light tree blobs missing");
  uint8_t* root  = nullptr;
  uint8_t* nodes = nullptr;
  LB_TRY(dev_alloc(d, &root, tree->root_size));
  LB_TRY(dev_alloc(d, &nodes, tree->nodes_size ? tree->nodes_size : 64));
  LB_TRY(dev_alloc(d, &d->d_light_handles, tree->num_lights));
  LB_CHECK(cudaMemcpyAsync(root, tree->root_data, tree->root_size, cudaMemcpyHostToDevice, d->stream));
  if (tree->nodes_size)
    LB_CHECK(cudaMemcpyAsync(nodes, tree->nodes_data, tree->nodes_size, cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d->d_light_handles, tree->tri_handle_map, sizeof(uint2) * tree->num_lights, cudaMemcpyHostToDevice, d->stream));
  {
    // header word 2, bits 16..23: number of 8-child root sections (DeviceLightTreeRootHeader, device_utils.h:305-312)
    const uint32_t num_sections = (((const uint32_t*) tree->root_data)[2] >> 16) & 0xFFu;
    LB_REQUIRE(16 + 48 * (size_t) num_sections <= tree->root_size, LUMB200_ERROR_INVALID_API_ARGUMENT, "light tree root blob is truncated");
    // LIGHT_TREE_ROOT_MAX_NUM_SECTIONS (device_utils.h:45-47); k_shade stages the 8 children of every section in shared memory
    LB_REQUIRE(num_sections <= 16, LUMB200_ERROR_INVALID_API_ARGUMENT, "light tree root has %u sections, at most 16 are allowed", num_sections);
    LB_TRY(dev_alloc(d, &d->d_light_root_children, 16 * (size_t) (num_sections ? num_sections : 1)));
    lb_launch_unpack_light_root(root, d->d_light_root_children, num_sections, d->stream);
    d->light_root_sections = num_sections;
    LB_CHECK(cudaGetLastError());
  }
  LB_CHECK(cudaStreamSynchronize(d->stream));
  d->d_light_root     = root;
  d->d_light_nodes    = nodes;
  d->num_lights       = tree->num_lights;
  d->light_root_bytes = (uint32_t) tree->root_size;
  d->light_handles_host.assign(tree->tri_handle_map, tree->tri_handle_map + 2 * (size_t) tree->num_lights);
  return LUMB200_SUCCESS;
}

static Lumb200Result ensure_paths(Lumb200Device* d, uint32_t capacity) {
  if (capacity <= d->paths_capacity)
    return LUMB200_SUCCESS;
  LB_TRY(make_current(d));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  free_paths(d);
  LB_TRY(dev_alloc(d, &d->paths.org, capacity));
  LB_TRY(dev_alloc(d, &d->paths.dir, capacity));
  LB_TRY(dev_alloc(d, &d->paths.prim, capacity));
  LB_TRY(dev_alloc(d, &d->paths.record, capacity));
  LB_TRY(dev_alloc(d, &d->paths.pixel, capacity));
  LB_TRY(dev_alloc(d, &d->paths.state, capacity));
  LB_TRY(dev_alloc(d, &d->paths.medium, capacity));
  LB_TRY(dev_alloc(d, &d->paths.result, capacity));
  LB_TRY(dev_alloc(d, &d->paths.sample_id, capacity));
  LB_TRY(dev_alloc(d, &d->paths.nee, LB_NEE_SLOTS * (size_t) capacity));
  LB_TRY(dev_alloc(d, &d->paths.sq_org, LB_NEE_SLOTS * (size_t) capacity));
  LB_TRY(dev_alloc(d, &d->paths.sq_dir, LB_NEE_SLOTS * (size_t) capacity));
  LB_TRY(dev_alloc(d, &d->paths.sq_col, LB_NEE_SLOTS * (size_t) capacity));
  LB_TRY(dev_alloc(d, &d->paths.eq_org, capacity));
  LB_TRY(dev_alloc(d, &d->paths.eq_dir, capacity));
  LB_TRY(dev_alloc(d, &d->paths.eq_weight, capacity));
  LB_TRY(dev_alloc(d, &d->paths.eq_rec, capacity));
  LB_TRY(dev_alloc(d, &d->paths.eq_hits, capacity));
  LB_TRY(dev_alloc(d, &d->queue[0], capacity));
  LB_TRY(dev_alloc(d, &d->queue[1], capacity));
  LB_TRY(dev_alloc(d, &d->d_uv, capacity));
  d->paths.capacity = capacity;
  d->paths_capacity = capacity;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_update_settings(Lumb200Device* d, const Lumb200Settings* s) {
  LB_REQUIRE(d && s, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(s->width > 0 && s->height > 0 && s->width <= 16384 && s->height <= 16384, LUMB200_ERROR_INVALID_API_ARGUMENT,
             "resolution %ux%u is outside 1..16384 (PathID holds 14 bits per axis)", s->width, s->height);
  LB_REQUIRE((uint64_t) s->width * s->height < (1ull << 30), LUMB200_ERROR_INVALID_API_ARGUMENT, "more than 2^30 pixels per pass");
  LB_REQUIRE(s->max_ray_depth < 64, LUMB200_ERROR_INVALID_API_ARGUMENT, "max_ray_depth must be < 64");
  const bool resized = (s->width != d->settings.width) || (s->height != d->settings.height);
  d->settings        = *s;
  LB_TRY(make_current(d));
  LB_TRY(ensure_paths(d, s->width * s->height));
  if (resized || !d->planes) {
    LB_CHECK(cudaStreamSynchronize(d->stream));
    // the asynchronous result slots are sized for the old frame: wait for copies still in flight, then drop them
    if (d->copy_stream)
      LB_CHECK(cudaStreamSynchronize(d->copy_stream));
    for (int k = 0; k < 2; k++) {
      dev_free(d, d->d_result_async[k]);
      d->slot_pending[k] = false;
    }
    if (!d->planes_external)
      dev_free(d, d->planes);
    d->planes          = nullptr;
    d->planes_external = false;
    d->planes_floats   = 4 * (size_t) s->width * s->height;
    LB_TRY(dev_alloc(d, &d->planes, d->planes_floats));
    dev_free(d, d->d_result);
    LB_TRY(dev_alloc(d, &d->d_result, 3 * (size_t) s->width * s->height));
    dev_free(d, d->d_output);
    LB_TRY(dev_alloc(d, &d->d_output, (size_t) s->width * s->height));
    LB_CHECK(cudaMemsetAsync(d->planes, 0, sizeof(float) * d->planes_floats, d->stream));
  }
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_update_camera(Lumb200Device* d, const Lumb200Camera* c) {
  LB_REQUIRE(d && c, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  float q[4];
  euler_to_quat(c->rotation, q);  // device_struct_camera_convert, device_structs.c:41-89
  d->camera.px = c->pos[0], d->camera.py = c->pos[1], d->camera.pz = c->pos[2];
  d->camera.qx = q[0], d->camera.qy = q[1], d->camera.qz = q[2], d->camera.qw = q[3];
  d->camera.fov                  = c->fov;
  d->camera.aperture_size        = c->aperture_size;
  d->camera.object_distance      = c->object_distance;
  d->camera.camera_scale         = c->camera_scale;
  d->camera.rr_threshold         = c->russian_roulette_threshold;
  d->camera.aperture_shape       = c->aperture_shape;
  d->camera.aperture_blade_count = c->aperture_blade_count;
  return LUMB200_SUCCESS;
}

extern "C" void lumb200_sky_default(Lumb200Sky* s) {  // sky_get_default, sky.c:6-42 (mode: this library starts with the constant colour)
  if (!s)
    return;
  memset(s, 0, sizeof(*s));
  s->mode = 2;
  s->constant_color[0] = s->constant_color[1] = s->constant_color[2] = 1.0f;
  s->geometry_offset[1]     = 0.1f;
  s->altitude               = 0.5f;
  s->azimuth                = 3.141f;
  s->moon_altitude          = -0.5f;
  s->sun_strength           = 1.0f;
  s->base_density           = 1.0f;
  s->rayleigh_density       = 1.0f;
  s->mie_density            = 1.0f;
  s->ozone_density          = 1.0f;
  s->ground_visibility      = 60.0f;
  s->mie_diameter           = 2.0f;
  s->ozone_layer_thickness  = 15.0f;
  s->rayleigh_falloff       = 8.0f;
  s->mie_falloff            = 1.7f;
  s->multiscattering_factor = 1.0f;
  s->steps                  = 40;
  s->ozone_absorption       = 1;
  s->stars_count            = 10000;
  s->stars_intensity        = 1.0f;
  s->hdri_dim               = 2048;
  s->hdri_samples           = 32;
}

// device_struct_sky_convert, device_structs.c:132-170: positions of the sun and the moon in sky space (double arithmetic)
static void celestial_position(float azimuth, float altitude, double distance, const float* offset, float* out) {
  // the reference is C: cos(float) promotes to double there, while C++ would pick the float overload
  const double az = (double) azimuth, al = (double) altitude;
  double x = cos(az) * cos(al);
  double y = sin(al);
  double z = sin(az) * cos(al);
  const double scale = 1.0 / (sqrt(x * x + y * y + z * z));
  x *= scale * distance;
  y *= scale * distance;
  z *= scale * distance;
  y -= LB_SKY_EARTH_RADIUS;
  x -= offset[0];
  y -= offset[1];
  z -= offset[2];
  out[0] = (float) x, out[1] = (float) y, out[2] = (float) z;
}

// _sky_stars_generate, device_sky.c:484-546: the catalogue is a function of (seed, count) through glibc's rand(), binned into a
// 64 x 32 (azimuth, altitude) grid of 0.1 rad cells. rand() is process-global state, like in the reference.
static Lumb200Result generate_stars(Lumb200Device* d, uint32_t count, uint32_t seed) {
  const uint32_t cells = LB_STARS_GRID_X * LB_STARS_GRID_Y;
  std::vector<float> buffer(4 * (size_t) count);
  std::vector<uint32_t> counts(cells, 0);
  srand(seed);
  auto random_float = []() { return (float) (((double) rand()) / RAND_MAX); };
  for (uint32_t i = 0; i < count; i++) {
    const float altitude  = -LB_SKY_PI * 0.5f + LB_SKY_PI * (1.0f - sqrtf(random_float()));
    const float azimuth   = 2.0f * LB_SKY_PI * random_float();
    const float radius    = 0.0001f + 0.0004f * (1.0f - sqrtf(random_float()));
    const float intensity = 0.0001f + 0.0015f * (0.1f + 0.9f * (1.0f - sqrtf(random_float())));
    const uint32_t x      = (uint32_t) (azimuth * 10.0f);
    const uint32_t y      = (uint32_t) ((altitude + LB_SKY_PI * 0.5f) * 10.0f);
    LB_REQUIRE(x < LB_STARS_GRID_X && y < LB_STARS_GRID_Y, LUMB200_ERROR_API_EXCEPTION, "Star generation exception.");
    counts[x + y * LB_STARS_GRID_X]++;
    buffer[4 * (size_t) i + 0] = altitude, buffer[4 * (size_t) i + 1] = azimuth, buffer[4 * (size_t) i + 2] = radius, buffer[4 * (size_t) i + 3] = intensity;
  }
  d->stars_offsets_host.assign(cells + 1, 0);
  uint32_t offset = 0;
  for (uint32_t i = 0; i < cells; i++) {
    d->stars_offsets_host[i] = offset;
    offset += counts[i];
    counts[i] = 0;
  }
  d->stars_offsets_host[cells] = offset;
  d->stars_host.assign(4 * (size_t) count, 0.0f);
  for (uint32_t i = 0; i < count; i++) {
    const uint32_t x = (uint32_t) (buffer[4 * (size_t) i + 1] * 10.0f);
    const uint32_t y = (uint32_t) ((buffer[4 * (size_t) i + 0] + LB_SKY_PI * 0.5f) * 10.0f);
    const uint32_t p = x + y * LB_STARS_GRID_X;
    memcpy(&d->stars_host[4 * (size_t) (d->stars_offsets_host[p] + counts[p]++)], &buffer[4 * (size_t) i], sizeof(float) * 4);
  }
  dev_free(d, d->d_stars);
  d->d_stars = nullptr;
  if (!d->d_stars_offsets)
    LB_TRY(dev_alloc(d, &d->d_stars_offsets, cells + 1));
  LB_CHECK(cudaMemcpyAsync(d->d_stars_offsets, d->stars_offsets_host.data(), sizeof(uint32_t) * (cells + 1), cudaMemcpyHostToDevice, d->stream));
  if (count) {
    LB_TRY(dev_alloc(d, &d->d_stars, count));
    LB_CHECK(cudaMemcpyAsync(d->d_stars, d->stars_host.data(), sizeof(float) * 4 * (size_t) count, cudaMemcpyHostToDevice, d->stream));
  }
  LB_CHECK(cudaStreamSynchronize(d->stream));
  d->stars_count = count;
  d->stars_seed  = seed;
  return LUMB200_SUCCESS;
}

// the fields sky_check_for_dirty (sky.c:44-110) flags as SCENE_DIRTY_FLAG_INTEGRATION and the LUT kernels read
static bool sky_medium_differs(const Lumb200Sky& a, const Lumb200Sky& b) {
  return a.base_density != b.base_density || a.rayleigh_density != b.rayleigh_density || a.mie_density != b.mie_density
         || a.ozone_density != b.ozone_density || a.rayleigh_falloff != b.rayleigh_falloff || a.mie_falloff != b.mie_falloff
         || a.mie_diameter != b.mie_diameter || a.ground_visibility != b.ground_visibility || a.ozone_layer_thickness != b.ozone_layer_thickness
         || a.multiscattering_factor != b.multiscattering_factor || a.ozone_absorption != b.ozone_absorption;
}

// sky_lut_generate + device_sky_lut_update (device_sky.c:80-220): transmittance table, then the multiscattering table that samples it.
// Textures as texture_create / sky_lut_create configure them: float4, linear filter, clamp, normalised coordinates, no mips.
static Lumb200Result build_sky_luts(Lumb200Device* d) {
  const uint32_t dims[4][2] = {{LB_SKY_TM_TEX_WIDTH, LB_SKY_TM_TEX_HEIGHT}, {LB_SKY_TM_TEX_WIDTH, LB_SKY_TM_TEX_HEIGHT},
                               {LB_SKY_MS_TEX_SIZE, LB_SKY_MS_TEX_SIZE}, {LB_SKY_MS_TEX_SIZE, LB_SKY_MS_TEX_SIZE}};
  const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<float4>();
  for (int k = 0; k < 4; k++) {
    if (!d->d_sky_lut[k])
      LB_TRY(dev_alloc(d, &d->d_sky_lut[k], (size_t) dims[k][0] * dims[k][1]));
    if (!d->sky_arrays[k]) {
      LB_CHECK(cudaMallocArray(&d->sky_arrays[k], &fmt, dims[k][0], dims[k][1]));
      cudaResourceDesc rd;
      memset(&rd, 0, sizeof(rd));
      rd.resType         = cudaResourceTypeArray;
      rd.res.array.array = d->sky_arrays[k];
      cudaTextureDesc td;
      memset(&td, 0, sizeof(td));
      td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
      td.filterMode       = cudaFilterModeLinear;
      td.readMode         = cudaReadModeElementType;
      td.normalizedCoords = 1;
      LB_CHECK(cudaCreateTextureObject(&d->sky_tex[k], &rd, &td, nullptr));
    }
  }
  d->sky_dev.tm_low = d->sky_tex[0], d->sky_dev.tm_high = d->sky_tex[1], d->sky_dev.ms_low = d->sky_tex[2], d->sky_dev.ms_high = d->sky_tex[3];
  auto to_array = [&](int k) {
    return cudaMemcpy2DToArrayAsync(d->sky_arrays[k], 0, 0, d->d_sky_lut[k], dims[k][0] * sizeof(float4), dims[k][0] * sizeof(float4), dims[k][1],
                                    cudaMemcpyDeviceToDevice, d->stream);
  };
  lb_launch_sky_transmittance_lut(d->sky_dev, d->d_sky_lut[0], d->d_sky_lut[1], d->stream);
  LB_CHECK(to_array(0));
  LB_CHECK(to_array(1));
  lb_launch_sky_multiscattering_lut(d->sky_dev, d->d_sky_lut[2], d->d_sky_lut[3], d->stream);
  LB_CHECK(to_array(2));
  LB_CHECK(to_array(3));
  d->launches += 2;
  LB_CHECK(cudaGetLastError());
  LB_CHECK(cudaStreamSynchronize(d->stream));
  d->sky_lut_valid  = true;
  d->sky_lut_params = d->sky;
  return LUMB200_SUCCESS;
}

// sky_hdri_generate + _sky_hdri_compute (device_sky.c:279-375): table of hdri_dim^2 float4 baked from the camera position; texture as
// sky_hdri_generate configures it (point filter) over texture_create's defaults (wrap addressing, normalised coordinates).
static Lumb200Result build_sky_hdri(Lumb200Device* d) {
  const uint32_t dim = std::max(d->sky.hdri_dim, 1u), samples = std::max(d->sky.hdri_samples, 1u);
  LB_REQUIRE(d->d_bluenoise, LUMB200_ERROR_API_EXCEPTION, "the blue-noise mask must be loaded before the sky HDRI is baked");
  if (dim != d->hdri_dim || !d->d_hdri) {
    if (d->hdri_tex)
      cudaDestroyTextureObject(d->hdri_tex);
    if (d->hdri_array)
      cudaFreeArray(d->hdri_array);
    dev_free(d, d->d_hdri);
    d->hdri_tex = 0, d->hdri_array = nullptr, d->d_hdri = nullptr, d->hdri_dim = 0;
    LB_TRY(dev_alloc(d, &d->d_hdri, (size_t) dim * dim));
    const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<float4>();
    LB_CHECK(cudaMallocArray(&d->hdri_array, &fmt, dim, dim));
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType         = cudaResourceTypeArray;
    rd.res.array.array = d->hdri_array;
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
    td.filterMode       = cudaFilterModePoint;
    td.readMode         = cudaReadModeElementType;
    td.normalizedCoords = 1;
    LB_CHECK(cudaCreateTextureObject(&d->hdri_tex, &rd, &td, nullptr));
    d->hdri_dim = dim;
  }
  d->sky_dev.hdri = d->hdri_tex;
  d->hdri_origin[0] = d->camera.px, d->hdri_origin[1] = d->camera.py, d->hdri_origin[2] = d->camera.pz;
  lb_launch_sky_hdri(d->sky_dev, d->d_bluenoise, d->hdri_origin, dim, samples, d->d_hdri, d->stream);
  d->launches += 1;
  LB_CHECK(cudaGetLastError());
  LB_CHECK(cudaMemcpy2DToArrayAsync(d->hdri_array, 0, 0, d->d_hdri, dim * sizeof(float4), dim * sizeof(float4), dim, cudaMemcpyDeviceToDevice, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  d->hdri_valid = true;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_update_sky(Lumb200Device* d, const Lumb200Sky* s) {
  LB_REQUIRE(d && s, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(s->mode <= 2, LUMB200_ERROR_INVALID_API_ARGUMENT, "invalid sky mode %u", s->mode);
  if (s->mode != 2) {
    LB_REQUIRE(s->steps >= 1 && s->steps < 1024, LUMB200_ERROR_INVALID_API_ARGUMENT, "sky steps %u outside 1..1023", s->steps);
    LB_REQUIRE(s->stars_count <= (1u << 24), LUMB200_ERROR_INVALID_API_ARGUMENT, "%u stars", s->stars_count);
    LB_REQUIRE(s->mode != 1 || s->hdri_dim <= 8192, LUMB200_ERROR_INVALID_API_ARGUMENT, "sky HDRI dimension %u exceeds 8192", s->hdri_dim);
  }
  // sky_hdri_update (device_sky.c:249-277): any change of the sky invalidates the baked table
  if (memcmp(&d->sky, s, sizeof(*s)) != 0)
    d->hdri_valid = false;
  d->sky = *s;
  if (s->mode == 2)
    return LUMB200_SUCCESS;
  LB_TRY(make_current(d));
  LbSkyDev& S = d->sky_dev;
  S.mode = s->mode, S.steps = s->steps, S.ozone_absorption = s->ozone_absorption ? 1u : 0u, S.aerial_perspective = s->aerial_perspective ? 1u : 0u;
  S.moon_tex_offset = s->moon_tex_offset;
  memcpy(S.geometry_offset, s->geometry_offset, sizeof(float) * 3);
  S.sun_strength = s->sun_strength, S.base_density = s->base_density, S.stars_intensity = s->stars_intensity;
  S.rayleigh_density = s->rayleigh_density, S.mie_density = s->mie_density, S.ozone_density = s->ozone_density;
  S.rayleigh_falloff = s->rayleigh_falloff, S.mie_falloff = s->mie_falloff, S.mie_diameter = s->mie_diameter;
  S.ground_visibility = s->ground_visibility, S.ozone_layer_thickness = s->ozone_layer_thickness, S.multiscattering_factor = s->multiscattering_factor;
  celestial_position(s->azimuth, s->altitude, LB_SKY_SUN_DISTANCE, s->geometry_offset, S.sun_pos);
  celestial_position(s->moon_azimuth, s->moon_altitude, LB_SKY_MOON_DISTANCE, s->geometry_offset, S.moon_pos);
  if (d->stars_count != s->stars_count || d->stars_seed != s->stars_seed)
    LB_TRY(generate_stars(d, s->stars_count, s->stars_seed));
  S.stars = d->d_stars, S.stars_offsets = d->d_stars_offsets, S.has_stars = d->stars_count ? 1u : 0u;
  if (!d->sky_lut_valid || sky_medium_differs(d->sky, d->sky_lut_params))
    LB_TRY(build_sky_luts(d));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_load_moon_textures(Lumb200Device* d, const Lumb200Texture* albedo, const Lumb200Texture* normal) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_TRY(make_current(d));
  const Lumb200Texture* src[2] = {albedo, normal};
  LbTexture* dst[2]            = {&d->sky_dev.moon_albedo, &d->sky_dev.moon_normal};
  for (int k = 0; k < 2; k++) {
    if (d->moon_tex[k])
      cudaDestroyTextureObject(d->moon_tex[k]);
    if (d->moon_array[k])
      cudaFreeArray(d->moon_array[k]);
    d->moon_tex[k] = 0, d->moon_array[k] = nullptr;
    dst[k]->handle = 0, dst[k]->gamma = 1.0f, dst[k]->size = 0;
    const Lumb200Texture* t = src[k];
    if (!t || !t->data)
      continue;
    LB_REQUIRE(t->width > 0 && t->height > 0 && t->width <= 0xFFFF && t->height <= 0xFFFF, LUMB200_ERROR_INVALID_API_ARGUMENT, "moon texture extent");
    LB_REQUIRE(t->num_components == 1 || t->num_components == 2 || t->num_components == 4, LUMB200_ERROR_API_EXCEPTION, "moon texture components");
    LB_REQUIRE(t->type <= LUMB200_TEXTURE_U16 && t->wrap_mode_u <= LUMB200_WRAP_BORDER && t->wrap_mode_v <= LUMB200_WRAP_BORDER
                 && t->filter <= LUMB200_FILTER_LINEAR,
               LUMB200_ERROR_API_EXCEPTION, "moon texture format");
    const int bits         = (t->type == LUMB200_TEXTURE_U8) ? 8 : (t->type == LUMB200_TEXTURE_U16) ? 16 : 32;
    const size_t row_bytes = (size_t) t->width * t->num_components * (bits / 8);
    LB_REQUIRE(t->pitch >= row_bytes, LUMB200_ERROR_INVALID_API_ARGUMENT, "moon texture pitch");
    const cudaChannelFormatKind kind = (t->type == LUMB200_TEXTURE_FP32) ? cudaChannelFormatKindFloat : cudaChannelFormatKindUnsigned;
    const int nc                     = (int) t->num_components;
    const cudaChannelFormatDesc desc = cudaCreateChannelDesc(bits, nc >= 2 ? bits : 0, nc == 4 ? bits : 0, nc == 4 ? bits : 0, kind);
    LB_CHECK(cudaMallocArray(&d->moon_array[k], &desc, t->width, t->height));
    LB_CHECK(cudaMemcpy2DToArrayAsync(d->moon_array[k], 0, 0, t->data, t->pitch, row_bytes, t->height, cudaMemcpyHostToDevice, d->stream));
    LB_CHECK(cudaStreamSynchronize(d->stream));
    cudaResourceDesc res;
    memset(&res, 0, sizeof(res));
    res.resType         = cudaResourceTypeArray;
    res.res.array.array = d->moon_array[k];
    cudaTextureDesc tex;
    memset(&tex, 0, sizeof(tex));
    const cudaTextureAddressMode modes[4] = {cudaAddressModeWrap, cudaAddressModeClamp, cudaAddressModeMirror, cudaAddressModeBorder};
    tex.addressMode[0] = modes[t->wrap_mode_u], tex.addressMode[1] = modes[t->wrap_mode_v], tex.addressMode[2] = cudaAddressModeClamp;
    tex.filterMode       = (t->filter == LUMB200_FILTER_LINEAR) ? cudaFilterModeLinear : cudaFilterModePoint;
    tex.readMode         = (t->type == LUMB200_TEXTURE_FP32) ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
    tex.normalizedCoords = 1;
    LB_CHECK(cudaCreateTextureObject(&d->moon_tex[k], &res, &tex, nullptr));
    dst[k]->handle = d->moon_tex[k], dst[k]->gamma = t->gamma, dst[k]->size = t->width | (t->height << 16);
    d->device_bytes += row_bytes * t->height;
  }
  d->hdri_valid = false;  // the table does not hold the moon (celestials off), but keep the rule simple: new inputs, new bake
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_build_sky_hdri(Lumb200Device* d) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(d->sky.mode == 1, LUMB200_ERROR_API_EXCEPTION, "the sky HDRI is baked under sky mode 1 (LUMINARY_SKY_MODE_HDRI) only");
  LB_TRY(make_current(d));
  return build_sky_hdri(d);
}

extern "C" Lumb200Result lumb200_device_get_sky_hdri(Lumb200Device* d, float* color, uint32_t capacity_texels, uint32_t* dim, float* origin) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(d->sky.mode == 1 && d->hdri_valid, LUMB200_ERROR_API_EXCEPTION, "no baked sky HDRI (sky mode 1 after build_sky_hdri / start_render)");
  if (dim)
    *dim = d->hdri_dim;
  if (origin)
    memcpy(origin, d->hdri_origin, sizeof(float) * 3);
  if (color) {
    LB_REQUIRE((uint64_t) capacity_texels >= (uint64_t) d->hdri_dim * d->hdri_dim, LUMB200_ERROR_INVALID_API_ARGUMENT, "capacity %u < %u^2 texels",
               capacity_texels, d->hdri_dim);
    LB_TRY(make_current(d));
    LB_CHECK(cudaMemcpyAsync(color, d->d_hdri, sizeof(float4) * (size_t) d->hdri_dim * d->hdri_dim, cudaMemcpyDeviceToHost, d->stream));
    LB_CHECK(cudaStreamSynchronize(d->stream));
  }
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_get_sky_lut(Lumb200Device* d, float* tm_low, float* tm_high, float* ms_low, float* ms_high) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(d->sky.mode != 2 && d->sky_lut_valid, LUMB200_ERROR_API_EXCEPTION, "the sky LUTs exist only under the procedural sky (modes 0 and 1)");
  LB_TRY(make_current(d));
  float* dst[4]         = {tm_low, tm_high, ms_low, ms_high};
  const size_t texels[4] = {(size_t) LB_SKY_TM_TEX_WIDTH * LB_SKY_TM_TEX_HEIGHT, (size_t) LB_SKY_TM_TEX_WIDTH * LB_SKY_TM_TEX_HEIGHT,
                            (size_t) LB_SKY_MS_TEX_SIZE * LB_SKY_MS_TEX_SIZE, (size_t) LB_SKY_MS_TEX_SIZE * LB_SKY_MS_TEX_SIZE};
  for (int k = 0; k < 4; k++)
    if (dst[k])
      LB_CHECK(cudaMemcpyAsync(dst[k], d->d_sky_lut[k], sizeof(float4) * texels[k], cudaMemcpyDeviceToHost, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_get_sky_info(Lumb200Device* d, float* sun_pos, float* moon_pos, float* stars, uint32_t capacity,
                                                      uint32_t* stars_offsets, uint32_t* stars_count) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(d->sky.mode != 2, LUMB200_ERROR_API_EXCEPTION, "the sky info exists only under the procedural sky (modes 0 and 1)");
  if (sun_pos)
    memcpy(sun_pos, d->sky_dev.sun_pos, sizeof(float) * 3);
  if (moon_pos)
    memcpy(moon_pos, d->sky_dev.moon_pos, sizeof(float) * 3);
  const uint32_t n = d->stars_count == 0xFFFFFFFFu ? 0u : d->stars_count;
  if (stars && n)
    memcpy(stars, d->stars_host.data(), sizeof(float) * 4 * (size_t) std::min(n, capacity));
  if (stars_offsets && !d->stars_offsets_host.empty())
    memcpy(stars_offsets, d->stars_offsets_host.data(), sizeof(uint32_t) * d->stars_offsets_host.size());
  if (stars_count)
    *stars_count = n;
  return LUMB200_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// acceleration structures
// ---------------------------------------------------------------------------------------------
__global__ void k_gather_light_tris(const float4* __restrict__ world, const uint32_t* __restrict__ light_prims, uint32_t n,
                                    float4* __restrict__ out) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n)
    return;
  const uint32_t p = light_prims[l];
  float4 v0        = world[3 * (size_t) p + 0];
  v0.w             = __uint_as_float(l);
  out[3 * (size_t) l + 0] = v0;
  out[3 * (size_t) l + 1] = world[3 * (size_t) p + 1];
  out[3 * (size_t) l + 2] = world[3 * (size_t) p + 2];
}

extern "C" Lumb200Result lumb200_device_build_accel(Lumb200Device* d) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_TRY(make_current(d));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  free_scene_tables(d);

  const uint32_t num_meshes    = (uint32_t) d->meshes.size();
  const uint32_t num_instances = (uint32_t) d->instances.size();

  std::vector<float4*> mv(num_meshes ? num_meshes : 1, nullptr);
  std::vector<uint4*> mt(num_meshes ? num_meshes : 1, nullptr);
  for (uint32_t m = 0; m < num_meshes; m++) {
    mv[m] = d->meshes[m].vertices;
    mt[m] = d->meshes[m].textris;
  }
  std::vector<uint32_t> inst_mesh(num_instances ? num_instances : 1, 0);
  std::vector<LbTransform> inst_x(num_instances ? num_instances : 1);
  std::vector<uint32_t> inst_off(num_instances + 1, 0);
  uint64_t total = 0;
  for (uint32_t i = 0; i < num_instances; i++) {
    const Lumb200Instance& in = d->instances[i];
    inst_mesh[i]              = in.mesh_id;
    inst_off[i]               = (uint32_t) total;
    if (in.active)
      total += d->meshes[in.mesh_id].num_tris;
    const LbTransform t = pack_transform(in);
    inst_x[i] = t;
  }
  LB_REQUIRE(total < 0x7FFFFFFFull, LUMB200_ERROR_INVALID_API_ARGUMENT, "scene has too many triangles (%llu)", (unsigned long long) total);
  inst_off[num_instances] = (uint32_t) total;
  d->num_prims            = (uint32_t) total;

  // per flattened primitive material id (sort key + shadow response lookup)
  std::vector<uint16_t> prim_mat(total ? total : 1, 0);
  for (uint32_t i = 0; i < num_instances; i++) {
    if (!d->instances[i].active)
      continue;
    const MeshDev& m = d->meshes[d->instances[i].mesh_id];
    memcpy(prim_mat.data() + inst_off[i], m.host_material.data(), sizeof(uint16_t) * m.num_tris);
  }

  LB_TRY(dev_alloc(d, &d->d_mesh_vertices, mv.size()));
  LB_TRY(dev_alloc(d, &d->d_mesh_textris, mt.size()));
  LB_TRY(dev_alloc(d, &d->d_instance_mesh, inst_mesh.size()));
  LB_TRY(dev_alloc(d, &d->d_instance_xform, inst_x.size()));
  LB_TRY(dev_alloc(d, &d->d_instance_offset, inst_off.size()));
  LB_TRY(dev_alloc(d, &d->d_prim_handle, (size_t) total));
  LB_TRY(dev_alloc(d, &d->d_prim_material, prim_mat.size()));
  LB_TRY(dev_alloc(d, &d->d_world_tris, 3 * (size_t) total));

  LB_CHECK(cudaMemcpyAsync(d->d_mesh_vertices, mv.data(), sizeof(float4*) * mv.size(), cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d->d_mesh_textris, mt.data(), sizeof(uint4*) * mt.size(), cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d->d_instance_mesh, inst_mesh.data(), sizeof(uint32_t) * inst_mesh.size(), cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d->d_instance_xform, inst_x.data(), sizeof(LbTransform) * inst_x.size(), cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d->d_instance_offset, inst_off.data(), sizeof(uint32_t) * inst_off.size(), cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d->d_prim_material, prim_mat.data(), sizeof(uint16_t) * prim_mat.size(), cudaMemcpyHostToDevice, d->stream));

  LbSceneTables tab;
  tab.mesh_vertices        = (const float4* const*) d->d_mesh_vertices;
  tab.mesh_textris         = (const uint4* const*) d->d_mesh_textris;
  tab.instance_mesh        = d->d_instance_mesh;
  tab.instance_transform   = d->d_instance_xform;
  tab.instance_prim_offset = d->d_instance_offset;
  tab.prim_handle          = d->d_prim_handle;
  tab.materials            = d->d_materials;
  tab.num_instances        = num_instances;
  tab.num_prims            = d->num_prims;
  tab.num_materials        = d->num_materials;

  LB_TRY(lb_flatten_instances(tab, d->d_world_tris, d->d_prim_handle, d->stream));
  d->launches++;

  float ms = 0.0f;
  LB_TRY(lb_bvh8_build(d->d_world_tris, d->num_prims, &d->bvh, d->stream, &ms));
  d->accel_seconds = ms * 1e-3;
  d->device_bytes += d->bvh.bytes;
  // the persistent traversal loop pushes at most two stack entries per level (trace_loop.cuh): refuse a tree it cannot walk
  LB_REQUIRE(d->bvh.depth <= LB_TRAVERSAL_MAX_LEVELS, LUMB200_ERROR_API_EXCEPTION,
             "scene BVH is %u levels deep; the traversal stack holds %u", d->bvh.depth, (unsigned) LB_TRAVERSAL_MAX_LEVELS);

  // emitter-only BVH for BSDF-sampled NEE (replaces optix_bvh_light_build, device/optix_bvh.c:382-478)
  if (d->num_lights) {
    std::vector<uint32_t> light_prims(d->num_lights);
    for (uint32_t l = 0; l < d->num_lights; l++) {
      const uint32_t inst = d->light_handles_host[2 * l + 0];
      const uint32_t tri  = d->light_handles_host[2 * l + 1];
      LB_REQUIRE(inst < num_instances && d->instances[inst].active && tri < d->meshes[d->instances[inst].mesh_id].num_tris,
                 LUMB200_ERROR_INVALID_API_ARGUMENT, "light %u references an invalid triangle handle (%u, %u)", l, inst, tri);
      light_prims[l] = inst_off[inst] + tri;
    }
    LB_TRY(dev_alloc(d, &d->d_light_prims, d->num_lights));
    LB_TRY(dev_alloc(d, &d->d_light_world, 3 * (size_t) d->num_lights));
    LB_CHECK(cudaMemcpyAsync(d->d_light_prims, light_prims.data(), sizeof(uint32_t) * d->num_lights, cudaMemcpyHostToDevice, d->stream));
    k_gather_light_tris<<<(d->num_lights + 255) / 256, 256, 0, d->stream>>>(d->d_world_tris, d->d_light_prims, d->num_lights, d->d_light_world);
    LB_CHECK(cudaGetLastError());
    float lms = 0.0f;
    LB_TRY(lb_bvh8_build(d->d_light_world, d->num_lights, &d->light_bvh, d->stream, &lms));
    d->accel_seconds += lms * 1e-3;
    d->device_bytes += d->light_bvh.bytes;
    LB_REQUIRE(d->light_bvh.depth <= LB_TRAVERSAL_MAX_LEVELS, LUMB200_ERROR_API_EXCEPTION, "emitter BVH is %u levels deep; the traversal stack holds %u",
               d->light_bvh.depth, (unsigned) LB_TRAVERSAL_MAX_LEVELS);
  }

  // keep BVH nodes resident in L2: persisting access-policy window over the node array
  {
    cudaDeviceProp prop;
    // Measured on B200 (profiles/r2_l2_window_ab.md): with the window the pass is SLOWER (atrium-1M 9.17 -> 8.70 ms without it,
    // terrain-10M 9.59 -> 9.35 ms): the persisting carve-out takes L2 capacity from the path state k_shade streams, and the node
    // array stays resident on its own (it is the hottest data of the pass). The window is therefore opt-in: LUMB200_L2_WINDOW=1.
    if (getenv("LUMB200_L2_WINDOW") && cudaGetDeviceProperties(&prop, d->cuda_index) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 &&
        d->bvh.nodes) {
      const size_t node_bytes = sizeof(Bvh8Node) * (size_t) d->bvh.num_nodes;
      const size_t carve      = (size_t) prop.persistingL2CacheMaxSize;
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
      cudaStreamAttrValue attr;
      memset(&attr, 0, sizeof(attr));
      const size_t window                     = node_bytes < (size_t) prop.accessPolicyMaxWindowSize ? node_bytes : (size_t) prop.accessPolicyMaxWindowSize;
      attr.accessPolicyWindow.base_ptr        = (void*) d->bvh.nodes;
      attr.accessPolicyWindow.num_bytes       = window;
      attr.accessPolicyWindow.hitRatio        = (window <= carve) ? 1.0f : (float) carve / (float) window;
      attr.accessPolicyWindow.hitProp         = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp        = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(d->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();
    }
  }

  LB_CHECK(cudaStreamSynchronize(d->stream));
  d->accel_dirty         = false;
  d->light_records_dirty = true;
  return LUMB200_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// BSDF LUTs
// ---------------------------------------------------------------------------------------------
extern "C" Lumb200Result lumb200_device_build_bsdf_lut(Lumb200Device* d) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(d->d_bluenoise, LUMB200_ERROR_MISSING_DATA, "blue-noise mask not loaded");
  LB_TRY(make_current(d));
  LB_TRY(lb_lut_generate(&d->luts, d->d_bluenoise, d->stream));
  d->launches += 3;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_get_bsdf_lut(Lumb200Device* d, uint16_t* conductor, uint16_t* glossy, uint16_t* dielectric,
                                                     uint16_t* dielectric_inv) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(d->luts.valid, LUMB200_ERROR_API_EXCEPTION, "BSDF LUTs have not been built");
  LB_TRY(make_current(d));
  return lb_lut_download(&d->luts, conductor, glossy, dielectric, dielectric_inv, d->stream);
}

extern "C" Lumb200Result lumb200_device_set_bsdf_lut(Lumb200Device* d, const uint16_t* conductor, const uint16_t* glossy,
                                                     const uint16_t* dielectric, const uint16_t* dielectric_inv) {
  LB_REQUIRE(d && conductor && glossy && dielectric && dielectric_inv, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_TRY(make_current(d));
  return lb_lut_upload(&d->luts, conductor, glossy, dielectric, dielectric_inv, d->stream);
}

// ---------------------------------------------------------------------------------------------
// rendering
// ---------------------------------------------------------------------------------------------
static LbFrame make_frame(const Lumb200Device* d) {
  LbFrame F;
  F.width     = d->settings.width;
  F.height    = d->settings.height;
  F.max_depth = d->settings.max_ray_depth;
  F.sky_mode  = d->sky.mode;
  F.sky_r     = d->sky.constant_color[0];
  F.sky_g     = d->sky.constant_color[1];
  F.sky_b     = d->sky.constant_color[2];
  return F;
}

static Bvh8 make_bvh(const LbBvhBuffers& b) {
  Bvh8 r;
  r.nodes     = b.nodes;
  r.tris      = b.tris;
  r.num_nodes = b.num_nodes;
  r.num_tris  = b.num_tris;
  return r;
}

static Lumb200Result check_ready(Lumb200Device* d, bool need_shading) {
  LB_REQUIRE(d->settings.width && d->settings.height, LUMB200_ERROR_API_EXCEPTION, "settings have not been set");
  LB_REQUIRE(d->d_bluenoise, LUMB200_ERROR_MISSING_DATA, "blue-noise mask not loaded");
  LB_REQUIRE(!d->accel_dirty && d->bvh.nodes, LUMB200_ERROR_API_EXCEPTION, "acceleration structure is out of date: call lumb200_device_build_accel");
  if (need_shading) {
    LB_REQUIRE(d->luts.valid, LUMB200_ERROR_API_EXCEPTION, "BSDF LUTs have not been built");
    LB_REQUIRE(d->num_materials > 0 || d->num_prims == 0, LUMB200_ERROR_API_EXCEPTION, "no materials uploaded");
  }
  // every id a kernel dereferences must be in range: triangle -> material (the any-hit kernels read it too), material -> texture
  for (const Lumb200Instance& in : d->instances) {
    if (!in.active || d->meshes[in.mesh_id].num_tris == 0)
      continue;
    if (d->num_materials == 0 && !need_shading)
      break;  // geometry-only use (closest-hit hooks): no kernel looks at materials
    LB_REQUIRE(d->meshes[in.mesh_id].max_material < d->num_materials, LUMB200_ERROR_INVALID_API_ARGUMENT, "mesh %u references material %u but only %u materials are uploaded", in.mesh_id,
               (unsigned) d->meshes[in.mesh_id].max_material, d->num_materials);
  }
  const MaterialPacked* mp = (const MaterialPacked*) d->materials_packed.data();
  for (uint32_t i = 0; i < d->num_materials; i++) {
    const uint16_t ids[5] = {mp[i].albedo_tex, mp[i].luminance_tex, mp[i].roughness_tex, mp[i].normal_tex, mp[i].metallic_tex};
    for (uint16_t id : ids)
      LB_REQUIRE(id == 0xFFFF || id < d->textures.size(), LUMB200_ERROR_INVALID_API_ARGUMENT,
                 "material %u references texture %u but only %zu textures are uploaded", i, (unsigned) id, d->textures.size());
  }
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_start_render(Lumb200Device* d) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_TRY(check_ready(d, true));
  LB_TRY(make_current(d));
  if (d->sky.mode == 1 && !d->hdri_valid)  // SCENE_DIRTY_FLAG_HDRI of a changed sky (device_manager.c:351-365)
    LB_TRY(build_sky_hdri(d));
  LB_CHECK(cudaMemsetAsync(d->planes, 0, sizeof(float) * d->planes_floats, d->stream));
  LB_CHECK(cudaMemsetAsync(d->counters, 0, sizeof(LbCounters), d->stream));
  // adaptive_sampler_setup + device_adaptive_sampler_reset (device_adaptive_sampler.c:29-56, 320-328)
  d->as_active = d->as_params.enable != 0;
  d->as_stage  = 0;
  d->as_paths  = 0;
  memset(d->as_executions, 0, sizeof(d->as_executions));
  if (d->as_active) {
    const uint32_t bw = (d->settings.width + 3u) >> 2, bh = (d->settings.height + 3u) >> 2;
    if (bw != d->as_bw || bh != d->as_bh || !d->d_as_words) {
      dev_free(d, d->d_as_words);
      dev_free(d, d->d_as_prefix);
      dev_free(d, d->d_as_block_var);
      LB_TRY(dev_alloc(d, &d->d_as_words, (size_t) bw * bh));
      LB_TRY(dev_alloc(d, &d->d_as_prefix, (size_t) bw * bh));
      LB_TRY(dev_alloc(d, &d->d_as_block_var, (size_t) bw * bh));
      if (!d->d_as_total)
        LB_TRY(dev_alloc(d, &d->d_as_total, 1));
      if (!d->d_as_var_sum)
        LB_TRY(dev_alloc(d, &d->d_as_var_sum, 1));
      d->as_bw = bw, d->as_bh = bh;
    }
    LB_CHECK(cudaMemsetAsync(d->d_as_words, 0, sizeof(uint32_t) * (size_t) bw * bh, d->stream));
    d->as_total_tasks = (bw * bh) << 4;  // adaptive_sampler_setup: upper_bound_tasks_per_sample
  }
  LB_CHECK(cudaStreamSynchronize(d->stream));
  d->render_seconds = 0.0;
  d->samples_done   = 0;
  d->events_pending = 0;
  d->launches       = 0;
  return LUMB200_SUCCESS;
}

static Lumb200Result collect_events(Lumb200Device* d) {
  for (size_t i = 0; i < d->events_pending; i++) {
    float ms = 0.0f;
    LB_CHECK(cudaEventSynchronize(d->ev_end[i]));
    LB_CHECK(cudaEventElapsedTime(&ms, d->ev_start[i], d->ev_end[i]));
    d->render_seconds += ms * 1e-3;
  }
  d->events_pending = 0;
  return LUMB200_SUCCESS;
}

static void prof_collect(Lumb200Device* d) {
  for (size_t k = 0; k < d->prof_used; k++) {
    float ms = 0.0f;
    if (cudaEventSynchronize(d->prof_events[2 * k + 1]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, d->prof_events[2 * k], d->prof_events[2 * k + 1]) == cudaSuccess) {
      d->profile.milliseconds[d->prof_class[k]] += ms;
      d->profile.launches[d->prof_class[k]]++;
    }
  }
  d->prof_used = 0;
}

struct ProfScope {
  Lumb200Device* d;
  size_t slot;
  bool on;
  ProfScope(Lumb200Device* dev, int cls) : d(dev), slot(0), on(dev->profiling) {
    if (!on)
      return;
    if (d->prof_used >= 4096)
      prof_collect(d);
    slot = d->prof_used++;
    while (d->prof_events.size() < 2 * (slot + 1)) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      d->prof_events.push_back(e);
    }
    if (d->prof_class.size() <= slot)
      d->prof_class.resize(slot + 1);
    d->prof_class[slot] = cls;
    cudaEventRecord(d->prof_events[2 * slot], d->stream);
  }
  ~ProfScope() {
    if (on)
      cudaEventRecord(d->prof_events[2 * slot + 1], d->stream);
  }
};

// One sample pass = the reference's per-tile action queue (device_renderer.c:53-134, 434-463):
//   tasks_create; for depth 0..D { trace; classify+sort; shade (geometry + sky); shadow } ; collect results.
static void fill_scene_params(const Lumb200Device* d, LbShadeParams& sp) {
  sp.prim_handle         = d->d_prim_handle;
  sp.mesh_vertices       = (const float4* const*) d->d_mesh_vertices;
  sp.mesh_textris        = (const uint4* const*) d->d_mesh_textris;
  sp.instance_mesh       = d->d_instance_mesh;
  sp.instance_xform      = d->d_instance_xform;
  sp.instance_offset     = d->d_instance_offset;
  sp.materials           = d->d_materials;
  sp.prim_material       = d->d_prim_material;
  sp.textures            = d->d_textures;
  sp.num_textures        = (uint32_t) d->textures.size();
  sp.textured            = d->any_material_tex ? 1u : 0u;
  sp.light_root          = (const uint4*) d->d_light_root;
  sp.light_root_children = d->d_light_root_children;
  sp.light_nodes         = (const uint4*) d->d_light_nodes;
  sp.light_handles       = d->d_light_handles;
  sp.light_prims         = d->d_light_prims;
  sp.light_records       = d->d_light_records;
  sp.num_lights          = d->num_lights;
}

// (re)builds the per-light records when the emitters, instances or materials changed since the last render
static Lumb200Result ensure_light_records(Lumb200Device* d) {
  if (!d->light_records_dirty)
    return LUMB200_SUCCESS;
  if (d->num_lights && d->d_light_prims && d->d_materials) {
    dev_free(d, d->d_light_records);
    LB_TRY(dev_alloc(d, &d->d_light_records, 4 * (size_t) d->num_lights));
    LbShadeParams sp;
    memset(&sp, 0, sizeof(sp));
    fill_scene_params(d, sp);
    lb_launch_build_light_records(sp, d->d_light_records, d->stream);
    LB_CHECK(cudaGetLastError());
    d->launches++;
  }
  d->light_records_dirty = false;
  return LUMB200_SUCCESS;
}

// everything the textured any-hit variants of the traversal kernels read (texture.cuh)
static LbTexScene make_tex_scene(const Lumb200Device* d) {
  LbTexScene T;
  T.textures      = d->d_textures;
  T.num_textures  = (uint32_t) d->textures.size();
  T.materials     = d->d_materials;
  T.prim_handle   = d->d_prim_handle;
  T.instance_mesh = d->d_instance_mesh;
  T.mesh_textris  = (const uint4* const*) d->d_mesh_textris;
  T.prim_material = d->d_prim_material;
  T.shadow_tab    = d->d_shadow_tab;
  return T;
}

static LbAdaptive make_adaptive(const Lumb200Device* d) {
  LbAdaptive A;
  A.words = d->d_as_words;
  A.bw    = d->as_bw;
  for (int k = 0; k <= LB_ADAPTIVE_STAGES; k++)
    A.executions[k] = d->as_executions[k];
  return A;
}

// AdaptiveChunk != nullptr: the wavefront holds tasks [begin, begin + count) of one execution of an adaptive stage >= 1
// (tasks_create_adaptive_sampling instead of tasks_create); paths carry their own sample ids and results are added atomically.
struct AdaptiveChunk {
  uint32_t begin, count;
};

// shading parameters of one pass (everything but the per-iteration queue pointers and depth)
static LbShadeParams make_shade_params(const Lumb200Device* d, const LbFrame& F, uint32_t sample_id, bool adaptive) {
  LbShadeParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.paths     = d->paths;
  sp.frame     = F;
  sp.camera    = d->camera;
  sp.bluenoise = d->d_bluenoise;
  sp.rng_table = d->d_rng_table;
  sp.sample_id = sample_id;
  sp.counters  = d->counters;
  fill_scene_params(d, sp);
  sp.luts      = d->luts.tex;
  sp.sky       = d->sky_dev;
  sp.light_bvh = make_bvh(d->light_bvh);
  sp.adaptive  = adaptive ? 1u : 0u;
  return sp;
}

// The surface stages of one wavefront iteration: material sort of queue[cur] into queue[cur ^ 1], shading (survivors are
// appended to queue[cur], NEE segments to the shadow queue), shadow rays. Shared by render_pass and the per-vertex parity hook.
static void surface_stages(Lumb200Device* d, LbShadeParams& sp, const Bvh8& bvh, int cur, uint32_t rng_depth, bool is_last, bool count,
                           const LbTexScene* tex) {
  cudaStream_t s = d->stream;
  if (d->sky.mode != 2 && d->sky_dev.aerial_perspective) {
    // render_inscattering (device_manager.c:475): aerial perspective between the trace and the sort (device_renderer.c:84-88)
    ProfScope ps(d, LUMB200_KERNEL_SHADE);
    sp.queue_in  = d->queue[cur];
    sp.rng_depth = rng_depth;
    lb_launch_sky_inscattering(sp, d->stream_grid, s);
    d->launches += 1;
  }
  // unsorted mode (sort_by_material = 0): hits / misses only, every hit is shaded by the GENERIC kernel
  const bool sorted           = d->settings.sort_by_material != 0;
  const LbSortClasses unsorted = {{0, LB_SORT_KEY_SKY, LB_SORT_KEY_SKY, LB_SORT_KEY_SKY}};
  {
    ProfScope ps(d, LUMB200_KERNEL_SORT);
    lb_launch_sort(d->paths, d->queue[cur], d->queue[cur ^ 1], d->counters, d->d_prim_material, sorted ? d->d_sort_rank : nullptr,
                   sorted ? d->sort_classes : unsorted, d->sort_bins, d->stream_grid, s);
  }
  sp.queue_in  = d->queue[cur ^ 1];
  sp.queue_out = d->queue[cur];
  sp.rng_depth = rng_depth;
  sp.is_last   = is_last ? 1u : 0u;
  sp.count     = count ? 1u : 0u;
  for (int c = 0; c < LB_NUM_CLASSES; c++)
    sp.class_materials[c] = sorted ? d->class_materials[c] : (c == LB_CLASS_GENERIC ? 1u : 0u);
  {
    ProfScope ps(d, LUMB200_KERNEL_SHADE);
    d->launches += lb_launch_shade(sp, d->shade_grid, s, d->shade_overlap ? d->aux_stream : nullptr, d->ev_fork, d->ev_join);
  }
  if (d->num_lights) {
    // BSDF-sampled NEE: enumerate the emitters along the queued directions, then evaluate the samples into slot-1 shadow segments
    ProfScope ps(d, LUMB200_KERNEL_TRACE_ENUM);
    lb_launch_trace_enum(make_bvh(d->light_bvh), d->paths, d->counters, d->d_light_prims, make_tex_scene(d), d->any_albedo_tex, d->trace_grid, s);
    lb_launch_enum_finish(sp, d->stream_grid, s);
    d->launches += 2;
  }
  {
    ProfScope ps(d, LUMB200_KERNEL_TRACE_SHADOW);
    lb_launch_trace_shadow(bvh, d->paths, d->counters, d->d_prim_material, d->d_shadow_tab, d->trace_grid, s, count, tex);
  }
  d->launches += 4;  // sort: count, scan, scatter; shadow
}

static Lumb200Result render_pass(Lumb200Device* d, uint32_t sample_id, bool count = false, bool accumulate = true,
                                 const AdaptiveChunk* chunk = nullptr) {
  const LbFrame F = make_frame(d);
  const Bvh8 bvh  = make_bvh(d->bvh);
  cudaStream_t s  = d->stream;
  const LbTexScene tex_scene = make_tex_scene(d);
  const LbTexScene* tex      = d->any_albedo_tex ? &tex_scene : nullptr;

  if (chunk) {
    ProfScope ps(d, LUMB200_KERNEL_RAYGEN);
    lb_launch_raygen_adaptive(d->paths, F, d->camera, d->d_bluenoise, make_adaptive(d), d->as_stage, d->d_as_prefix, d->as_bw * d->as_bh,
                              chunk->begin, chunk->count, d->queue[0], d->counters, d->stream_grid, s);
    d->launches += 2;
  }
  else {
    {
      ProfScope ps(d, LUMB200_KERNEL_RAYGEN);
      lb_launch_raygen(d->paths, F, d->camera, d->d_bluenoise, sample_id, d->queue[0], d->counters, d->stream_grid, s, d->tile_order);
    }
    lb_launch_rng_table(d->d_rng_table, sample_id, F.max_depth + 1, s);
    d->launches += 2;
  }

  LbShadeParams sp = make_shade_params(d, F, sample_id, chunk != nullptr);

  const int cur = 0;
  if (d->shading_mode != 0) {
    // _device_renderer_build_debug_kernel_queue (device_renderer.c:136-182): raytrace, [inscattering], sort, the *_process_tasks_debug kernels
    {
      ProfScope ps(d, LUMB200_KERNEL_TRACE_CLOSEST);
      lb_launch_trace_closest(bvh, d->paths, d->queue[cur], d->counters, nullptr, d->trace_grid, s, count, tex);
    }
    if (d->sky.mode != 2 && d->sky_dev.aerial_perspective) {
      ProfScope ps(d, LUMB200_KERNEL_SHADE);
      sp.queue_in  = d->queue[cur];
      sp.rng_depth = 0;
      lb_launch_sky_inscattering(sp, d->stream_grid, s);
      d->launches += 1;
    }
    const LbSortClasses hit_miss = {{0, LB_SORT_KEY_SKY, LB_SORT_KEY_SKY, LB_SORT_KEY_SKY}};
    {
      ProfScope ps(d, LUMB200_KERNEL_SORT);
      lb_launch_sort(d->paths, d->queue[cur], d->queue[cur ^ 1], d->counters, d->d_prim_material, nullptr, hit_miss, d->sort_bins, d->stream_grid, s);
    }
    sp.queue_in  = d->queue[cur ^ 1];
    sp.queue_out = d->queue[cur];
    sp.rng_depth = 0;
    {
      ProfScope ps(d, LUMB200_KERNEL_SHADE);
      d->launches += lb_launch_shade_debug(sp, d->shading_mode, d->shade_grid, s);
    }
    lb_launch_next_bounce(d->counters, s);
    d->launches += 5;  // closest; sort: count, scan, scatter; next_bounce
  }
  else
  for (uint32_t depth = 0; depth <= F.max_depth; depth++) {
    // device.state.depth as seen by the kernels: the reference skips the UPDATE_DEPTH action when
    // depth + 1 == max_depth (device_renderer.c:126-130), so the last iteration re-uses the previous value.
    uint32_t rng_depth = depth;
    if (depth == F.max_depth && depth > 0)
      rng_depth = depth - 1;

    {
      ProfScope ps(d, LUMB200_KERNEL_TRACE_CLOSEST);
      lb_launch_trace_closest(bvh, d->paths, d->queue[cur], d->counters, nullptr, d->trace_grid, s, count, tex);
    }
    surface_stages(d, sp, bvh, cur, rng_depth, depth == F.max_depth, count, tex);
    lb_launch_next_bounce(d->counters, s);
    d->launches += 2;
    // survivors were appended to queue[cur]; it is the active queue of the next bounce
  }

  if (accumulate) {
    ProfScope ps(d, LUMB200_KERNEL_ACCUMULATE);
    if (chunk)
      lb_launch_accumulate_adaptive(d->paths, chunk->count, F.width * F.height, d->planes, d->counters, d->stream_grid, s);
    else
      lb_launch_accumulate(d->paths, F, d->planes, d->counters, d->stream_grid, s);
    d->launches++;
  }
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_set_shading_mode(Lumb200Device* d, uint32_t shading_mode) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(shading_mode <= 5, LUMB200_ERROR_INVALID_API_ARGUMENT, "Invalid shading mode %u.", shading_mode);
  d->shading_mode = shading_mode;
  return LUMB200_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// adaptive sampler control (device/device_adaptive_sampler.c, device_renderer.c:350-376)
// ---------------------------------------------------------------------------------------------
extern "C" Lumb200Result lumb200_device_update_adaptive_sampling(Lumb200Device* d, const Lumb200AdaptiveSampling* params) {
  LB_REQUIRE(d && params, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(params->update_interval >= 1 || !params->enable, LUMB200_ERROR_INVALID_API_ARGUMENT, "update_interval must be >= 1");
  LB_REQUIRE(params->output_mode <= 3, LUMB200_ERROR_INVALID_API_ARGUMENT, "Invalid adaptive sampling output mode.");
  d->as_params = *params;
  // adaptive_sampler_setup, device_adaptive_sampler.c:46-48
  uint32_t mx = params->max_sampling_rate < 1 ? 1 : params->max_sampling_rate;
  mx          = mx > 256 ? 256 : mx;
  uint32_t av = params->avg_sampling_rate < 1 ? 1 : params->avg_sampling_rate;
  d->as_params.max_sampling_rate = mx;
  d->as_params.avg_sampling_rate = av > mx ? mx : av;
  if (!params->exposure_aware)
    d->as_params.exposure = 0.0f;
  return LUMB200_SUCCESS;
}

static Lumb200Result adaptive_build_stage(Lumb200Device* d) {
  Lumb200OutputParams tm;
  memset(&tm, 0, sizeof(tm));
  tm.exposure       = d->as_params.exposure;
  tm.tonemap        = d->as_params.tonemap;
  tm.agx_slope      = d->as_params.agx_slope;
  tm.agx_power      = d->as_params.agx_power;
  tm.agx_saturation = d->as_params.agx_saturation;
  lb_launch_adaptive_build_stage(d->planes, d->settings.width, d->settings.height, make_adaptive(d), tm, d->as_stage, d->as_params.max_sampling_rate,
                                 d->as_params.avg_sampling_rate, d->d_as_words, d->d_as_block_var, d->d_as_var_sum, d->d_as_prefix, d->d_as_total,
                                 d->stream);
  d->launches += 3;
  // the host needs the task count of the new stage to tile its executions (adaptive_sampler_compute_next_stage downloads it too)
  LB_CHECK(cudaMemcpyAsync(&d->as_total_tasks, d->d_as_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  d->as_stage++;
  return LUMB200_SUCCESS;
}

static Lumb200Result adaptive_execute(Lumb200Device* d);

// next free (start, end) event pair of the render-time bookkeeping (device_renderer.c:593-652: cumulative GPU seconds of the passes)
static Lumb200Result next_time_events(Lumb200Device* d, size_t* slot) {
  if (d->events_pending == d->ev_start.size()) {
    if (d->ev_start.size() >= 64) {
      LB_TRY(collect_events(d));
    }
    else {
      cudaEvent_t a, b;
      LB_CHECK(cudaEventCreate(&a));
      LB_CHECK(cudaEventCreate(&b));
      d->ev_start.push_back(a);
      d->ev_end.push_back(b);
    }
  }
  *slot = d->events_pending++;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_render_executions(Lumb200Device* d, uint32_t count) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_TRY(check_ready(d, true));
  LB_REQUIRE(d->as_active, LUMB200_ERROR_API_EXCEPTION, "adaptive sampling is not enabled (lumb200_device_update_adaptive_sampling + start_render)");
  LB_TRY(make_current(d));
  LB_TRY(ensure_light_records(d));
  for (uint32_t e = 0; e < count; e++) {
    if (d->events_pending == d->ev_start.size()) {
      if (d->ev_start.size() >= 64) {
        LB_TRY(collect_events(d));
      }
      else {
        cudaEvent_t a, b;
        LB_CHECK(cudaEventCreate(&a));
        LB_CHECK(cudaEventCreate(&b));
        d->ev_start.push_back(a);
        d->ev_end.push_back(b);
      }
    }
    const size_t ev = d->events_pending++;
    LB_CHECK(cudaEventRecord(d->ev_start[ev], d->stream));
    LB_TRY(adaptive_execute(d));
    LB_CHECK(cudaEventRecord(d->ev_end[ev], d->stream));
    d->as_executions[d->as_stage]++;
    d->samples_done++;
    if (d->as_stage < LB_ADAPTIVE_STAGES && d->as_executions[d->as_stage] >= (d->as_params.update_interval << d->as_stage))
      LB_TRY(adaptive_build_stage(d));
  }
  LB_CHECK(cudaGetLastError());
  return LUMB200_SUCCESS;
}

// one execution of the current stage, no bookkeeping (shared by render_executions and render_allocated_execution)
static Lumb200Result adaptive_execute(Lumb200Device* d) {
  if (d->as_stage == 0) {
    // stage 0: one sample per pixel with the same sample id everywhere -> the table-driven uniform pass
    LB_REQUIRE(d->as_executions[0] < (1u << 20), LUMB200_ERROR_INVALID_API_ARGUMENT, "sample id exceeds MAX_NUM_GLOBAL_SAMPLES");
    LB_TRY(render_pass(d, d->as_executions[0]));
    d->as_paths += (uint64_t) d->settings.width * d->settings.height;
  }
  else {
    const uint32_t capacity = d->paths_capacity;
    for (uint32_t begin = 0; begin < d->as_total_tasks; begin += capacity) {
      AdaptiveChunk chunk = {begin, (d->as_total_tasks - begin < capacity) ? d->as_total_tasks - begin : capacity};
      LB_TRY(render_pass(d, 0, false, true, &chunk));
    }
    d->as_paths += d->as_total_tasks;
  }
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_set_adaptive_state(Lumb200Device* d, uint32_t stage_id, const uint32_t* executions, const uint32_t* words,
                                                           size_t num_words) {
  LB_REQUIRE(d && executions, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d->as_active, LUMB200_ERROR_API_EXCEPTION, "adaptive sampling is not enabled");
  LB_REQUIRE(stage_id <= LB_ADAPTIVE_STAGES, LUMB200_ERROR_INVALID_API_ARGUMENT, "stage id %u is out of range", stage_id);
  LB_REQUIRE(!words || num_words == (size_t) d->as_bw * d->as_bh, LUMB200_ERROR_INVALID_API_ARGUMENT, "expected blocks_x * blocks_y words");
  LB_TRY(make_current(d));
  d->as_stage = stage_id;
  for (int k = 0; k <= LB_ADAPTIVE_STAGES; k++)
    d->as_executions[k] = executions[k];
  if (words) {
    LB_CHECK(cudaMemcpyAsync(d->d_as_words, words, sizeof(uint32_t) * num_words, cudaMemcpyHostToDevice, d->stream));
    if (stage_id > 0) {
      // tasks per block + prefix sums of the imposed stage (what the building device derived with the counts)
      std::vector<uint32_t> prefix(num_words);
      uint32_t run = 0;
      for (size_t b = 0; b < num_words; b++) {
        run += (((words[b] >> ((stage_id - 1u) * 8u)) & 0xFFu) + 1u) * 16u;
        prefix[b] = run;
      }
      LB_CHECK(cudaMemcpyAsync(d->d_as_prefix, prefix.data(), sizeof(uint32_t) * num_words, cudaMemcpyHostToDevice, d->stream));
      LB_CHECK(cudaStreamSynchronize(d->stream));
      d->as_total_tasks = run;
    }
    else
      LB_CHECK(cudaStreamSynchronize(d->stream));
  }
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_render_allocated_execution(Lumb200Device* d, const uint32_t* executions_before) {
  LB_REQUIRE(d && executions_before, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_TRY(check_ready(d, true));
  LB_REQUIRE(d->as_active, LUMB200_ERROR_API_EXCEPTION, "adaptive sampling is not enabled");
  LB_TRY(make_current(d));
  LB_TRY(ensure_light_records(d));
  uint32_t keep[LB_ADAPTIVE_STAGES + 1];
  for (int k = 0; k <= LB_ADAPTIVE_STAGES; k++) {
    keep[k]             = d->as_executions[k];
    d->as_executions[k] = executions_before[k];
  }
  size_t ev = 0;
  LB_TRY(next_time_events(d, &ev));
  LB_CHECK(cudaEventRecord(d->ev_start[ev], d->stream));
  const Lumb200Result r = adaptive_execute(d);
  LB_CHECK(cudaEventRecord(d->ev_end[ev], d->stream));
  for (int k = 0; k <= LB_ADAPTIVE_STAGES; k++)
    d->as_executions[k] = keep[k];
  d->samples_done++;
  LB_TRY(r);
  LB_CHECK(cudaGetLastError());
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_build_adaptive_stage(Lumb200Device* d) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(d->as_active && d->as_stage < LB_ADAPTIVE_STAGES, LUMB200_ERROR_API_EXCEPTION, "no adaptive stage left to build");
  LB_TRY(make_current(d));
  return adaptive_build_stage(d);
}

extern "C" Lumb200Result lumb200_device_get_adaptive_state(Lumb200Device* d, Lumb200AdaptiveState* state) {
  LB_REQUIRE(d && state, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  memset(state, 0, sizeof(*state));
  state->stage_id = d->as_stage;
  for (int k = 0; k <= LB_ADAPTIVE_STAGES; k++)
    state->executions[k] = d->as_executions[k];
  state->tasks_per_execution = d->as_active ? d->as_total_tasks : d->settings.width * d->settings.height;
  state->blocks_x            = d->as_bw;
  state->blocks_y            = d->as_bh;
  state->paths_traced        = d->as_paths;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_download_adaptive_words(Lumb200Device* d, uint32_t* words, size_t count) {
  LB_REQUIRE(d && words, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d->as_active && d->d_as_words && count == (size_t) d->as_bw * d->as_bh, LUMB200_ERROR_INVALID_API_ARGUMENT,
             "adaptive sampling is off or the buffer does not hold blocks_x * blocks_y words");
  LB_TRY(make_current(d));
  LB_CHECK(cudaMemcpyAsync(words, d->d_as_words, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_render_samples(Lumb200Device* d, uint32_t first_sample_id, uint32_t count, uint32_t stride) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_TRY(check_ready(d, true));
  LB_REQUIRE(stride >= 1, LUMB200_ERROR_INVALID_API_ARGUMENT, "stride must be >= 1");
  LB_TRY(make_current(d));
  LB_TRY(ensure_light_records(d));
  for (uint32_t k = 0; k < count; k++) {
    const uint32_t sample_id = first_sample_id + k * stride;
    LB_REQUIRE(sample_id < (1u << 20), LUMB200_ERROR_INVALID_API_ARGUMENT, "sample id %u exceeds MAX_NUM_GLOBAL_SAMPLES", sample_id);
    if (d->events_pending == d->ev_start.size()) {
      if (d->ev_start.size() >= 64) {
        LB_TRY(collect_events(d));
      }
      else {
        cudaEvent_t a, b;
        LB_CHECK(cudaEventCreate(&a));
        LB_CHECK(cudaEventCreate(&b));
        d->ev_start.push_back(a);
        d->ev_end.push_back(b);
      }
    }
    const size_t e = d->events_pending++;
    LB_CHECK(cudaEventRecord(d->ev_start[e], d->stream));
    LB_TRY(render_pass(d, sample_id));
    LB_CHECK(cudaEventRecord(d->ev_end[e], d->stream));
    d->samples_done++;
  }
  LB_CHECK(cudaGetLastError());
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_sync(Lumb200Device* d) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_TRY(make_current(d));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LB_TRY(collect_events(d));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_get_frame_planes(Lumb200Device* d, void** device_ptr, size_t* num_floats) {
  LB_REQUIRE(d && device_ptr && num_floats, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d->planes, LUMB200_ERROR_API_EXCEPTION, "settings have not been set");
  *device_ptr = d->planes;
  *num_floats = d->planes_floats;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_bind_frame_planes(Lumb200Device* d, void* device_ptr, size_t num_floats) {
  LB_REQUIRE(d && device_ptr, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d->settings.width, LUMB200_ERROR_API_EXCEPTION, "settings have not been set");
  LB_REQUIRE(num_floats == 4 * (size_t) d->settings.width * d->settings.height, LUMB200_ERROR_INVALID_API_ARGUMENT,
             "plane buffer must hold 4 * width * height floats");
  LB_TRY(make_current(d));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  if (!d->planes_external)
    dev_free(d, d->planes);
  d->planes          = (float*) device_ptr;
  d->planes_external = true;
  d->planes_floats   = num_floats;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_download_frame_planes(Lumb200Device* d, float* dst, size_t num_floats) {
  LB_REQUIRE(d && dst, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d->planes && num_floats == d->planes_floats, LUMB200_ERROR_INVALID_API_ARGUMENT, "plane buffer size mismatch");
  LB_TRY(make_current(d));
  LB_CHECK(cudaMemcpyAsync(dst, d->planes, sizeof(float) * num_floats, cudaMemcpyDeviceToHost, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LB_TRY(collect_events(d));
  return LUMB200_SUCCESS;
}

// accumulation_generate_result with the adaptive sampler's per-pixel counts / output mode and the camera's local error minimisation
static void resolve_adaptive(Lumb200Device* d, uint32_t sample_count, uint32_t local_error_minimization) {
  Lumb200OutputParams tm;
  memset(&tm, 0, sizeof(tm));
  tm.exposure       = d->as_params.exposure != 0.0f ? d->as_params.exposure : 1.0f;
  tm.tonemap        = d->as_params.tonemap;
  tm.agx_slope      = d->as_params.agx_slope;
  tm.agx_power      = d->as_params.agx_power;
  tm.agx_saturation = d->as_params.agx_saturation;
  LbAdaptive A      = make_adaptive(d);
  if (!d->as_active)
    A.words = nullptr;
  lb_launch_resolve(d->planes, d->d_result, d->settings.width, d->settings.height, A, sample_count, d->as_active ? d->as_params.output_mode : 0,
                    local_error_minimization, d->as_stage, tm, d->stream_grid, d->stream);
}

extern "C" Lumb200Result lumb200_device_download_result(Lumb200Device* d, uint32_t sample_count, float* dst) {
  LB_REQUIRE(d && dst, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d->planes && sample_count > 0, LUMB200_ERROR_INVALID_API_ARGUMENT, "nothing to resolve");
  LB_TRY(make_current(d));
  const size_t n = (size_t) d->settings.width * d->settings.height;
  if (d->as_active)
    resolve_adaptive(d, sample_count, 0);
  else
    lb_launch_generate_result(d->planes, d->d_result, (uint32_t) n, sample_count, d->stream_grid, d->stream);
  d->launches++;
  LB_CHECK(cudaMemcpyAsync(dst, d->d_result, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LB_TRY(collect_events(d));
  return LUMB200_SUCCESS;
}


extern "C" Lumb200Result lumb200_device_download_result_async(Lumb200Device* d, uint32_t sample_count, float* dst, uint32_t slot) {
  LB_REQUIRE(d && dst, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(slot < 2, LUMB200_ERROR_INVALID_API_ARGUMENT, "slot must be 0 or 1");
  LB_REQUIRE(d->planes && sample_count > 0, LUMB200_ERROR_INVALID_API_ARGUMENT, "nothing to resolve");
  LB_REQUIRE(!d->slot_pending[slot], LUMB200_ERROR_API_EXCEPTION, "slot %u still has a download in flight: wait for it first", slot);
  LB_TRY(make_current(d));
  const size_t n = (size_t) d->settings.width * d->settings.height;
  if (!d->copy_stream) {
    LB_CHECK(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      LB_CHECK(cudaEventCreateWithFlags(&d->ev_resolved[k], cudaEventDisableTiming));
      LB_CHECK(cudaEventCreateWithFlags(&d->ev_copied[k], cudaEventDisableTiming));
    }
  }
  if (!d->d_result_async[slot])
    LB_TRY(dev_alloc(d, &d->d_result_async[slot], 3 * n));
  float* keep = d->d_result;
  d->d_result = d->d_result_async[slot];  // resolve_adaptive / generate_result write d->d_result
  if (d->as_active)
    resolve_adaptive(d, sample_count, 0);
  else
    lb_launch_generate_result(d->planes, d->d_result, (uint32_t) n, sample_count, d->stream_grid, d->stream);
  d->d_result = keep;
  d->launches++;
  LB_CHECK(cudaEventRecord(d->ev_resolved[slot], d->stream));
  LB_CHECK(cudaStreamWaitEvent(d->copy_stream, d->ev_resolved[slot], 0));
  LB_CHECK(cudaMemcpyAsync(dst, d->d_result_async[slot], sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, d->copy_stream));
  LB_CHECK(cudaEventRecord(d->ev_copied[slot], d->copy_stream));
  d->slot_pending[slot] = true;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_wait_download(Lumb200Device* d, uint32_t slot) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(slot < 2, LUMB200_ERROR_INVALID_API_ARGUMENT, "slot must be 0 or 1");
  if (!d->slot_pending[slot])
    return LUMB200_SUCCESS;
  LB_TRY(make_current(d));
  LB_CHECK(cudaEventSynchronize(d->ev_copied[slot]));
  d->slot_pending[slot] = false;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_load_bluenoise_1d(Lumb200Device* d, const uint16_t* bluenoise_1d, size_t count) {
  LB_REQUIRE(d && bluenoise_1d, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(count == 256 * 256, LUMB200_ERROR_INVALID_API_ARGUMENT, "the 1D blue-noise mask must hold 256 x 256 entries, got %zu", count);
  LB_TRY(make_current(d));
  if (!d->d_bluenoise_1d)
    LB_TRY(dev_alloc(d, &d->d_bluenoise_1d, count));
  LB_CHECK(cudaMemcpyAsync(d->d_bluenoise_1d, bluenoise_1d, count * sizeof(uint16_t), cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_download_output_argb8(Lumb200Device* d, uint32_t sample_count, const Lumb200OutputParams* params,
                                                              uint8_t* dst) {
  LB_REQUIRE(d && params && dst, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d->planes && d->d_output && sample_count > 0, LUMB200_ERROR_INVALID_API_ARGUMENT, "nothing to resolve");
  LB_REQUIRE(params->tonemap <= 6, LUMB200_ERROR_INVALID_API_ARGUMENT, "unknown tone map %u", params->tonemap);
  LB_REQUIRE(params->supersampling <= 3 && (d->settings.width >> params->supersampling) > 0 && (d->settings.height >> params->supersampling) > 0,
             LUMB200_ERROR_INVALID_API_ARGUMENT, "supersampling %u does not fit the rendered resolution", params->supersampling);
  LB_REQUIRE(!params->dithering || d->d_bluenoise_1d, LUMB200_ERROR_MISSING_DATA, "dithering needs the 1D blue-noise mask");
  LB_TRY(make_current(d));
  const size_t n = (size_t) (d->settings.width >> params->supersampling) * (d->settings.height >> params->supersampling);
  const uint32_t W = d->settings.width, H = d->settings.height;
  const bool raw           = d->shading_mode != 0;  // tonemap_apply and device_post_apply are no-ops under a debug shading mode
  const uint32_t mip_count = (params->bloom_blend > 0.0f && !raw) ? lb_bloom_mip_count(W, H) : 0;
  if (mip_count > 1 || d->as_active || params->local_error_minimization) {
    // device_output_generate_output: accumulation_generate_result -> device_post_apply (bloom) -> generate_final_image
    if (mip_count > 1 && (d->bloom_w != W || d->bloom_h != H)) {
      for (float*& m : d->bloom_mips)
        dev_free(d, m);
      d->bloom_mips.assign(mip_count, nullptr);
      for (uint32_t i = 0; i < mip_count; i++)
        LB_TRY(dev_alloc(d, &d->bloom_mips[i], (size_t) (W >> (i + 1)) * (H >> (i + 1))));
      d->bloom_w = W, d->bloom_h = H;
    }
    if (d->as_active || params->local_error_minimization)
      resolve_adaptive(d, sample_count, params->local_error_minimization);
    else
      lb_launch_generate_result(d->planes, d->d_result, W * H, sample_count, d->stream_grid, d->stream);
    // device_post_apply only blooms the beauty output (device_post.c:216-220)
    if (mip_count > 1 && !(d->as_active && d->as_params.output_mode != 0))
      lb_launch_bloom(d->d_result, W, H, d->bloom_mips.data(), mip_count, params->bloom_blend, d->stream_grid, d->stream);
    lb_launch_output_argb8(d->d_result, W, H, 1, *params, d->d_bluenoise_1d, d->d_output, d->stream_grid, d->stream, raw);
    d->launches += 2 + 6 * mip_count;
  }
  else {
    lb_launch_output_argb8(d->planes, W, H, sample_count, *params, d->d_bluenoise_1d, d->d_output, d->stream_grid, d->stream, raw);
    d->launches++;
  }
  LB_CHECK(cudaMemcpyAsync(dst, d->d_output, 4 * n, cudaMemcpyDeviceToHost, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LB_TRY(collect_events(d));
  return LUMB200_SUCCESS;
}

// dst[i] += src[i]; the reference's buffer_add (cuda/kernels.cuh:646-675)
__global__ void __launch_bounds__(256) k_planes_add(float4* __restrict__ dst, const float4* __restrict__ src, size_t n4) {
  for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x) {
    float4 a       = dst[i];
    const float4 b = src[i];
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    dst[i] = a;
  }
}

extern "C" Lumb200Result lumb200_device_add_planes_from(Lumb200Device* d, Lumb200Device* other) {
  LB_REQUIRE(d && other, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d != other, LUMB200_ERROR_INVALID_API_ARGUMENT, "a device cannot be combined with itself");
  LB_REQUIRE(d->planes && other->planes && d->planes_floats == other->planes_floats && (d->planes_floats % 4) == 0,
             LUMB200_ERROR_INVALID_API_ARGUMENT, "the devices render different resolutions");
  // drain the producer first
  LB_TRY(lumb200_device_sync(other));
  LB_TRY(make_current(d));
  if (d->peer_floats != d->planes_floats) {
    dev_free(d, d->d_peer_planes);
    LB_TRY(dev_alloc(d, &d->d_peer_planes, d->planes_floats));
    d->peer_floats = d->planes_floats;
  }
  LB_CHECK(cudaMemcpyPeerAsync(d->d_peer_planes, d->cuda_index, other->planes, other->cuda_index, sizeof(float) * d->planes_floats, d->stream));
  k_planes_add<<<d->stream_grid, 256, 0, d->stream>>>((float4*) d->planes, (const float4*) d->d_peer_planes, d->planes_floats / 4);
  LB_CHECK(cudaGetLastError());
  d->launches++;
  LB_CHECK(cudaStreamSynchronize(d->stream));
  return LUMB200_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// parity / measurement hooks
// ---------------------------------------------------------------------------------------------
static Lumb200Result fetch_hits(Lumb200Device* d, uint32_t n, uint32_t* instance_ids, uint32_t* tri_ids, float* t, float* u, float* v) {
  DevTmp d_inst, d_tri, d_t, d_u, d_v;
  LB_CHECK(d_inst.alloc(sizeof(uint32_t) * n));
  LB_CHECK(d_tri.alloc(sizeof(uint32_t) * n));
  LB_CHECK(d_t.alloc(sizeof(float) * n));
  LB_CHECK(d_u.alloc(sizeof(float) * n));
  LB_CHECK(d_v.alloc(sizeof(float) * n));
  lb_launch_extract_hits(d->paths, d->d_prim_handle, d->d_uv, n, d_inst.as<uint32_t>(), d_tri.as<uint32_t>(), d_t.as<float>(), d_u.as<float>(),
                         d_v.as<float>(), d->stream_grid, d->stream);
  d->launches++;
  LB_CHECK(cudaGetLastError());
  if (instance_ids)
    LB_CHECK(cudaMemcpyAsync(instance_ids, d_inst.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, d->stream));
  if (tri_ids)
    LB_CHECK(cudaMemcpyAsync(tri_ids, d_tri.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, d->stream));
  if (t)
    LB_CHECK(cudaMemcpyAsync(t, d_t.p, sizeof(float) * n, cudaMemcpyDeviceToHost, d->stream));
  if (u)
    LB_CHECK(cudaMemcpyAsync(u, d_u.p, sizeof(float) * n, cudaMemcpyDeviceToHost, d->stream));
  if (v)
    LB_CHECK(cudaMemcpyAsync(v, d_v.p, sizeof(float) * n, cudaMemcpyDeviceToHost, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_trace_primary(Lumb200Device* d, uint32_t sample_id, uint32_t* instance_ids, uint32_t* tri_ids, float* t,
                                                      float* u, float* v) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_TRY(check_ready(d, false));
  LB_TRY(make_current(d));
  const LbFrame F  = make_frame(d);
  const uint32_t n = F.width * F.height;
  lb_launch_raygen(d->paths, F, d->camera, d->d_bluenoise, sample_id, d->queue[0], d->counters, d->stream_grid, d->stream);
  {
    const LbTexScene tex_scene = make_tex_scene(d);
    lb_launch_trace_closest(make_bvh(d->bvh), d->paths, d->queue[0], d->counters, d->d_uv, d->trace_grid, d->stream, false,
                            d->any_albedo_tex ? &tex_scene : nullptr);
  }
  d->launches += 2;
  LB_CHECK(cudaGetLastError());
  return fetch_hits(d, n, instance_ids, tri_ids, t, u, v);
}

// device_get_gbuffer_meta (device/device.h:190; filled by optix_kernel_raytrace.cu:45-75 on the first pass): closest hit of ONE pixel's
// primary ray. Must not run while sample passes are queued on the wavefront (the call synchronises first).
extern "C" Lumb200Result lumb200_device_query_pixel(Lumb200Device* d, uint32_t x, uint32_t y, uint32_t sample_id, uint32_t* instance_id,
                                                    uint32_t* tri_id, float* depth, float* ray) {
  LB_REQUIRE(d && instance_id && tri_id && depth && ray, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_TRY(check_ready(d, false));
  LB_REQUIRE(x < d->settings.width && y < d->settings.height, LUMB200_ERROR_INVALID_API_ARGUMENT, "pixel (%u, %u) is outside the %ux%u frame", x, y,
             d->settings.width, d->settings.height);
  LB_TRY(make_current(d));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LbCounters before;
  LB_CHECK(cudaMemcpy(&before, d->counters, sizeof(before), cudaMemcpyDeviceToHost));
  const LbFrame F = make_frame(d);
  lb_launch_raygen_pixel(d->paths, F, d->camera, d->d_bluenoise, x, y, sample_id, d->queue[0], d->counters, d->stream);
  {
    const LbTexScene tex_scene = make_tex_scene(d);
    lb_launch_trace_closest(make_bvh(d->bvh), d->paths, d->queue[0], d->counters, d->d_uv, d->trace_grid, d->stream, false,
                            d->any_albedo_tex ? &tex_scene : nullptr);
  }
  d->launches += 2;
  float t = 0.0f, u = 0.0f, v = 0.0f;
  LB_TRY(fetch_hits(d, 1, instance_id, tri_id, &t, &u, &v));
  float4 dir;
  LB_CHECK(cudaMemcpy(&dir, d->paths.dir, sizeof(dir), cudaMemcpyDeviceToHost));
  ray[0] = dir.x, ray[1] = dir.y, ray[2] = dir.z;
  if (*instance_id == LB_HIT_SKY) {
    *instance_id = 0xFFFFFFFFu;  // HIT_TYPE_INVALID
    *tri_id      = 0xFFFFFFFFu;
    *depth       = 3.402823466e+38f;
  }
  else
    *depth = t;
  LB_CHECK(cudaMemcpy(d->counters, &before, sizeof(before), cudaMemcpyHostToDevice));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_trace_rays(Lumb200Device* d, const float* origins, const float* directions, uint32_t count,
                                                   uint32_t* instance_ids, uint32_t* tri_ids, float* t, float* u, float* v) {
  LB_REQUIRE(d && (count == 0 || (origins && directions)), LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(!d->accel_dirty && d->bvh.nodes, LUMB200_ERROR_API_EXCEPTION, "acceleration structure is out of date: call lumb200_device_build_accel");
  if (count == 0)
    return LUMB200_SUCCESS;
  LB_TRY(make_current(d));
  LB_TRY(ensure_paths(d, count));
  DevTmp d_o, d_d;
  LB_CHECK(d_o.alloc(sizeof(float) * 3 * (size_t) count));
  LB_CHECK(d_d.alloc(sizeof(float) * 3 * (size_t) count));
  LB_CHECK(cudaMemcpyAsync(d_o.p, origins, sizeof(float) * 3 * (size_t) count, cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d_d.p, directions, sizeof(float) * 3 * (size_t) count, cudaMemcpyHostToDevice, d->stream));
  lb_launch_load_rays(d->paths, d_o.as<float>(), d_d.as<float>(), count, d->queue[0], d->counters, d->stream_grid, d->stream);
  {
    const LbTexScene tex_scene = make_tex_scene(d);
    lb_launch_trace_closest(make_bvh(d->bvh), d->paths, d->queue[0], d->counters, d->d_uv, d->trace_grid, d->stream, false,
                            d->any_albedo_tex ? &tex_scene : nullptr);
  }
  d->launches += 2;
  return fetch_hits(d, count, instance_ids, tri_ids, t, u, v);
}

extern "C" Lumb200Result lumb200_device_shade_vertices(Lumb200Device* d, uint32_t sample_id, uint32_t rng_depth, uint32_t is_last_iteration,
                                                        const Lumb200VertexIn* vertices, uint32_t count, Lumb200VertexOut* out) {
  LB_REQUIRE(d && (count == 0 || (vertices && out)), LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_TRY(check_ready(d, true));
  if (d->sky.mode == 1 && !d->hdri_valid) {
    LB_TRY(make_current(d));
    LB_TRY(build_sky_hdri(d));
  }
  LB_REQUIRE(count <= d->paths_capacity, LUMB200_ERROR_INVALID_API_ARGUMENT, "%u vertices exceed the wavefront capacity %u", count, d->paths_capacity);
  LB_REQUIRE(sample_id < (1u << 20) && rng_depth < LB_RNG_TABLE_DEPTHS, LUMB200_ERROR_INVALID_API_ARGUMENT, "sample id / depth out of range");
  for (uint32_t i = 0; i < count; i++)
    LB_REQUIRE((vertices[i].prim < d->num_prims || vertices[i].prim == 0xFFFFFFFFu) && vertices[i].pixel_x < d->settings.width
                 && vertices[i].pixel_y < d->settings.height,
               LUMB200_ERROR_INVALID_API_ARGUMENT, "vertex %u: primitive %u / pixel (%u, %u) out of range", i, vertices[i].prim, vertices[i].pixel_x,
               vertices[i].pixel_y);
  if (count == 0)
    return LUMB200_SUCCESS;
  LB_TRY(make_current(d));
  LB_TRY(ensure_light_records(d));
  const LbFrame F = make_frame(d);
  const Bvh8 bvh  = make_bvh(d->bvh);
  const LbTexScene tex_scene = make_tex_scene(d);
  DevTmp d_in, d_out;
  LB_CHECK(d_in.alloc(sizeof(Lumb200VertexIn) * (size_t) count));
  LB_CHECK(d_out.alloc(sizeof(Lumb200VertexOut) * (size_t) count));
  LB_CHECK(cudaMemcpyAsync(d_in.p, vertices, sizeof(Lumb200VertexIn) * (size_t) count, cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemsetAsync(d_out.p, 0, sizeof(Lumb200VertexOut) * (size_t) count, d->stream));
  // the hook must not disturb the public ray counters
  LbCounters before;
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LB_CHECK(cudaMemcpy(&before, d->counters, sizeof(before), cudaMemcpyDeviceToHost));
  lb_launch_load_vertices(d->paths, d_in.as<Lumb200VertexIn>(), count, F.width, sample_id, d->queue[0], d->counters, d->stream_grid, d->stream);
  lb_launch_rng_table(d->d_rng_table, sample_id, rng_depth + 1, d->stream);
  LbShadeParams sp = make_shade_params(d, F, sample_id, false);
  const bool was_profiling = d->profiling;
  d->profiling             = false;
  surface_stages(d, sp, bvh, 0, rng_depth, is_last_iteration != 0, false, d->any_albedo_tex ? &tex_scene : nullptr);
  d->profiling             = was_profiling;
  lb_launch_extract_segments(d->paths, d->counters, d_out.as<Lumb200VertexOut>(), d->stream_grid, d->stream);
  lb_launch_extract_vertices(d->paths, count, d->queue[0], d->counters, d_out.as<Lumb200VertexOut>(), d->stream_grid, d->stream);
  d->launches += 4;
  LB_CHECK(cudaGetLastError());
  LB_CHECK(cudaMemcpyAsync(out, d_out.p, sizeof(Lumb200VertexOut) * (size_t) count, cudaMemcpyDeviceToHost, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LbCounters after;
  LB_CHECK(cudaMemcpy(&after, d->counters, sizeof(after), cudaMemcpyDeviceToHost));
  before.stack_overflow = after.stack_overflow;
  LB_CHECK(cudaMemcpy(d->counters, &before, sizeof(before), cudaMemcpyHostToDevice));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_trace_shadow_rays(Lumb200Device* d, const float* origins, const float* directions, const float* max_dist,
                                                           const uint32_t* ignore_prims, const uint32_t* target_prims, uint32_t count,
                                                           float* visibility) {
  LB_REQUIRE(d && (count == 0 || (origins && directions && max_dist && ignore_prims && target_prims && visibility)), LUMB200_ERROR_ARGUMENT_NULL,
             "NULL argument");
  LB_TRY(check_ready(d, true));  // the any-hit response needs the materials
  LB_REQUIRE(count <= d->paths_capacity, LUMB200_ERROR_INVALID_API_ARGUMENT, "%u rays exceed the wavefront capacity %u", count, d->paths_capacity);
  if (count == 0)
    return LUMB200_SUCCESS;
  LB_TRY(make_current(d));
  const Bvh8 bvh = make_bvh(d->bvh);
  const LbTexScene tex_scene = make_tex_scene(d);
  DevTmp d_o, d_d, d_m, d_i, d_t, d_v;
  LB_CHECK(d_o.alloc(sizeof(float) * 3 * (size_t) count));
  LB_CHECK(d_d.alloc(sizeof(float) * 3 * (size_t) count));
  LB_CHECK(d_m.alloc(sizeof(float) * (size_t) count));
  LB_CHECK(d_i.alloc(sizeof(uint32_t) * (size_t) count));
  LB_CHECK(d_t.alloc(sizeof(uint32_t) * (size_t) count));
  LB_CHECK(d_v.alloc(sizeof(float) * 3 * (size_t) count));
  LB_CHECK(cudaMemcpyAsync(d_o.p, origins, sizeof(float) * 3 * (size_t) count, cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d_d.p, directions, sizeof(float) * 3 * (size_t) count, cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d_m.p, max_dist, sizeof(float) * (size_t) count, cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d_i.p, ignore_prims, sizeof(uint32_t) * (size_t) count, cudaMemcpyHostToDevice, d->stream));
  LB_CHECK(cudaMemcpyAsync(d_t.p, target_prims, sizeof(uint32_t) * (size_t) count, cudaMemcpyHostToDevice, d->stream));
  LbCounters before;
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LB_CHECK(cudaMemcpy(&before, d->counters, sizeof(before), cudaMemcpyDeviceToHost));
  lb_launch_load_shadow_rays(d->paths, d_o.as<float>(), d_d.as<float>(), d_m.as<float>(), d_i.as<uint32_t>(), d_t.as<uint32_t>(), count, d->counters,
                             d->stream_grid, d->stream);
  lb_launch_trace_shadow(bvh, d->paths, d->counters, d->d_prim_material, d->d_shadow_tab, d->trace_grid, d->stream, false,
                         d->any_albedo_tex ? &tex_scene : nullptr);
  lb_launch_extract_visibility(d->paths, count, d_v.as<float>(), d->stream_grid, d->stream);
  d->launches += 4;
  LB_CHECK(cudaGetLastError());
  LB_CHECK(cudaMemcpyAsync(visibility, d_v.p, sizeof(float) * 3 * (size_t) count, cudaMemcpyDeviceToHost, d->stream));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LbCounters after;
  LB_CHECK(cudaMemcpy(&after, d->counters, sizeof(after), cudaMemcpyDeviceToHost));
  before.stack_overflow = after.stack_overflow;
  LB_CHECK(cudaMemcpy(d->counters, &before, sizeof(before), cudaMemcpyHostToDevice));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_time_primary_trace(Lumb200Device* d, uint32_t sample_id, uint32_t repeats, float* avg_ms) {
  LB_REQUIRE(d && avg_ms, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(repeats > 0, LUMB200_ERROR_INVALID_API_ARGUMENT, "repeats must be > 0");
  LB_TRY(check_ready(d, false));
  LB_TRY(make_current(d));
  const LbFrame F = make_frame(d);
  cudaEvent_t a, b;
  LB_CHECK(cudaEventCreate(&a));
  LB_CHECK(cudaEventCreate(&b));
  float total = 0.0f;
  for (uint32_t k = 0; k < repeats; k++) {
    lb_launch_raygen(d->paths, F, d->camera, d->d_bluenoise, sample_id + k, d->queue[0], d->counters, d->stream_grid, d->stream);
    cudaEventRecord(a, d->stream);
    {
      const LbTexScene tex_scene = make_tex_scene(d);
      lb_launch_trace_closest(make_bvh(d->bvh), d->paths, d->queue[0], d->counters, nullptr, d->trace_grid, d->stream, false,
                              d->any_albedo_tex ? &tex_scene : nullptr);
    }
    cudaEventRecord(b, d->stream);
    LB_CHECK(cudaEventSynchronize(b));
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    total += ms;
    d->launches += 2;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *avg_ms = total / repeats;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_download_bvh(Lumb200Device* d, uint32_t which, void* nodes, size_t node_capacity, float* triangles,
                                                     size_t triangle_capacity, uint32_t* num_nodes, uint32_t* num_triangles) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(which <= 1, LUMB200_ERROR_INVALID_API_ARGUMENT, "which must be 0 or 1");
  LB_TRY(make_current(d));
  const LbBvhBuffers& b = which ? d->light_bvh : d->bvh;
  if (num_nodes)
    *num_nodes = b.nodes ? b.num_nodes : 0;
  if (num_triangles)
    *num_triangles = b.nodes ? b.num_tris : 0;
  if (!b.nodes)
    return LUMB200_SUCCESS;
  LB_CHECK(cudaStreamSynchronize(d->stream));
  if (nodes && node_capacity)
    LB_CHECK(cudaMemcpy(nodes, b.nodes, sizeof(Bvh8Node) * (node_capacity < b.num_nodes ? node_capacity : b.num_nodes), cudaMemcpyDeviceToHost));
  if (triangles && triangle_capacity && b.num_tris)
    LB_CHECK(cudaMemcpy(triangles, b.tris, sizeof(float) * 12 * (triangle_capacity < b.num_tris ? triangle_capacity : b.num_tris),
                        cudaMemcpyDeviceToHost));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_get_stats(Lumb200Device* d, Lumb200Stats* stats) {
  LB_REQUIRE(d && stats, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_TRY(make_current(d));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LB_TRY(collect_events(d));
  LbCounters c;
  LB_CHECK(cudaMemcpy(&c, d->counters, sizeof(c), cudaMemcpyDeviceToHost));
  memset(stats, 0, sizeof(*stats));
  stats->closest_rays        = c.closest_rays;
  stats->shadow_rays         = c.shadow_rays;
  stats->light_rays          = c.light_rays;
  stats->kernel_launches     = d->launches;
  stats->render_seconds      = d->render_seconds;
  stats->accel_build_seconds = d->accel_seconds;
  stats->samples_done        = d->samples_done;
  stats->bvh_nodes           = d->bvh.num_nodes;
  stats->bvh_tris            = d->bvh.num_tris;
  stats->light_bvh_nodes     = d->light_bvh.num_nodes;
  stats->device_bytes        = d->device_bytes;
  stats->bvh_depth           = d->bvh.depth;
  stats->light_bvh_depth     = d->light_bvh.depth;
  stats->bvh_sah_cost        = d->bvh.sah_cost;
  stats->bvh_ploc_radius     = (uint32_t) d->bvh.ploc_radius;
  stats->stack_overflows     = c.stack_overflow;
  stats->nonfinite_samples   = c.nonfinite_samples;
  stats->nonfinite_pixel     = c.nonfinite_pixel;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_set_profiling(Lumb200Device* d, uint32_t enable) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_TRY(make_current(d));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  d->prof_used = 0;
  memset(&d->profile, 0, sizeof(d->profile));
  d->profiling = enable != 0;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_get_profile(Lumb200Device* d, Lumb200Profile* profile) {
  LB_REQUIRE(d && profile, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_TRY(make_current(d));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  prof_collect(d);
  *profile = d->profile;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_measure_traversal(Lumb200Device* d, uint32_t sample_id, Lumb200TraversalStats* stats) {
  LB_REQUIRE(d && stats, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_TRY(check_ready(d, true));
  LB_TRY(make_current(d));
  LB_TRY(ensure_light_records(d));
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LbCounters before, after;
  LB_CHECK(cudaMemcpy(&before, d->counters, sizeof(before), cudaMemcpyDeviceToHost));
  const bool was_profiling = d->profiling;
  d->profiling             = false;
  Lumb200Result r          = render_pass(d, sample_id, true, false);
  d->profiling             = was_profiling;
  LB_TRY(r);
  LB_CHECK(cudaStreamSynchronize(d->stream));
  LB_CHECK(cudaMemcpy(&after, d->counters, sizeof(after), cudaMemcpyDeviceToHost));
  stats->closest_rays  = after.closest_rays - before.closest_rays;
  stats->closest_nodes = after.closest_nodes - before.closest_nodes;
  stats->closest_tris  = after.closest_tris - before.closest_tris;
  stats->shadow_rays   = after.shadow_rays - before.shadow_rays;
  stats->shadow_nodes  = after.shadow_nodes - before.shadow_nodes;
  stats->shadow_tris   = after.shadow_tris - before.shadow_tris;
  stats->light_rays    = after.light_rays - before.light_rays;
  stats->shaded_vertices     = after.shaded_vertices - before.shaded_vertices;
  stats->light_tree_nodes    = after.light_tree_nodes - before.light_tree_nodes;
  stats->light_root_sections = d->light_root_sections;
  stats->_pad                = 0;
  // keep the public ray counters untouched by the measurement
  LB_CHECK(cudaMemcpy(d->counters, &before, sizeof(before), cudaMemcpyHostToDevice));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_get_cuda_index(Lumb200Device* d, uint32_t* index) {
  LB_REQUIRE(d && index, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  *index = (uint32_t) d->cuda_index;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_get_adaptive_words_device(Lumb200Device* d, void** words, void** task_prefix, size_t* count) {
  LB_REQUIRE(d && words && task_prefix && count, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d->as_active && d->d_as_words, LUMB200_ERROR_API_EXCEPTION, "adaptive sampling is not active (enable it and start a render)");
  *words       = d->d_as_words;
  *task_prefix = d->d_as_prefix;
  *count       = (size_t) d->as_bw * d->as_bh;
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_adopt_adaptive_stage(Lumb200Device* d, uint32_t stage_id, const uint32_t* executions) {
  LB_REQUIRE(d && executions, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  LB_REQUIRE(d->as_active, LUMB200_ERROR_API_EXCEPTION, "adaptive sampling is not enabled");
  LB_REQUIRE(stage_id <= LB_ADAPTIVE_STAGES, LUMB200_ERROR_INVALID_API_ARGUMENT, "stage id %u is out of range", stage_id);
  LB_TRY(make_current(d));
  d->as_stage = stage_id;
  for (int k = 0; k <= LB_ADAPTIVE_STAGES; k++)
    d->as_executions[k] = executions[k];
  if (stage_id > 0) {
    // the stage's task count is the last entry of the (broadcast) inclusive prefix sums
    const size_t n = (size_t) d->as_bw * d->as_bh;
    LB_CHECK(cudaMemcpyAsync(&d->as_total_tasks, d->d_as_prefix + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, d->stream));
    LB_CHECK(cudaStreamSynchronize(d->stream));
  }
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_clear_frame_planes(Lumb200Device* d) {
  LB_REQUIRE(d, LUMB200_ERROR_ARGUMENT_NULL, "device is NULL");
  LB_REQUIRE(d->planes, LUMB200_ERROR_API_EXCEPTION, "settings have not been set");
  LB_TRY(make_current(d));
  LB_CHECK(cudaMemsetAsync(d->planes, 0, sizeof(float) * d->planes_floats, d->stream));
  return LUMB200_SUCCESS;
}

extern "C" Lumb200Result lumb200_device_get_stream(Lumb200Device* d, void** stream) {
  LB_REQUIRE(d && stream, LUMB200_ERROR_ARGUMENT_NULL, "NULL argument");
  *stream = (void*) d->stream;
  return LUMB200_SUCCESS;
}
