/* ref_dev_harness.cu - flat C entry points that launch the REFERENCE's own CUDA kernels, unmodified.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as oracle/lum_oracle.h): only tests/ may load oracle/_ref/librefdev.so. The product
 * never does. This file contains no reference code: it #includes the reference's device headers from the throw-away
 * copy that oracle/ref/ref_patch.sh prepares from /root/reference (three nvcc-12.9 compile fixes, no arithmetic change)
 * and compiles them with the reference's own flags (--use_fast_math) for sm_100a.
 *
 * What runs is the reference's kernel, as written:
 *   tasks_create                         cuda/kernels.cuh:45-193      (ray generation)
 *   geometry_process_tasks               cuda/geometry.cuh:11-180     (surface shading, NEE task creation, bounce, RR)
 *   sky_process_tasks                    cuda/sky.cuh:609-633         (miss shading; constant colour and the procedural atmosphere)
 *   sky_compute_transmittance_lut / sky_compute_multiscattering_lut   cuda/sky.cuh:144-330
 *   sky_compute_hdri                     cuda/sky_hdri.cuh:60-158     (HDRI mode bake)
 *   sky_process_inscattering_events      cuda/kernels.cuh:356-389     (aerial perspective)
 *   accumulation_collect_results[_first_sample], accumulation_generate_result   cuda/accumulation.cuh:36-190
 *   bsdf_generate_ss_lut / glossy_lut / dielectric_lut               cuda/bsdf_lut.cuh:20-209
 * The harness only owns what the reference's host C code owns: the `device` constant block (device_utils.h:567-617),
 * the work buffers (device_work_buffers.c) and the launch geometry (THREADS_PER_BLOCK x num_blocks, kernel.c:140-159).
 * Task records travel as raw bytes in the reference's warp-interleaved layout (cuda/memory.cuh:114-131); the tests do the
 * address arithmetic in numpy.
 *
 * NOT covered: the OptiX programs (device/optix/*.cu - closest hit, shadow / light-enumeration any-hit evaluation). They
 * need libnvoptix.so.1, which neither this container nor the GPU box has.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>

#include "accumulation.cuh"
#include "bsdf_lut.cuh"
#include "geometry.cuh"
#include "kernels.cuh"
#include "sky.cuh"
#include "sky_hdri.cuh"
#include "utils.cuh"

#define RD_CHECK(expr)                                                                                         \
  do {                                                                                                         \
    const cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) {                                                                                   \
      fprintf(stderr, "librefdev: %s failed: %s (%s:%d)\n", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                                                \
    }                                                                                                          \
  } while (0)

namespace {

struct Buffer {
  void* ptr    = nullptr;
  size_t bytes = 0;
};

struct Harness {
  DeviceConstantMemory host;  // mirror of the __constant__ block
  std::map<std::string, Buffer> buffers;
  uint32_t num_blocks = 0, tasks_per_thread = 0;
  cudaArray_t lut_arrays[4]      = {nullptr, nullptr, nullptr, nullptr};
  cudaTextureObject_t lut_tex[4] = {0, 0, 0, 0};
  cudaTextureObject_t sky_tex[4] = {0, 0, 0, 0};
  cudaTextureObject_t hdri_tex   = 0;
  cudaArray_t moon_arrays[2]     = {nullptr, nullptr};
  cudaTextureObject_t moon_tex[2] = {0, 0};
  bool dirty                     = true;
} g;

int alloc_buffer(const char* name, size_t bytes, void** out) {
  Buffer& b = g.buffers[name];
  if (b.ptr && b.bytes >= bytes) {
    RD_CHECK(cudaMemset(b.ptr, 0, b.bytes));
    *out = b.ptr;
    return 0;
  }
  if (b.ptr)
    cudaFree(b.ptr);
  b.bytes = bytes ? bytes : 16;
  RD_CHECK(cudaMalloc(&b.ptr, b.bytes));
  RD_CHECK(cudaMemset(b.ptr, 0, b.bytes));
  *out = b.ptr;
  return 0;
}

int upload_new(const char* name, const void* src, size_t bytes, void** out) {
  if (alloc_buffer(name, bytes, out))
    return 1;
  if (bytes)
    RD_CHECK(cudaMemcpy(*out, src, bytes, cudaMemcpyHostToDevice));
  return 0;
}

int sync_constant() {
  if (g.dirty) {
    RD_CHECK(cudaMemcpyToSymbol(device, &g.host, sizeof(DeviceConstantMemory)));
    g.dirty = false;
  }
  return 0;
}

int finish(const char* what) {
  const cudaError_t e1 = cudaGetLastError();
  const cudaError_t e2 = cudaDeviceSynchronize();
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    fprintf(stderr, "librefdev: %s: %s / %s\n", what, cudaGetErrorString(e1), cudaGetErrorString(e2));
    return 1;
  }
  return 0;
}

}  // namespace

extern "C" {

size_t refdev_sizeof(const char* what) {
  const std::string w(what);
  if (w == "DeviceConstantMemory") return sizeof(DeviceConstantMemory);
  if (w == "DeviceRendererSettings") return sizeof(DeviceRendererSettings);
  if (w == "DeviceCamera") return sizeof(DeviceCamera);
  if (w == "DeviceSky") return sizeof(DeviceSky);
  if (w == "DeviceTaskState") return sizeof(DeviceTaskState);
  if (w == "DeviceTaskDirectLight") return sizeof(DeviceTaskDirectLight);
  if (w == "DeviceTaskResult") return sizeof(DeviceTaskResult);
  return 0;
}

/* device_create + device_allocate_work_buffers (device.c:422-488, device_work_buffers.c): zeroed constant block, abort flag. */
int refdev_create(int cuda_device) {
  RD_CHECK(cudaSetDevice(cuda_device));
  memset(&g.host, 0, sizeof(g.host));
  void* p;
  if (alloc_buffer("abort_flag", sizeof(uint32_t), &p))
    return 1;
  g.host.ptrs.abort_flag = (uint32_t*) p;
  /* textures that are never supplied here read as absent (texture_utils.cuh:24-31), not as CUDA texture object 0: the moon's surface
   * (data/moon/*.png, embedded by the reference's build) is one of them - its disc is an occluder with albedo 0 */
  g.host.moon_albedo_tex.handle = TEXTURE_OBJECT_INVALID;
  g.host.moon_normal_tex.handle = TEXTURE_OBJECT_INVALID;
  g.dirty                       = true;
  return 0;
}

int refdev_set_settings(const void* data, size_t bytes) {
  if (bytes != sizeof(DeviceRendererSettings)) return 2;
  memcpy(&g.host.settings, data, bytes);
  g.dirty = true;
  return 0;
}

int refdev_set_camera(const void* data, size_t bytes) {
  if (bytes != sizeof(DeviceCamera)) return 2;
  memcpy(&g.host.camera, data, bytes);
  g.dirty = true;
  return 0;
}

int refdev_set_sky(const void* data, size_t bytes) {
  if (bytes != sizeof(DeviceSky)) return 2;
  memcpy(&g.host.sky, data, bytes);
  g.dirty = true;
  return 0;
}

/* device_load_embedded_data: bluenoise_1D (u16 x 65536) and bluenoise_2D (u32 x 65536) */
int refdev_set_bluenoise(const uint16_t* bn1d, size_t n1, const uint32_t* bn2d, size_t n2) {
  void* p;
  if (upload_new("bluenoise_1D", bn1d, n1 * sizeof(uint16_t), &p)) return 1;
  g.host.ptrs.bluenoise_1D = (const uint16_t*) p;
  if (upload_new("bluenoise_2D", bn2d, n2 * sizeof(uint32_t), &p)) return 1;
  g.host.ptrs.bluenoise_2D = (const uint32_t*) p;
  g.dirty                  = true;
  return 0;
}

/* device_add_mesh / device_update_instances / device_update_materials payloads, already in device format (packed by the
 * reference's own host packers through oracle/_ref/libref_host.so). */
int refdev_set_scene(uint32_t num_meshes, const void* const* vertices, const void* const* textris, const uint32_t* tri_counts,
                     uint32_t num_instances, const void* transforms, const uint32_t* mesh_ids, uint32_t num_materials, const void* materials) {
  const DeviceTriangleVertex** vptrs  = new const DeviceTriangleVertex*[num_meshes];
  const DeviceTriangleTexture** tptrs = new const DeviceTriangleTexture*[num_meshes];
  void* p;
  for (uint32_t m = 0; m < num_meshes; m++) {
    char name[64];
    snprintf(name, sizeof(name), "mesh_vertices_%u", m);
    if (upload_new(name, vertices[m], (size_t) tri_counts[m] * 3 * sizeof(DeviceTriangleVertex), &p)) return 1;
    vptrs[m] = (const DeviceTriangleVertex*) p;
    snprintf(name, sizeof(name), "mesh_textris_%u", m);
    if (upload_new(name, textris[m], (size_t) tri_counts[m] * sizeof(DeviceTriangleTexture), &p)) return 1;
    tptrs[m] = (const DeviceTriangleTexture*) p;
  }
  if (upload_new("vertices", vptrs, num_meshes * sizeof(void*), &p)) return 1;
  g.host.ptrs.vertices = (const DeviceTriangleVertex**) p;
  if (upload_new("texture_triangles", tptrs, num_meshes * sizeof(void*), &p)) return 1;
  g.host.ptrs.texture_triangles = (const DeviceTriangleTexture**) p;
  delete[] vptrs;
  delete[] tptrs;
  if (upload_new("instance_transforms", transforms, (size_t) num_instances * sizeof(DeviceTransform), &p)) return 1;
  g.host.ptrs.instance_transforms = (const DeviceTransform*) p;
  if (upload_new("instance_mesh_ids", mesh_ids, (size_t) num_instances * sizeof(uint32_t), &p)) return 1;
  g.host.ptrs.instance_mesh_ids = (const uint32_t*) p;
  if (upload_new("materials", materials, (size_t) num_materials * sizeof(DeviceMaterialCompressed), &p)) return 1;
  g.host.ptrs.materials = (const DeviceMaterialCompressed*) p;
  g.dirty               = true;
  return 0;
}

/* device_update_light_tree_data, device.c (LightTree blobs of device_light.h:102-113); num_lights == 0 removes the tree */
int refdev_set_light_tree(const void* root, size_t root_bytes, const void* nodes, size_t nodes_bytes, const uint32_t* handles, uint32_t num_lights) {
  if (num_lights == 0) {
    g.host.ptrs.light_tree_root           = nullptr;
    g.host.ptrs.light_tree_nodes          = nullptr;
    g.host.ptrs.light_tree_tri_handle_map = nullptr;
    g.dirty                               = true;
    return 0;
  }
  void* p;
  if (upload_new("light_tree_root", root, root_bytes, &p)) return 1;
  g.host.ptrs.light_tree_root = (const DeviceLightTreeRootHeader*) p;
  if (upload_new("light_tree_nodes", nodes, nodes_bytes, &p)) return 1;
  g.host.ptrs.light_tree_nodes = (const DeviceLightTreeNode*) p;
  if (upload_new("light_tree_tri_handle_map", handles, (size_t) num_lights * sizeof(TriangleHandle), &p)) return 1;
  g.host.ptrs.light_tree_tri_handle_map = (const TriangleHandle*) p;
  g.dirty                               = true;
  return 0;
}

/* bsdf_lut_generate (device_bsdf.c:56-125): the three reference kernels, then R16 unorm / linear / clamp / normalised
 * textures as device_bsdf.c:7-54 + device_texture.c:262-271 configure them. Host copies are returned for comparison. */
int refdev_build_bsdf_lut(uint16_t* conductor, uint16_t* glossy, uint16_t* dielectric, uint16_t* dielectric_inv) {
  const size_t n2 = BSDF_LUT_SIZE * BSDF_LUT_SIZE, n3 = n2 * BSDF_LUT_SIZE;
  void* d[4];
  if (alloc_buffer("lut_conductor", n2 * 2, &d[0]) || alloc_buffer("lut_glossy", n2 * 2, &d[1]) || alloc_buffer("lut_dielectric", n3 * 2, &d[2])
      || alloc_buffer("lut_dielectric_inv", n3 * 2, &d[3]))
    return 1;
  if (sync_constant()) return 1;
  const uint32_t blocks = (uint32_t) ((n3 + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK);
  // NUM_THREADS is read from device.config.num_blocks by THREAD_ID-independent code only; keep it consistent anyway
  const uint32_t saved  = g.host.config.num_blocks;
  g.host.config.num_blocks = blocks;
  g.dirty                  = true;
  if (sync_constant()) return 1;
  KernelArgsBSDFGenerateSSLUT a0;
  a0.dst = (uint16_t*) d[0];
  bsdf_generate_ss_lut<<<blocks, THREADS_PER_BLOCK>>>(a0);
  KernelArgsBSDFGenerateGlossyLUT a1;
  a1.dst           = (uint16_t*) d[1];
  a1.src_energy_ss = (const uint16_t*) d[0];
  bsdf_generate_glossy_lut<<<blocks, THREADS_PER_BLOCK>>>(a1);
  KernelArgsBSDFGenerateDielectricLUT a2;
  a2.dst     = (uint16_t*) d[2];
  a2.dst_inv = (uint16_t*) d[3];
  bsdf_generate_dielectric_lut<<<blocks, THREADS_PER_BLOCK>>>(a2);
  if (finish("bsdf_generate_*_lut")) return 1;
  g.host.config.num_blocks = saved;

  uint16_t* host_out[4] = {conductor, glossy, dielectric, dielectric_inv};
  const cudaChannelFormatDesc fmt = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
  DeviceTextureObject* objs[4]    = {&g.host.bsdf_lut_conductor, &g.host.bsdf_lut_glossy, &g.host.bsdf_lut_dielectric, &g.host.bsdf_lut_dielectric_inv};
  for (int k = 0; k < 4; k++) {
    const bool is3d = k >= 2;
    if (host_out[k])
      RD_CHECK(cudaMemcpy(host_out[k], d[k], (is3d ? n3 : n2) * 2, cudaMemcpyDeviceToHost));
    if (g.lut_tex[k]) {
      cudaDestroyTextureObject(g.lut_tex[k]);
      g.lut_tex[k] = 0;
    }
    if (!g.lut_arrays[k]) {
      if (is3d)
        RD_CHECK(cudaMalloc3DArray(&g.lut_arrays[k], &fmt, make_cudaExtent(BSDF_LUT_SIZE, BSDF_LUT_SIZE, BSDF_LUT_SIZE)));
      else
        RD_CHECK(cudaMallocArray(&g.lut_arrays[k], &fmt, BSDF_LUT_SIZE, BSDF_LUT_SIZE));
    }
    if (is3d) {
      cudaMemcpy3DParms cp;
      memset(&cp, 0, sizeof(cp));
      cp.srcPtr   = make_cudaPitchedPtr(d[k], BSDF_LUT_SIZE * 2, BSDF_LUT_SIZE, BSDF_LUT_SIZE);
      cp.dstArray = g.lut_arrays[k];
      cp.extent   = make_cudaExtent(BSDF_LUT_SIZE, BSDF_LUT_SIZE, BSDF_LUT_SIZE);
      cp.kind     = cudaMemcpyDeviceToDevice;
      RD_CHECK(cudaMemcpy3D(&cp));
    }
    else {
      RD_CHECK(cudaMemcpy2DToArray(g.lut_arrays[k], 0, 0, d[k], BSDF_LUT_SIZE * 2, BSDF_LUT_SIZE * 2, BSDF_LUT_SIZE, cudaMemcpyDeviceToDevice));
    }
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType         = cudaResourceTypeArray;
    rd.res.array.array = g.lut_arrays[k];
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode       = cudaFilterModeLinear;
    td.readMode         = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    RD_CHECK(cudaCreateTextureObject(&g.lut_tex[k], &rd, &td, nullptr));
    objs[k]->handle = (DeviceTextureHandle) g.lut_tex[k];
    objs[k]->gamma  = 1.0f;
    objs[k]->width  = BSDF_LUT_SIZE;
    objs[k]->height = BSDF_LUT_SIZE;
  }
  g.dirty = true;
  return 0;
}

/* sky_lut_generate + device_sky_lut_update (device_sky.c:80-220): the reference's two LUT kernels with its launch geometry
 * (kernel_execute_with_args: num_blocks x THREADS_PER_BLOCK; kernel_execute_custom: SKY_MS_ITER threads, SKY_MS_TEX_SIZE^2 blocks),
 * then float4 / linear / clamp / normalised pitch-2D textures over the same linear memory, as device_texture_create
 * (device_texture.c:255-330) configures them. device.sky must have been set (refdev_set_sky). Host copies: 256*64*4, 256*64*4,
 * 32*32*4, 32*32*4 floats. */
int refdev_build_sky_lut(float* tm_low, float* tm_high, float* ms_low, float* ms_high) {
  const size_t dims[4][2] = {{SKY_TM_TEX_WIDTH, SKY_TM_TEX_HEIGHT}, {SKY_TM_TEX_WIDTH, SKY_TM_TEX_HEIGHT}, {SKY_MS_TEX_SIZE, SKY_MS_TEX_SIZE},
                             {SKY_MS_TEX_SIZE, SKY_MS_TEX_SIZE}};
  const char* names[4]    = {"sky_tm_low", "sky_tm_high", "sky_ms_low", "sky_ms_high"};
  void* d[4];
  for (int k = 0; k < 4; k++)
    if (alloc_buffer(names[k], dims[k][0] * dims[k][1] * sizeof(float4), &d[k]))
      return 1;
  DeviceTextureObject objs[4];
  const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<float4>();
  for (int k = 0; k < 4; k++) {
    if (g.sky_tex[k]) {
      cudaDestroyTextureObject(g.sky_tex[k]);
      g.sky_tex[k] = 0;
    }
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType                  = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr       = d[k];
    rd.res.pitch2D.desc         = fmt;
    rd.res.pitch2D.width        = dims[k][0];
    rd.res.pitch2D.height       = dims[k][1];
    rd.res.pitch2D.pitchInBytes = dims[k][0] * sizeof(float4);
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode       = cudaFilterModeLinear;
    td.readMode         = cudaReadModeElementType;
    td.normalizedCoords = 1;
    RD_CHECK(cudaCreateTextureObject(&g.sky_tex[k], &rd, &td, nullptr));
    memset(&objs[k], 0, sizeof(objs[k]));
    objs[k].handle = (DeviceTextureHandle) g.sky_tex[k];
    objs[k].gamma  = 1.0f;
    objs[k].width  = (uint16_t) dims[k][0];
    objs[k].height = (uint16_t) dims[k][1];
  }
  const uint32_t saved     = g.host.config.num_blocks;
  const uint32_t blocks    = (uint32_t) ((SKY_TM_TEX_WIDTH * SKY_TM_TEX_HEIGHT + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK);
  g.host.config.num_blocks = blocks;
  g.dirty                  = true;
  if (sync_constant()) return 1;
  KernelArgsSkyComputeTransmittanceLUT a0;
  a0.dst_low  = (float4*) d[0];
  a0.dst_high = (float4*) d[1];
  sky_compute_transmittance_lut<<<blocks, THREADS_PER_BLOCK>>>(a0);
  if (finish("sky_compute_transmittance_lut")) return 1;
  KernelArgsSkyComputeMultiscatteringLUT a1;
  a1.transmission_low_tex  = objs[0];
  a1.transmission_high_tex = objs[1];
  a1.dst_low               = (float4*) d[2];
  a1.dst_high              = (float4*) d[3];
  sky_compute_multiscattering_lut<<<dim3(SKY_MS_TEX_SIZE, SKY_MS_TEX_SIZE, 1), dim3(SKY_MS_ITER, 1, 1)>>>(a1);
  if (finish("sky_compute_multiscattering_lut")) return 1;
  g.host.config.num_blocks = saved;
  float* host_out[4]       = {tm_low, tm_high, ms_low, ms_high};
  for (int k = 0; k < 4; k++)
    if (host_out[k])
      RD_CHECK(cudaMemcpy(host_out[k], d[k], dims[k][0] * dims[k][1] * sizeof(float4), cudaMemcpyDeviceToHost));
  g.host.sky_lut_transmission_low_tex     = objs[0];
  g.host.sky_lut_transmission_high_tex    = objs[1];
  g.host.sky_lut_multiscattering_low_tex  = objs[2];
  g.host.sky_lut_multiscattering_high_tex = objs[3];
  g.dirty                                 = true;
  return 0;
}

/* Replaces the contents of the four sky LUTs (built before by refdev_build_sky_lut) with tables supplied by the caller. */
int refdev_set_sky_lut(const float* tm_low, const float* tm_high, const float* ms_low, const float* ms_high) {
  const size_t texels[4] = {SKY_TM_TEX_WIDTH * SKY_TM_TEX_HEIGHT, SKY_TM_TEX_WIDTH * SKY_TM_TEX_HEIGHT, SKY_MS_TEX_SIZE * SKY_MS_TEX_SIZE,
                            SKY_MS_TEX_SIZE * SKY_MS_TEX_SIZE};
  const char* names[4]   = {"sky_tm_low", "sky_tm_high", "sky_ms_low", "sky_ms_high"};
  const float* src[4]    = {tm_low, tm_high, ms_low, ms_high};
  for (int k = 0; k < 4; k++) {
    auto it = g.buffers.find(names[k]);
    if (it == g.buffers.end()) return 2;
    RD_CHECK(cudaMemcpy(it->second.ptr, src[k], texels[k] * sizeof(float4), cudaMemcpyHostToDevice));
  }
  return 0;
}

/* sky_hdri_generate + _sky_hdri_compute (device_sky.c:279-375): the reference's sky_compute_hdri with its launch geometry
 * (num_pixels * 32 threads in blocks of THREADS_PER_BLOCK), then the colour table bound as device.sky_hdri_color_tex the way
 * sky_hdri_generate configures it (float4, point filter, wrap addressing, normalised coordinates). device.sky, the sky LUTs and
 * the blue-noise masks must have been set. Host copy: dim * dim * 4 floats. */
int refdev_build_sky_hdri(uint32_t dim, uint32_t sample_count, const float* origin, float* color) {
  void *dc, *ds;
  if (alloc_buffer("sky_hdri_color", (size_t) dim * dim * sizeof(float4), &dc)) return 1;
  if (alloc_buffer("sky_hdri_shadow", (size_t) dim * dim * sizeof(float), &ds)) return 1;
  const uint32_t saved      = g.host.config.num_blocks;
  const uint32_t num_blocks = (uint32_t) (((size_t) dim * dim * 32 + THREADS_PER_BLOCK - 1) / THREADS_PER_BLOCK);
  g.host.config.num_blocks  = num_blocks;
  memset(&g.host.state, 0, sizeof(g.host.state));
  g.dirty = true;
  if (sync_constant()) return 1;
  KernelArgsSkyComputeHDRI args;
  args.dst_color    = (float4*) dc;
  args.dst_shadow   = (float*) ds;
  args.dim          = dim;
  args.ld_color     = dim;
  args.ld_shadow    = dim;
  args.origin.x     = origin[0], args.origin.y = origin[1], args.origin.z = origin[2];
  args.sample_count = sample_count;
  sky_compute_hdri<<<num_blocks, THREADS_PER_BLOCK>>>(args);
  if (finish("sky_compute_hdri")) return 1;
  g.host.config.num_blocks = saved;
  if (color)
    RD_CHECK(cudaMemcpy(color, dc, (size_t) dim * dim * sizeof(float4), cudaMemcpyDeviceToHost));
  if (g.hdri_tex) {
    cudaDestroyTextureObject(g.hdri_tex);
    g.hdri_tex = 0;
  }
  cudaResourceDesc rd;
  memset(&rd, 0, sizeof(rd));
  rd.resType                  = cudaResourceTypePitch2D;
  rd.res.pitch2D.devPtr       = dc;
  rd.res.pitch2D.desc         = cudaCreateChannelDesc<float4>();
  rd.res.pitch2D.width        = dim;
  rd.res.pitch2D.height       = dim;
  rd.res.pitch2D.pitchInBytes = (size_t) dim * sizeof(float4);
  cudaTextureDesc td;
  memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
  td.filterMode       = cudaFilterModePoint;
  td.readMode         = cudaReadModeElementType;
  td.normalizedCoords = 1;
  RD_CHECK(cudaCreateTextureObject(&g.hdri_tex, &rd, &td, nullptr));
  memset(&g.host.sky_hdri_color_tex, 0, sizeof(g.host.sky_hdri_color_tex));
  g.host.sky_hdri_color_tex.handle = (DeviceTextureHandle) g.hdri_tex;
  g.host.sky_hdri_color_tex.gamma  = 1.0f;
  g.host.sky_hdri_color_tex.width  = (uint16_t) dim;
  g.host.sky_hdri_color_tex.height = (uint16_t) dim;
  g.dirty                          = true;
  return 0;
}

/* device_embedded_data_update (device_embedded_data.c:62-100): the moon's surface as png_load + device_texture_create deliver it -
 * RGBA8 unorm, wrap addressing, linear filter, normalised coordinates, gamma 1. Either pointer may be NULL (absent). */
int refdev_set_moon_textures(const uint8_t* albedo_rgba8, uint32_t aw, uint32_t ah, const uint8_t* normal_rgba8, uint32_t nw, uint32_t nh) {
  const uint8_t* src[2]      = {albedo_rgba8, normal_rgba8};
  const uint32_t dims[2][2]  = {{aw, ah}, {nw, nh}};
  DeviceTextureObject* dst[2] = {&g.host.moon_albedo_tex, &g.host.moon_normal_tex};
  for (int k = 0; k < 2; k++) {
    if (g.moon_tex[k]) cudaDestroyTextureObject(g.moon_tex[k]);
    if (g.moon_arrays[k]) cudaFreeArray(g.moon_arrays[k]);
    g.moon_tex[k] = 0, g.moon_arrays[k] = nullptr;
    memset(dst[k], 0, sizeof(DeviceTextureObject));
    dst[k]->handle = TEXTURE_OBJECT_INVALID;
    if (!src[k]) continue;
    const cudaChannelFormatDesc fmt = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
    RD_CHECK(cudaMallocArray(&g.moon_arrays[k], &fmt, dims[k][0], dims[k][1]));
    RD_CHECK(cudaMemcpy2DToArray(g.moon_arrays[k], 0, 0, src[k], dims[k][0] * 4, dims[k][0] * 4, dims[k][1], cudaMemcpyHostToDevice));
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof(rd));
    rd.resType         = cudaResourceTypeArray;
    rd.res.array.array = g.moon_arrays[k];
    cudaTextureDesc td;
    memset(&td, 0, sizeof(td));
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
    td.filterMode       = cudaFilterModeLinear;
    td.readMode         = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    RD_CHECK(cudaCreateTextureObject(&g.moon_tex[k], &rd, &td, nullptr));
    dst[k]->handle = (DeviceTextureHandle) g.moon_tex[k];
    dst[k]->gamma  = 1.0f;
    dst[k]->width  = (uint16_t) dims[k][0];
    dst[k]->height = (uint16_t) dims[k][1];
  }
  g.dirty = true;
  return 0;
}

/* device_sky_stars_update (device_sky.c:586-640): the catalogue and the STARS_GRID_LD x 32 + 1 cell offsets */
int refdev_set_stars(const float* stars, uint32_t count, const uint32_t* offsets) {
  void* p;
  if (upload_new("stars", stars, (size_t) count * sizeof(Star), &p)) return 1;
  g.host.ptrs.stars = (const Star*) p;
  if (upload_new("stars_offsets", offsets, sizeof(uint32_t) * (STARS_GRID_LD * 32 + 1), &p)) return 1;
  g.host.ptrs.stars_offsets = (const uint32_t*) p;
  g.dirty                   = true;
  return 0;
}

/* device_allocate_work_buffers (device_work_buffers.c:20-90) for a launch geometry chosen by the caller */
int refdev_configure(uint32_t num_blocks, uint32_t tasks_per_thread) {
  g.num_blocks       = num_blocks;
  g.tasks_per_thread = tasks_per_thread;
  const size_t T     = (size_t) num_blocks * THREADS_PER_BLOCK;
  const size_t K     = tasks_per_thread;
  void* p;
  if (alloc_buffer("task_states", sizeof(DeviceTaskState) * TASK_STATE_BUFFER_INDEX_COUNT * K * T, &p)) return 1;
  g.host.ptrs.task_states = (DeviceTaskState*) p;
  if (alloc_buffer("task_direct_light", sizeof(DeviceTaskDirectLight) * TASK_STATE_BUFFER_INDEX_DIRECT_LIGHT_COUNT * K * T, &p)) return 1;
  g.host.ptrs.task_direct_light = (DeviceTaskDirectLight*) p;
  if (alloc_buffer("task_results", sizeof(DeviceTaskResult) * TASK_STATE_BUFFER_INDEX_RESULT_COUNT * K * T, &p)) return 1;
  g.host.ptrs.task_results = (DeviceTaskResult*) p;
  if (alloc_buffer("results_counts", sizeof(uint16_t) * T, &p)) return 1;
  g.host.ptrs.results_counts = (uint16_t*) p;
  if (alloc_buffer("trace_counts", sizeof(uint16_t) * T, &p)) return 1;
  g.host.ptrs.trace_counts = (uint16_t*) p;
  if (alloc_buffer("task_counts", sizeof(uint16_t) * T * SHADING_TASK_INDEX_TOTAL, &p)) return 1;
  g.host.ptrs.task_counts = (uint16_t*) p;
  if (alloc_buffer("task_offsets", sizeof(uint16_t) * T * SHADING_TASK_INDEX_TOTAL, &p)) return 1;
  g.host.ptrs.task_offsets = (uint16_t*) p;
  const size_t px = (size_t) g.host.settings.width * g.host.settings.height;
  const char* names[10] = {"frame_first_moment_r", "frame_first_moment_g", "frame_first_moment_b", "frame_second_moment_luminance",
                           "frame_result_r",       "frame_result_g",       "frame_result_b",       "frame_output_r",
                           "frame_output_g",       "frame_output_b"};
  float* planes[10];
  for (int i = 0; i < 10; i++) {
    if (alloc_buffer(names[i], sizeof(float) * px, &p)) return 1;
    planes[i] = (float*) p;
  }
  for (int c = 0; c < 3; c++) {
    g.host.ptrs.frame_first_moment[c] = planes[c];
    g.host.ptrs.frame_result[c]       = planes[4 + c];
    g.host.ptrs.frame_output[c]       = planes[7 + c];
  }
  g.host.ptrs.frame_second_moment_luminance = planes[3];
  const size_t blocks                       = ((size_t) (g.host.settings.width + 3) / 4) * ((g.host.settings.height + 3) / 4) + 16;
  if (alloc_buffer("stage_sample_counts", sizeof(uint32_t) * blocks, &p)) return 1;
  g.host.ptrs.stage_sample_counts    = (uint32_t*) p;
  g.host.config.num_blocks           = num_blocks;
  g.host.config.num_tasks_per_thread = tasks_per_thread;
  g.dirty                            = true;
  return 0;
}

/* the per-launch execution state the reference's renderer writes (device_renderer.c:378-467): bounce depth, tile, and the
 * sample allocation of a non-adaptive pass (stage 0, one sample, first id = sample_id); accumulated_samples = number of
 * samples already in the planes BEFORE this pass (accumulation_generate_result divides by accumulated + 1). */
int refdev_set_state(uint32_t depth, uint32_t tile_id, uint32_t sample_id, uint32_t accumulated_samples) {
  memset(&g.host.state, 0, sizeof(g.host.state));
  g.host.state.depth                                       = (uint8_t) depth;
  g.host.state.tile_id                                     = tile_id;
  g.host.state.sample_allocation.stage_sample_offsets[0]   = sample_id;
  g.host.state.sample_allocation.upper_bound_tasks_per_sample = 1;
  g.host.state.sample_allocation.stage_id                  = 0;
  g.host.state.sample_allocation.num_samples               = 1;
  g.host.state.adaptive_sampling_accumulated_stages[0]     = accumulated_samples;
  g.dirty                                                  = true;
  return 0;
}

int refdev_upload(const char* name, size_t offset, const void* src, size_t bytes) {
  auto it = g.buffers.find(name);
  if (it == g.buffers.end() || offset + bytes > it->second.bytes) return 2;
  RD_CHECK(cudaMemcpy((uint8_t*) it->second.ptr + offset, src, bytes, cudaMemcpyHostToDevice));
  return 0;
}

int refdev_download(const char* name, size_t offset, void* dst, size_t bytes) {
  auto it = g.buffers.find(name);
  if (it == g.buffers.end() || offset + bytes > it->second.bytes) return 2;
  RD_CHECK(cudaMemcpy(dst, (const uint8_t*) it->second.ptr + offset, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

int refdev_clear(const char* name) {
  auto it = g.buffers.find(name);
  if (it == g.buffers.end()) return 2;
  RD_CHECK(cudaMemset(it->second.ptr, 0, it->second.bytes));
  return 0;
}

size_t refdev_buffer_size(const char* name) {
  auto it = g.buffers.find(name);
  return it == g.buffers.end() ? 0 : it->second.bytes;
}

#define RD_LAUNCH(kernel)                                 \
  do {                                                    \
    if (sync_constant()) return 1;                        \
    kernel<<<g.num_blocks, THREADS_PER_BLOCK>>>();        \
    return finish(#kernel);                               \
  } while (0)

int refdev_tasks_create(void) { RD_LAUNCH(tasks_create); }
int refdev_geometry_process_tasks(void) { RD_LAUNCH(geometry_process_tasks); }
int refdev_sky_process_tasks(void) { RD_LAUNCH(sky_process_tasks); }
int refdev_sky_process_inscattering_events(void) { RD_LAUNCH(sky_process_inscattering_events); }
int refdev_accumulation_collect_results(void) { RD_LAUNCH(accumulation_collect_results); }
int refdev_accumulation_collect_results_first_sample(void) { RD_LAUNCH(accumulation_collect_results_first_sample); }
int refdev_accumulation_generate_result(void) { RD_LAUNCH(accumulation_generate_result); }

/* Times `repeats` launches of geometry_process_tasks on the state currently uploaded (CUDA events around each launch on the
 * launch stream, one untimed launch first). The kernel appends its bounce tasks behind trace_counts[thread], so the counts
 * are zeroed before every launch (outside the timed interval); its other outputs are overwritten, the emission it adds to the
 * result records accumulates (irrelevant for timing). Used by tools/ref_shade_compare.py to report the reference's own shading
 * kernel, recompiled for sm_100a, beside the product's k_shade. */
int refdev_time_geometry_process_tasks(int repeats, float* avg_ms) {
  if (sync_constant()) return 1;
  auto tc = g.buffers.find("trace_counts");
  if (tc == g.buffers.end()) return 2;
  cudaEvent_t e0, e1;
  RD_CHECK(cudaEventCreate(&e0));
  RD_CHECK(cudaEventCreate(&e1));
  float total = 0.0f;
  for (int k = -1; k < repeats; k++) {
    RD_CHECK(cudaMemsetAsync(tc->second.ptr, 0, tc->second.bytes, 0));
    RD_CHECK(cudaEventRecord(e0, 0));
    geometry_process_tasks<<<g.num_blocks, THREADS_PER_BLOCK>>>();
    RD_CHECK(cudaEventRecord(e1, 0));
    if (finish("geometry_process_tasks (timed)")) return 1;
    float ms = 0.0f;
    RD_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (k >= 0)
      total += ms;
  }
  *avg_ms = total / (float) (repeats > 0 ? repeats : 1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

}  // extern "C"
