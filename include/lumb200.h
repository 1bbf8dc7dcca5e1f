/*
 * lumb200.h - C ABI of the B200-native wavefront path-tracing device path.
 *
 * This is the drop-in boundary for the per-bounce hot path of MilchRatchet/Luminary. Luminary's host and
 * device-manager layers talk to a GPU only through `device/device.h:141-198`; every entry point below
 * replaces one (or one group) of those functions and receives the same payload the reference passes
 * today (file:line of the replaced interface is cited per function). Plain pointers and sizes only; no
 * CUDA, torch or C++ types appear in a signature, so the library can be bound from C (the reference's
 * host language), ctypes, cgo, JNI, ...
 *
 * All functions return a Lumb200Result that uses Luminary's LuminaryResult numbering
 * (include/luminary/error.h:24-99): 0 = success, 1 = NULL argument, 3 = invalid argument, 7 = API
 * exception, 8 = CUDA error, 13 = invalid device. lumb200_last_error() returns a human readable string
 * for the calling thread's most recent failure.
 *
 * Threading: one Lumb200Device is driven by one thread at a time (Luminary: the device-manager worker,
 * device/device_manager.c:828-832). Work is queued on the device's own stream; calls that hand data
 * back to the host synchronise that stream themselves.
 */
#ifndef LUMB200_H
#define LUMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t Lumb200Result;

#define LUMB200_SUCCESS 0ull
#define LUMB200_ERROR_ARGUMENT_NULL 1ull
#define LUMB200_ERROR_NOT_IMPLEMENTED 2ull
#define LUMB200_ERROR_INVALID_API_ARGUMENT 3ull
#define LUMB200_ERROR_OUT_OF_MEMORY 5ull
#define LUMB200_ERROR_API_EXCEPTION 7ull
#define LUMB200_ERROR_CUDA 8ull
#define LUMB200_ERROR_MISSING_DATA 12ull
#define LUMB200_ERROR_INVALID_DEVICE 13ull

typedef struct Lumb200Device Lumb200Device;

/* Host-side triangle soup of one mesh: `Mesh` / `TriangleGeomData`, reference mesh.h:8-20. Non-indexed:
 * 9 floats of position and 9 floats of normal per triangle, 6 floats of uv, one material id. */
typedef struct Lumb200Mesh {
  uint32_t triangle_count;
  const float* vertex_buffer;
  const float* normal_buffer;
  const float* uv_buffer;
  const uint16_t* material_id_buffer;
} Lumb200Mesh;

/* `MeshInstance` as consumed by device_struct_instance_transform_convert, device_structs.c:402-413.
 * rotation is in Euler angles (radians), exactly what LuminaryInstance carries (structs.h:385-391). */
typedef struct Lumb200Instance {
  uint32_t mesh_id;
  float translation[3];
  float rotation[3];
  float scale[3];
  uint32_t active; /* inactive instances are skipped by the accel build */
} Lumb200Instance;

/* `LuminaryMaterial` (include/luminary/structs.h:360-381); packed on upload the way
 * device_struct_material_convert does (device_structs.c:270-330). Texture ids index the array handed to
 * lumb200_device_add_textures; 0xFFFF (LUMB200_TEXTURE_NONE) = no texture. */
#define LUMB200_TEXTURE_NONE 0xFFFFu
typedef struct Lumb200Material {
  uint32_t base_substrate; /* 0 opaque, 1 translucent */
  float albedo[4];
  float emission[3];
  float emission_scale;
  float roughness;
  float roughness_clamp;
  float refraction_index;
  uint8_t emission_active;
  uint8_t thin_walled;
  uint8_t metallic;
  uint8_t colored_transparency;
  uint8_t roughness_as_smoothness;
  uint8_t normal_map_is_compressed;
  uint8_t bidirectional_emission;
  uint8_t _pad;
  uint16_t albedo_tex;    /* rgb albedo + alpha (alpha == 0 texels are cut out of closest-hit and shadow rays) */
  uint16_t luminance_tex; /* emission colour, scaled by emission_scale */
  uint16_t roughness_tex; /* x channel */
  uint16_t metallic_tex;  /* carried, not evaluated (as in the reference, geometry_utils.cuh:160-162) */
  uint16_t normal_tex;    /* tangent-space normal map */
  uint16_t _pad2;
} Lumb200Material;

/* `Texture` (reference texture.h:21-40) as consumed by device_texture_create (device/device_texture.c), 2D only.
 * The path itself samples mip level 0 only (texture_get_default_args, cuda/texture_utils.cuh:13-22); the chain is generated
 * for textures that ask for it, as the reference does for scene textures (host/wavefront.c:267-268). */
enum { LUMB200_TEXTURE_FP32 = 0, LUMB200_TEXTURE_U8 = 1, LUMB200_TEXTURE_U16 = 2 };                              /* TextureDataType */
enum { LUMB200_WRAP_WRAP = 0, LUMB200_WRAP_CLAMP = 1, LUMB200_WRAP_MIRROR = 2, LUMB200_WRAP_BORDER = 3 };        /* TextureWrappingMode */
enum { LUMB200_FILTER_POINT = 0, LUMB200_FILTER_LINEAR = 1 };                                                    /* TextureFilterMode */
typedef struct Lumb200Texture {
  uint32_t width;
  uint32_t height;
  uint32_t pitch;          /* bytes per row of `data` */
  uint32_t type;           /* LUMB200_TEXTURE_* */
  uint32_t num_components; /* 1, 2 or 4 */
  uint32_t wrap_mode_u;
  uint32_t wrap_mode_v;
  uint32_t filter;
  float gamma;             /* applied to rgb on load, never to alpha (texture_utils.cuh:36-41) */
  uint32_t mipmap;         /* TextureMipmapMode: 0 none, 1 generate the mip chain on the device (4-component textures;
                              _device_texture_generate_mipmaps, device_texture.c:128-245 + cuda/mipmap.cuh) */
  const void* data;        /* HOST memory; NULL = invalid texture: loads return their default value */
} Lumb200Texture;

/* Subset of `LuminaryRendererSettings` the path consumes (structs.h:59-77). width/height are the
 * internal resolution (the reference shifts by `supersampling`, device_structs.c:21-22; callers do it). */
typedef struct Lumb200Settings {
  uint32_t width;
  uint32_t height;
  uint32_t max_ray_depth;
  uint32_t sort_by_material; /* 1: material-keyed compaction between trace and shade (default), 0: hit/miss only */
} Lumb200Settings;

/* Subset of `LuminaryCamera` (structs.h:160-211): thin-lens model. rotation in Euler angles. */
typedef struct Lumb200Camera {
  float pos[3];
  float rotation[3];
  float fov;
  float aperture_size;
  float object_distance;
  float camera_scale;
  float russian_roulette_threshold;
  uint32_t aperture_shape;
  uint32_t aperture_blade_count;
} Lumb200Camera;

/* `LuminarySky` (structs.h:262-292) as device_struct_sky_convert (device_structs.c:107-172), the LUT / star builders and the HDRI
 * bake (device_sky.c) read it. mode: 0 procedural atmosphere (LUMINARY_SKY_MODE_DEFAULT), 1 the atmosphere baked into a
 * latitude / longitude table (LUMINARY_SKY_MODE_HDRI: hdri_dim^2 texels, hdri_samples samples each, seen from the camera position at
 * bake time), 2 constant colour. lumb200_sky_default() fills the reference's defaults (sky.c:6-42) with mode = 2, which is this
 * library's initial state. aerial_perspective adds sky_process_inscattering_events (cuda/kernels.cuh:356-389) between the trace and
 * the sort of every bounce. Clouds are out of scope. */
typedef struct Lumb200Sky {
  uint32_t mode;
  float constant_color[3];
  float geometry_offset[3];
  float azimuth, altitude;           /* sun */
  float moon_azimuth, moon_altitude, moon_tex_offset;
  float sun_strength, base_density;
  float rayleigh_density, mie_density, ozone_density;
  float rayleigh_falloff, mie_falloff, mie_diameter, ground_visibility, ozone_layer_thickness, multiscattering_factor;
  float stars_intensity;
  uint32_t steps;                    /* ray-march steps of a miss, 1..1023 */
  uint32_t ozone_absorption;
  uint32_t aerial_perspective;
  uint32_t stars_count, stars_seed;  /* catalogue of _sky_stars_generate (device_sky.c:484-546: glibc rand()) */
  uint32_t hdri_dim, hdri_samples;   /* mode 1: table edge (1..8192) and samples per texel (>= 1) */
} Lumb200Sky;

/* `LightTree` as uploaded by device_update_light_tree_data (device_light.h:102-113): root blob =
 * DeviceLightTreeRootHeader + sections (device_utils.h:305-327), nodes = DeviceLightTreeNode[] (:283-303),
 * tri_handle_map = TriangleHandle[] (instance_id, tri_id) per light id. */
typedef struct Lumb200LightTree {
  const void* root_data;
  size_t root_size;
  const void* nodes_data;
  size_t nodes_size;
  const uint32_t* tri_handle_map;
  uint32_t num_lights;
} Lumb200LightTree;

/* Owned output of lumb200_host_build_light_tree; release with lumb200_host_free_light_tree. */
typedef struct Lumb200LightTreeBuffers {
  void* root_data;
  size_t root_size;
  void* nodes_data;
  size_t nodes_size;
  uint32_t* tri_handle_map;
  uint32_t num_lights;
} Lumb200LightTreeBuffers;

/* Output conversion parameters: the `LuminaryCamera` fields consumed by generate_final_image / convert_RGBF_to_ARGB8
 * (cuda/kernels.cuh:503-644, cuda/tonemap.cuh:175-246). exposure is the LINEAR scale expf(LuminaryCamera.exposure)
 * (device_structs.c:77); tonemap uses LuminaryToneMap numbering (0 none, 1 ACES, 2 Reinhard, 3 Uncharted 2, 4 AgX,
 * 5 AgX punchy, 6 AgX custom). */
typedef struct Lumb200OutputParams {
  float exposure;
  uint32_t tonemap;
  float agx_slope;
  float agx_power;
  float agx_saturation;
  uint32_t dithering; /* 1: add the 1D blue-noise mask before quantisation (needs lumb200_device_load_bluenoise_1d) */
  uint32_t purkinje;  /* 1: Purkinje shift of dark pixels before exposure (cuda/purkinje.cuh:19-90) */
  float purkinje_kappa1;
  float purkinje_kappa2;
  uint32_t supersampling; /* s: the frame was rendered at (w << s) x (h << s); every output pixel is the mean of the
                             (1 << s)^2 tone-mapped internal pixels (generate_final_image, kernels.cuh:503-560) */
  uint32_t local_error_minimization; /* LuminaryCamera.use_local_error_minimization: beauty pixels whose own standard error dominates
                                        are blended towards the mean of their 3 x 3 neighbourhood (accumulation.cuh:105-143) */
  float bloom_blend;      /* LuminaryCamera.bloom_blend; > 0: mip-chain bloom of the mean radiance before the tone map
                             (device_post_apply, device/device_post.c:62-140,210-231; cuda/post_common.cuh:71-143) */
  uint32_t filter;        /* LuminaryFilter: 0 none, 1 gray, 2 sepia, 3 gameboy, 4 2-bit gray, 5 CRT, 6 black & white; applied to the
                             tone-mapped output pixel before dithering (convert_RGBF_to_ARGB8, kernels.cuh:615-637; math.cuh:1081-1168).
                             Gameboy / 2-bit gray / black & white threshold against the 1D blue-noise mask (needs it loaded) */
  uint32_t use_color_correction; /* LuminaryCamera.use_color_correction: add color_correction to the pixel in HSV (tonemap.cuh:217-232) */
  float color_correction[3];     /* hue, saturation, value offsets */
  float film_grain;              /* LuminaryCamera.film_grain: white noise of this amplitude after exposure (tonemap.cuh:237-241) */
} Lumb200OutputParams;

/* Adaptive sampling (LuminaryRendererSettings.enable_adaptive_sampling & co., structs.h:59-77; device/device_adaptive_sampler.c,
 * cuda/adaptive_sampling.cuh). The image is tiled into 4 x 4 pixel blocks. Stage 0 renders one sample per pixel and execution;
 * after update_interval << s executions of stage s the samples per pixel of stage s + 1 (1 .. max_sampling_rate per execution)
 * are set per block from the variance estimate so that they average avg_sampling_rate. exposure_aware weighs the variance by
 * the squared compression of the tone map (exposure = linear exposure, tonemap / agx_* as in Lumb200OutputParams). */
#define LUMB200_ADAPTIVE_STAGES 4
typedef struct Lumb200AdaptiveSampling {
  uint32_t enable;
  uint32_t max_sampling_rate; /* clamped to 1 .. 256 (ADAPTIVE_SAMPLING_MAX_SAMPLING_RATE) */
  uint32_t avg_sampling_rate;
  uint32_t update_interval;
  uint32_t exposure_aware;
  float exposure;
  uint32_t tonemap;
  float agx_slope, agx_power, agx_saturation;
  uint32_t output_mode; /* LuminaryAdaptiveSamplingOutputMode: 0 beauty, 1 variance, 2 error, 3 sample distribution
                           (accumulation_generate_result, cuda/accumulation.cuh:86-190) */
} Lumb200AdaptiveSampling;

typedef struct Lumb200AdaptiveState {
  uint32_t stage_id;                                 /* 0 .. 4 */
  uint32_t executions[LUMB200_ADAPTIVE_STAGES + 1];  /* finished executions per stage */
  uint32_t tasks_per_execution;                      /* paths one execution of the current stage traces (upper bound incl. block padding) */
  uint32_t blocks_x, blocks_y;
  uint64_t paths_traced;                             /* since start_render */
} Lumb200AdaptiveState;

typedef struct Lumb200Stats {
  uint64_t closest_rays;   /* closest-hit rays traced since start_render */
  uint64_t shadow_rays;    /* transmittance shadow rays */
  uint64_t light_rays;     /* emitter-BVH enumeration rays */
  uint64_t kernel_launches;
  double render_seconds;   /* cumulative GPU seconds of the sample passes (device_renderer.c:593-652) */
  double accel_build_seconds;
  uint32_t samples_done;
  uint32_t bvh_nodes;
  uint32_t bvh_tris;
  uint32_t light_bvh_nodes;
  uint64_t device_bytes;
  uint32_t bvh_depth;       /* levels of the 8-wide scene BVH (the traversal stack is sized for 16) */
  uint32_t light_bvh_depth;
  float bvh_sah_cost;       /* SAH cost of the collapsed scene BVH, C(root) / A(root), c_node = 1 */
  uint32_t bvh_ploc_radius; /* PLOC search radius the build selected by that cost */
  uint64_t stack_overflows; /* traversal-stack entries that did not fit since start_render: MUST be 0 (rays would lose subtrees) */
  uint64_t nonfinite_samples; /* path samples dropped by the accumulation because their radiance was NaN / Inf, since start_render */
  uint32_t nonfinite_pixel;   /* pixel index (x + y * width) of the last such sample */
  uint32_t reserved0;
} Lumb200Stats;

/* Kernel classes of one sample pass, for lumb200_device_get_profile. */
enum {
  LUMB200_KERNEL_RAYGEN = 0,
  LUMB200_KERNEL_TRACE_CLOSEST = 1,
  LUMB200_KERNEL_SORT = 2,
  LUMB200_KERNEL_SHADE = 3,
  LUMB200_KERNEL_TRACE_SHADOW = 4,
  LUMB200_KERNEL_ACCUMULATE = 5,
  LUMB200_KERNEL_TRACE_ENUM = 6, /* emitter enumeration of the BSDF-sampled NEE direction + its evaluation */
  LUMB200_KERNEL_CLASS_COUNT = 8
};

/* Device time per kernel class, measured with CUDA events on the device's stream around every launch while
 * profiling is enabled (the reference's DEVICE_RENDERER_DO_PER_KERNEL_TIMING, device_renderer.h:10,46-63). */
typedef struct Lumb200Profile {
  double milliseconds[LUMB200_KERNEL_CLASS_COUNT];
  uint64_t launches[LUMB200_KERNEL_CLASS_COUNT];
} Lumb200Profile;

/* BVH work of one instrumented sample pass: nodes visited and triangles tested, summed over all rays. */
typedef struct Lumb200TraversalStats {
  uint64_t closest_rays, closest_nodes, closest_tris;
  uint64_t shadow_rays, shadow_nodes, shadow_tris;
  uint64_t light_rays;
  uint64_t shaded_vertices;     /* surface hits shaded by k_shade */
  uint64_t light_tree_nodes;    /* 64-byte light-tree nodes descended by the 8 reservoir lanes of those vertices */
  uint32_t light_root_sections; /* 48-byte root sections every vertex streams over (8 children each) */
  uint32_t _pad;
} Lumb200TraversalStats;

const char* lumb200_last_error(void);
Lumb200Result lumb200_get_device_count(uint32_t* count);
/* name and total memory of CUDA device `cuda_index` (luminary_host_get_device_info, device.c:218-300) */
Lumb200Result lumb200_get_device_properties(uint32_t cuda_index, char* name, size_t name_capacity, size_t* memory_bytes);

/* device_create / device_destroy, device/device.h:141,198 */
Lumb200Result lumb200_device_create(Lumb200Device** device, uint32_t cuda_index);
Lumb200Result lumb200_device_destroy(Lumb200Device** device);

/* device_load_embedded_data (device/device.h:151, device_embedded_data.c): blue-noise mask 256x256 u32 */
Lumb200Result lumb200_device_load_bluenoise(Lumb200Device* device, const uint32_t* bluenoise_2d, size_t count);

/* device_add_mesh, device/device.h:165 (device.c:899 -> device_mesh.c:19-51) */
Lumb200Result lumb200_device_add_mesh(Lumb200Device* device, const Lumb200Mesh* mesh, uint32_t* mesh_id);
/* device_update_instances, device/device.h:167 */
Lumb200Result lumb200_device_update_instances(Lumb200Device* device, const Lumb200Instance* instances, uint32_t count);
/* device_update_materials, device/device.h:166 */
Lumb200Result lumb200_device_update_materials(Lumb200Device* device, const Lumb200Material* materials, uint32_t count);
/* same, already in DeviceMaterialCompressed form (32 bytes each) */
Lumb200Result lumb200_device_update_materials_packed(Lumb200Device* device, const void* materials, uint32_t count);
/* device_add_textures, device/device.h:160 (-> device_texture_create, device/device_texture.c): APPENDS `count` textures
 * to the device's texture array, ids continue from the current count. The texel data is copied. */
Lumb200Result lumb200_device_add_textures(Lumb200Device* device, const Lumb200Texture* textures, uint32_t count);
/* Parity hook: raw tex2D<float4> fetches (no v flip, no gamma) of texture `texture_id` at `count` HOST (u, v) pairs;
 * rgba_out = 4 * count floats of HOST memory. */
Lumb200Result lumb200_device_sample_texture(Lumb200Device* device, uint32_t texture_id, const float* uv, uint32_t count, float* rgba_out);
/* same at an explicit mip level (tex2DLod; levels are point-selected, mipmapFilterMode POINT as in device_texture.c:269) */
Lumb200Result lumb200_device_sample_texture_lod(
  Lumb200Device* device, uint32_t texture_id, const float* uv, uint32_t count, float lod, float* rgba_out);
/* device_update_light_tree_data, device/device.h:171 */
Lumb200Result lumb200_device_update_light_tree(Lumb200Device* device, const Lumb200LightTree* tree);
/* Host-side (CPU, plain C) build of the light tree from the same scene description the device receives:
 * light_tree_build, device/device_light.c:2236 (called from device_manager.c:443). num_lights == 0 on return
 * means the scene has no emitters. */
Lumb200Result lumb200_host_build_light_tree(
  const Lumb200Mesh* meshes, uint32_t num_meshes, const Lumb200Instance* instances, uint32_t num_instances, const Lumb200Material* materials,
  uint32_t num_materials, Lumb200LightTreeBuffers* out);
/* Same, for scenes with luminance-textured emitters: triangle_intensities[mesh_id] is NULL or an array of triangle_count
 * floats holding the integrated texture intensity of each triangle as computed by lumb200_device_compute_light_intensities
 * (the reference's LightTreeCacheTriangle.average_intensity, device_light.c:2014, 2082-2094). Triangles of materials
 * without a luminance texture ignore the array. */
Lumb200Result lumb200_host_build_light_tree_textured(
  const Lumb200Mesh* meshes, uint32_t num_meshes, const Lumb200Instance* instances, uint32_t num_instances, const Lumb200Material* materials,
  uint32_t num_materials, const float* const* triangle_intensities, Lumb200LightTreeBuffers* out);
void lumb200_host_free_light_tree(Lumb200LightTreeBuffers* tree);
/* Replaces _light_tree_integrate + the light_compute_intensity kernel (device_light.c:1952-2018, cuda/light.cuh:190-262):
 * for `count` (mesh_id, triangle_id) pairs returns the largest colour importance of the triangle's luminance texture, scanned
 * texel by texel over 64 micro-triangles (one warp per triangle). Needs the meshes, materials and textures on the device.
 * mesh_ids / triangle_ids / intensities are HOST arrays. */
Lumb200Result lumb200_device_compute_light_intensities(
  Lumb200Device* device, const uint32_t* mesh_ids, const uint32_t* triangle_ids, uint32_t count, float* intensities);
/* device_update_scene_entity, device/device.h:159 (settings / camera / sky entities) */
Lumb200Result lumb200_device_update_settings(Lumb200Device* device, const Lumb200Settings* settings);
Lumb200Result lumb200_device_update_camera(Lumb200Device* device, const Lumb200Camera* camera);
/* LuminaryRendererSettings.shading_mode (structs.h LuminaryShadingMode; device_struct_settings_convert, device_structs.c:16): 0 the path
 * tracer; 1 albedo + emission, 2 depth, 3 shading normal, 4 identification colour of (instance, triangle), 5 lights - the one-bounce
 * queue of _device_renderer_build_debug_kernel_queue (device_renderer.c:136-182) with geometry_process_tasks_debug (cuda/geometry.cuh:182-246)
 * and sky_process_tasks_debug (cuda/sky.cuh:635-668). Under a debug mode the output chain skips the tone map and the bloom
 * (tonemap.cuh:207-208, device_post.c:214-215). */
Lumb200Result lumb200_device_set_shading_mode(Lumb200Device* device, uint32_t shading_mode);
Lumb200Result lumb200_device_update_sky(Lumb200Device* device, const Lumb200Sky* sky);
void lumb200_sky_default(Lumb200Sky* sky);
/* Sky LUTs of the procedural atmosphere (sky_lut_generate, device_sky.c:80-139), built by update_sky whenever a medium
 * parameter changed: transmittance 256 x 64 and multiscattering 32 x 32 texels, two float4 tables each (wavelengths 0-3 / 4-7).
 * HOST arrays of 256*64*4, 256*64*4, 32*32*4, 32*32*4 floats; any may be NULL. Fails under the constant-colour sky. */
Lumb200Result lumb200_device_get_sky_lut(
  Lumb200Device* device, float* transmittance_low, float* transmittance_high, float* multiscattering_low, float* multiscattering_high);
/* The star catalogue of the current sky: up to `capacity` stars of 4 floats (altitude, azimuth, radius, intensity) sorted by grid
 * cell, the 64 * 32 + 1 cell offsets, and the sun / moon positions in sky space (3 floats each). Any pointer may be NULL. */
Lumb200Result lumb200_device_get_sky_info(
  Lumb200Device* device, float* sun_pos, float* moon_pos, float* stars, uint32_t capacity, uint32_t* stars_offsets, uint32_t* stars_count);
/* The moon's surface textures of device_load_embedded_data (device_embedded_data.c:62-100: data/moon/moon_albedo.png and
 * moon_normal.png through png_load, i.e. RGBA8 / wrap / linear). Either may be NULL (absent: the moon is a black occluder, which is
 * also the state before this call). The texels are copied. */
Lumb200Result lumb200_device_load_moon_textures(Lumb200Device* device, const Lumb200Texture* albedo, const Lumb200Texture* normal);
/* device_build_sky_hdri (device.h; sky_hdri_generate, device_sky.c:323-375): bakes the sky as seen from the CURRENT camera position
 * (sky_compute_hdri, cuda/sky_hdri.cuh:60-158). Needs sky mode 1. start_render bakes it implicitly when the sky changed since the
 * last bake; a moved camera needs this call (luminary_host_request_sky_hdri_build). */
Lumb200Result lumb200_device_build_sky_hdri(Lumb200Device* device);
/* The baked table: dim * dim float4 texels (HOST array, `capacity_texels` of them; NULL to query *dim only) and its origin. */
Lumb200Result lumb200_device_get_sky_hdri(Lumb200Device* device, float* color, uint32_t capacity_texels, uint32_t* dim, float* origin);

/* device_build_bsdf_lut / device_update_bsdf_lut, device/device.h:172-173. get: 4 tables, R16:
 * conductor[32*32], glossy[32*32], dielectric[32^3], dielectric_inv[32^3]. */
Lumb200Result lumb200_device_build_bsdf_lut(Lumb200Device* device);
Lumb200Result lumb200_device_get_bsdf_lut(
  Lumb200Device* device, uint16_t* conductor, uint16_t* glossy, uint16_t* dielectric, uint16_t* dielectric_inv);
Lumb200Result lumb200_device_set_bsdf_lut(
  Lumb200Device* device, const uint16_t* conductor, const uint16_t* glossy, const uint16_t* dielectric, const uint16_t* dielectric_inv);

/* Replaces optix_bvh_gas_build / optix_bvh_ias_build / optix_bvh_light_build (device/optix_bvh.c:185-478):
 * flattens instances to world space and builds the compressed 8-wide BVHs on the device. */
Lumb200Result lumb200_device_build_accel(Lumb200Device* device);

/* device_start_render, device/device.h:182: clears the accumulation planes and counters. */
Lumb200Result lumb200_device_start_render(Lumb200Device* device);
/* device_continue_render, device/device.h:183: queues `count` full-frame sample passes with sample ids
 * first_sample_id + k * stride (k < count). Asynchronous. */
Lumb200Result lumb200_device_render_samples(Lumb200Device* device, uint32_t first_sample_id, uint32_t count, uint32_t stride);
Lumb200Result lumb200_device_sync(Lumb200Device* device);
/* device_update_scene_entity(SETTINGS) for the adaptive sampler (adaptive_sampler_setup, device_adaptive_sampler.c:29-56).
 * Takes effect at the next lumb200_device_start_render. */
Lumb200Result lumb200_device_update_adaptive_sampling(Lumb200Device* device, const Lumb200AdaptiveSampling* params);
/* device_continue_render under the adaptive sampler: queues `count` executions of the current schedule (sample ids follow from the
 * sampler state: a pixel's ids are consecutive over its history). A stage is built exactly when update_interval << stage
 * executions of it have finished (device_renderer.c:350-376 queues the build there; the reference finishes it asynchronously).
 * While adaptive sampling is enabled the outputs (download_result / download_output_argb8) divide every pixel by its own sample
 * count and ignore their sample_count argument. Single device: the stage build needs the combined planes. */
Lumb200Result lumb200_device_render_executions(Lumb200Device* device, uint32_t count);
Lumb200Result lumb200_device_get_adaptive_state(Lumb200Device* device, Lumb200AdaptiveState* state);
/* Shared-sampler mode for several devices / processes (the reference's AdaptiveSampler is shared by all devices:
 * adaptive_sampler_allocate_sample hands every execution its DeviceSampleAllocation, device_adaptive_sampler_update uploads the
 * shared stage sample counts, device_adaptive_sampler.c:58-71,330-420). The caller owns the schedule:
 *   set_adaptive_state     imposes stage id, executions per stage and (optionally, HOST array of blocks_x * blocks_y words) the
 *                          stage sample counts, e.g. the ones another device built;
 *   render_allocated_execution  renders ONE execution of the current stage whose allocation is `executions_before` (executions of
 *                          every stage finished globally before it - they fix the sample ids); the schedule is not advanced;
 *   build_adaptive_stage   builds the next stage from this device's planes (which must hold the COMBINED moments) and the
 *                          imposed execution totals, and advances the stage id. */
Lumb200Result lumb200_device_set_adaptive_state(
  Lumb200Device* device, uint32_t stage_id, const uint32_t executions[LUMB200_ADAPTIVE_STAGES + 1], const uint32_t* words, size_t num_words);
Lumb200Result lumb200_device_render_allocated_execution(Lumb200Device* device, const uint32_t executions_before[LUMB200_ADAPTIVE_STAGES + 1]);
Lumb200Result lumb200_device_build_adaptive_stage(Lumb200Device* device);
/* stage sample counts: one 32-bit word per block (blocks_x * blocks_y), to HOST memory */
Lumb200Result lumb200_device_download_adaptive_words(Lumb200Device* device, uint32_t* words, size_t count);

/* Accumulation planes [sum R | sum G | sum B | sum luminance(colour^2)], 4 * width * height floats
 * (device_utils.h:483-486). get: device pointer for a peer/NCCL reduce (device_result_interface.c);
 * bind: use caller-owned device memory instead (e.g. a torch tensor); download: D2H copy. */
Lumb200Result lumb200_device_get_frame_planes(Lumb200Device* device, void** device_ptr, size_t* num_floats);
Lumb200Result lumb200_device_bind_frame_planes(Lumb200Device* device, void* device_ptr, size_t num_floats);
Lumb200Result lumb200_device_download_frame_planes(Lumb200Device* device, float* dst, size_t num_floats);
/* accumulation_generate_result (cuda/accumulation.cuh:86-190, beauty mode): mean = sum / sample_count,
 * written to dst as 3 planes R,G,B of width*height floats (host memory). */
Lumb200Result lumb200_device_download_result(Lumb200Device* device, uint32_t sample_count, float* dst_rgb_planes);
/* Asynchronous flavour for progressive output (the reference's output callbacks run off stream_main too, device_output.c:160-233):
 * the frame is resolved on the render stream into one of two staging buffers and copied to dst (HOST memory, ideally pinned) on a
 * second stream, so the next sample passes overlap the transfer. `slot` is 0 or 1; lumb200_device_wait_download(slot) blocks until
 * that slot's copy has landed. A slot must be waited for before it is reused. */
Lumb200Result lumb200_device_download_result_async(Lumb200Device* device, uint32_t sample_count, float* dst_rgb_planes, uint32_t slot);
Lumb200Result lumb200_device_wait_download(Lumb200Device* device, uint32_t slot);
/* device output chain (device_output.c + generate_final_image + convert_RGBF_to_ARGB8, cuda/kernels.cuh:503-644):
 * mean -> bloom -> Purkinje shift -> exposure -> tone map -> supersampling box filter -> sRGB -> dither -> LuminaryARGB8
 * {b, g, r, a}; dst = (width >> s) * (height >> s) * 4 bytes of HOST memory, s = params->supersampling. */
Lumb200Result lumb200_device_load_bluenoise_1d(Lumb200Device* device, const uint16_t* bluenoise_1d, size_t count);
Lumb200Result lumb200_device_download_output_argb8(
  Lumb200Device* device, uint32_t sample_count, const Lumb200OutputParams* params, uint8_t* dst_argb8);
/* In-process multi-GPU combine (replaces device_result_interface.c:107-299, which stages through pinned host memory):
 * adds the accumulation planes of `other` (another CUDA device of this process) into `device` with a peer-to-peer
 * copy over NVLink plus one add kernel. Both devices must have finished their queued passes (the call synchronises). */
Lumb200Result lumb200_device_add_planes_from(Lumb200Device* device, Lumb200Device* other);
/* ---- the exchange step of the path: NCCL over NVLink / NVSwitch (csrc/comm.cu) ----
 * Replaces device_handle_result_sharing (device/device.c:1587-1612 -> device_result_interface.c:107-299: D2H to pinned host memory,
 * H2D, buffer_add; at most 4 devices): one in-place ncclReduce (sum) of the 4 accumulation planes of every member onto `root`,
 * queued on the members' render streams - it starts when the sample passes queued before it retire, nothing blocks the host.
 * The members that are not the root keep their partial planes; call lumb200_device_start_render on them before rendering the
 * next batch, as the reference does with its per-device staging (device_result_interface.c:214-262).
 *   create_all   one process, several devices (the device manager; ncclCommInitAll): comms[k] belongs to devices[k], rank k;
 *                collectives of such communicators go through the *_all entry points (one NCCL group).
 *   create_rank  one process per device: rank 0 obtains an id, the application distributes the LUMB200_COMM_ID_BYTES bytes.
 * NCCL is loaded at run time (libnccl.so.2); LUMB200_ERROR_MISSING_DATA when it is not installed. */
typedef struct Lumb200Comm Lumb200Comm;
#define LUMB200_COMM_ID_BYTES 128
Lumb200Result lumb200_comm_get_unique_id(void* id);
Lumb200Result lumb200_comm_create_rank(Lumb200Comm** comm, Lumb200Device* device, uint32_t world_size, uint32_t rank, const void* id);
Lumb200Result lumb200_comm_create_all(Lumb200Comm** comms, Lumb200Device* const* devices, uint32_t count);
Lumb200Result lumb200_comm_destroy(Lumb200Comm** comm);
Lumb200Result lumb200_comm_get_info(Lumb200Comm* comm, uint32_t* world_size, uint32_t* rank, uint32_t* nccl_version);
Lumb200Result lumb200_comm_reduce_planes(Lumb200Comm* comm, uint32_t root);
Lumb200Result lumb200_comm_reduce_planes_all(Lumb200Comm* const* comms, uint32_t count, uint32_t root);
/* Shared adaptive sampler (device_adaptive_sampler.c:330-420: the main device builds a stage, every device samples from the same
 * counts): broadcast of the per-block stage words of `root` into every member's own word buffer, on the render streams. */
Lumb200Result lumb200_comm_broadcast_adaptive_words(Lumb200Comm* comm, uint32_t root);
Lumb200Result lumb200_comm_broadcast_adaptive_words_all(Lumb200Comm* const* comms, uint32_t count, uint32_t root);
/* after the broadcast: a member that did not build the stage adopts it (stage id, the global executions per stage that fix its
 * sample ids; the task count comes from the broadcast prefix sums) */
Lumb200Result lumb200_device_adopt_adaptive_stage(Lumb200Device* device, uint32_t stage_id, const uint32_t executions[LUMB200_ADAPTIVE_STAGES + 1]);
/* zeroes the accumulation planes only (a member whose samples have been combined onto the root), keeps the sampler state */
Lumb200Result lumb200_device_clear_frame_planes(Lumb200Device* device);
/* CUDA ordinal of a device; device pointers + entry count of its adaptive stage counts / task prefix sums (for the collectives above) */
Lumb200Result lumb200_device_get_cuda_index(Lumb200Device* device, uint32_t* index);
Lumb200Result lumb200_device_get_adaptive_words_device(Lumb200Device* device, void** words, void** task_prefix, size_t* count);

/* device_get_gbuffer_meta stand-in / config-1 parity hook: traces the primary rays of one sample pass and
 * returns closest-hit handles to HOST arrays of width*height entries (any pointer may be NULL). */
Lumb200Result lumb200_device_trace_primary(
  Lumb200Device* device, uint32_t sample_id, uint32_t* instance_ids, uint32_t* tri_ids, float* t, float* u, float* v);
/* device_get_gbuffer_meta (device/device.h:190): closest hit of the primary ray of ONE pixel - instance id (0xFFFFFFFF on a miss),
 * triangle id, hit distance (FLT_MAX on a miss) and the ray direction (3 floats), for luminary_host_get_pixel_info. */
Lumb200Result lumb200_device_query_pixel(
  Lumb200Device* device, uint32_t x, uint32_t y, uint32_t sample_id, uint32_t* instance_id, uint32_t* tri_id, float* depth, float* ray);
/* Generic closest-hit batch on HOST ray arrays (3 floats each): H2D, trace, D2H. prim = flattened index. */
Lumb200Result lumb200_device_trace_rays(
  Lumb200Device* device, const float* origins, const float* directions, uint32_t count, uint32_t* instance_ids, uint32_t* tri_ids, float* t,
  float* u, float* v);

/* ---- per-vertex / per-ray parity hooks of the shading and shadow stages (tests; tools) ----
 * One path vertex as geometry_process_tasks (cuda/geometry.cuh:11-180) loads it: DeviceTask + DeviceTaskTrace +
 * DeviceTaskThroughput + DeviceTaskMediumStack (device_utils.h:365-397), with the hit given as a flattened primitive index. */
typedef struct Lumb200VertexIn {
  uint32_t pixel_x, pixel_y; /* PathID pixel (the sample id is the call's) */
  uint32_t state;            /* StateFlag bits, cuda/utils.cuh:113-120 */
  float origin[3];
  float ray[3];
  uint32_t prim;             /* flattened primitive index of the hit (instance prim offset + triangle id); 0xFFFFFFFF = miss:
                              * the vertex is shaded by the miss shader and `emission` returns the sky radiance x throughput */
  float t;                   /* hit distance */
  uint32_t record[2];        /* packed throughput */
  uint32_t medium;           /* packed IOR stack */
} Lumb200VertexIn;

/* One NEE shadow segment as queued for k_trace_shadow, plus what the shadow stage made of it. */
typedef struct Lumb200NeeSegment {
  uint32_t valid;
  float ray[3];
  float dist;          /* FLT_MAX for the ambient segment */
  float color[3];      /* unshadowed contribution x path throughput */
  uint32_t target_prim;
  float visible[3];    /* contribution that survived the transmittance test (color x visibility) */
} Lumb200NeeSegment;

/* Everything the shading + shadow stages write for one vertex: DeviceTaskDirectLight* evaluated (direct_lighting.cuh:445-669),
 * the emission added to the result record, and the bounce task (geometry.cuh:99-177). */
typedef struct Lumb200VertexOut {
  Lumb200NeeSegment nee[4]; /* 0 light-tree light, 1 BSDF-sampled light, 2 ambient, 3 sun */
  float emission[3];
  uint32_t alive;           /* the bounce task survived Russian roulette (and this is not the last iteration) */
  uint32_t state;
  float origin[3];
  float ray[3];
  uint32_t record[2];
  uint32_t medium;
} Lumb200VertexOut;

/* Runs the surface stages of ONE wavefront iteration (material sort -> shading -> shadow rays) on `count` caller-supplied
 * vertices instead of the output of the closest-hit stage, with the random numbers of (sample_id, rng_depth), and returns what
 * they wrote per vertex. HOST arrays; count <= width * height of the current settings. The accumulation planes are untouched. */
Lumb200Result lumb200_device_shade_vertices(
  Lumb200Device* device, uint32_t sample_id, uint32_t rng_depth, uint32_t is_last_iteration, const Lumb200VertexIn* vertices, uint32_t count,
  Lumb200VertexOut* out);
/* Runs k_trace_shadow on explicit segments: transmittance (3 floats per ray) with the reference's any-hit rules
 * (optix_anyhit.cuh:49-139): tmin = eps, hits with t < max_dist count, `ignore_prims[i]` (the surface the ray leaves) and
 * `target_prims[i]` (the emitter it aims at; 0xFFFFFFFF = none) are skipped. HOST arrays. */
Lumb200Result lumb200_device_trace_shadow_rays(
  Lumb200Device* device, const float* origins, const float* directions, const float* max_dist, const uint32_t* ignore_prims,
  const uint32_t* target_prims, uint32_t count, float* visibility);
/* Host-side packers of the upload path, exposed so that tests can compare them byte for byte with the reference's
 * (device_struct_material_convert device_structs.c:257-330 -> 32 bytes; device_struct_triangles_convert :332-386 ->
 * 3 x 16 bytes of DeviceTriangleVertex + 16 bytes of DeviceTriangleTexture per triangle; Quaternion16 + DeviceTransform
 * :388-413 -> 32 bytes). */
Lumb200Result lumb200_host_pack_material(const Lumb200Material* material, void* dst32);
Lumb200Result lumb200_host_pack_triangles(const Lumb200Mesh* mesh, void* vertices48, void* textris16);
Lumb200Result lumb200_host_pack_transform(const Lumb200Instance* instance, void* dst32);

/* Structural inspection of the acceleration structures (tests / debugging). which: 0 = scene BVH, 1 = emitter BVH.
 * Copies up to the given capacities of 80-byte nodes and 12-float triangles (v0.xyz, id bits, v1.xyz, 0, v2.xyz, 0)
 * to host memory and always reports the real counts. Either buffer may be NULL. */
Lumb200Result lumb200_device_download_bvh(
  Lumb200Device* device, uint32_t which, void* nodes, size_t node_capacity, float* triangles, size_t triangle_capacity, uint32_t* num_nodes,
  uint32_t* num_triangles);

Lumb200Result lumb200_device_get_stats(Lumb200Device* device, Lumb200Stats* stats);
/* Per-kernel-class timing. set: enables / disables event recording and clears the totals; get: synchronises. */
Lumb200Result lumb200_device_set_profiling(Lumb200Device* device, uint32_t enable);
Lumb200Result lumb200_device_get_profile(Lumb200Device* device, Lumb200Profile* profile);
/* Runs one sample pass with the instrumented traversal kernels (results are discarded, the planes untouched). */
Lumb200Result lumb200_device_measure_traversal(Lumb200Device* device, uint32_t sample_id, Lumb200TraversalStats* stats);
/* CUDA stream the device queues its work on (cudaStream_t as void*), so callers can time / order against it. */
Lumb200Result lumb200_device_get_stream(Lumb200Device* device, void** stream);
/* Time only the closest-hit kernel over the primary rays of `sample_id`, `repeats` times; returns average ms. */
Lumb200Result lumb200_device_time_primary_trace(Lumb200Device* device, uint32_t sample_id, uint32_t repeats, float* avg_ms);

#ifdef __cplusplus
}
#endif

#endif /* LUMB200_H */
