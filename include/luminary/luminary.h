/*
 * luminary/luminary.h - the public C API of Luminary as served by the B200-native path (luminary_b200).
 *
 * Applications written against MilchRatchet/Luminary include <luminary/luminary.h> and call luminary_init,
 * luminary_host_* and luminary_path_*; this header declares the same names, argument meanings, result codes and
 * struct layouts for the part of that API that drives the path-tracing hot path (reference: include/luminary/host.h:29-129,
 * structs.h, error.h:24-101, path.h:26-28, luminary.h:45-50), so such an application re-links against
 * libluminary_b200.so and renders through the CUDA kernels of liblumb200.so without source changes.
 *
 * What is different, on purpose:
 *   - entities that are not on the path (ocean, clouds, fog, particles, pixel queries, sky HDRI) keep their
 *     entry points but are opaque here and answer LUMINARY_ERROR_NOT_IMPLEMENTED;
 *   - adaptive sampling is on by default as in the reference; its stages switch deterministically after update_interval << stage
 *     executions (the reference finishes the stage build asynchronously) and its executions run on the main device;
 *     undersampling (preview passes) is ignored with a warning;
 *   - the host renders until every requested output has been produced and then idles.
 */
#ifndef LUMINARY_B200_PUBLIC_API_H
#define LUMINARY_B200_PUBLIC_API_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUMINARY_API

/* ---- result codes (reference error.h:24-99) ---------------------------------------------------------------- */
typedef uint64_t LuminaryResult;

#define LUMINARY_SUCCESS (0ull)
#define LUMINARY_ERROR_ARGUMENT_NULL (1ull)
#define LUMINARY_ERROR_NOT_IMPLEMENTED (2ull)
#define LUMINARY_ERROR_INVALID_API_ARGUMENT (3ull)
#define LUMINARY_ERROR_MEMORY_LEAK (4ull)
#define LUMINARY_ERROR_OUT_OF_MEMORY (5ull)
#define LUMINARY_ERROR_C_STD (6ull)
#define LUMINARY_ERROR_API_EXCEPTION (7ull)
#define LUMINARY_ERROR_CUDA (8ull)
#define LUMINARY_ERROR_OPTIX (9ull)
#define LUMINARY_ERROR_PREVIOUS_ERROR (10ull)
#define LUMINARY_ERROR_DEBUG_ASSERT (11ull)
#define LUMINARY_ERROR_MISSING_DATA (12ull)
#define LUMINARY_ERROR_INVALID_DEVICE (13ull)
#define LUMINARY_ERROR_PROPAGATED (0x8000000000000000ull)

LUMINARY_API const char* luminary_result_to_string(LuminaryResult result);

/* ---- small value types (reference api_utils.h:26-51) ------------------------------------------------------- */
typedef struct LuminaryVec3 { float x, y, z; } LuminaryVec3;
typedef struct LuminaryRGBF { float r, g, b; } LuminaryRGBF;
typedef struct LuminaryRGBAF { float r, g, b, a; } LuminaryRGBAF;
typedef struct LuminaryARGB8 { uint8_t b, g, r, a; } LuminaryARGB8;

/* ---- library life time (reference luminary.h:45-50) -------------------------------------------------------- */
LUMINARY_API void luminary_init(void);
LUMINARY_API void luminary_shutdown(void);

/* ---- paths (reference path.h:23-28) ------------------------------------------------------------------------ */
typedef struct LuminaryPath LuminaryPath;
LUMINARY_API LuminaryResult luminary_path_create(LuminaryPath** path);
LUMINARY_API LuminaryResult luminary_path_set_from_string(LuminaryPath* path, const char* string);
LUMINARY_API LuminaryResult luminary_path_destroy(LuminaryPath** path);

/* ---- host creation (reference structs.h:26-34) ------------------------------------------------------------- */
#define LUMINARY_HOST_CREATE_INFO_DEVICE_MASK_ALL_DEVICES (0xFFFFFFFF)
typedef struct LuminaryHostCreateInfo {
  uint32_t device_mask; /* bit i enables CUDA device i; bits above the device count are ignored */
} LuminaryHostCreateInfo;

/* ---- renderer settings (reference structs.h:40-77) --------------------------------------------------------- */
typedef enum LuminaryShadingMode {
  LUMINARY_SHADING_MODE_DEFAULT = 0,
  LUMINARY_SHADING_MODE_ALBEDO = 1,
  LUMINARY_SHADING_MODE_DEPTH = 2,
  LUMINARY_SHADING_MODE_NORMAL = 3,
  LUMINARY_SHADING_MODE_IDENTIFICATION = 4,
  LUMINARY_SHADING_MODE_LIGHTS = 5,
  LUMINARY_SHADING_MODE_COUNT
} LuminaryShadingMode;

typedef enum LuminaryAdaptiveSamplingOutputMode {
  LUMINARY_ADAPTIVE_SAMPLING_OUTPUT_MODE_BEAUTY = 0,
  LUMINARY_ADAPTIVE_SAMPLING_OUTPUT_MODE_VARIANCE = 1,
  LUMINARY_ADAPTIVE_SAMPLING_OUTPUT_MODE_ERROR = 2,
  LUMINARY_ADAPTIVE_SAMPLING_OUTPUT_MODE_SAMPLE_DISTRIBUTION = 3,
  LUMINARY_ADAPTIVE_SAMPLING_OUTPUT_MODE_COUNT
} LuminaryAdaptiveSamplingOutputMode;

typedef struct LuminaryRendererSettings {
  uint32_t width;
  uint32_t height;
  uint32_t max_ray_depth;
  uint32_t bridge_max_num_vertices;
  uint32_t undersampling;
  uint32_t supersampling;
  bool enable_adaptive_sampling;
  uint32_t adaptive_sampling_max_sampling_rate;
  uint32_t adaptive_sampling_avg_sampling_rate;
  uint32_t adaptive_sampling_update_interval;
  bool adaptive_sampling_exposure_aware;
  LuminaryAdaptiveSamplingOutputMode adaptive_sampling_output_mode;
  LuminaryShadingMode shading_mode;
  float region_x;
  float region_y;
  float region_width;
  float region_height;
} LuminaryRendererSettings;

typedef struct LuminaryDeviceInfo {
  bool is_main_device;
  bool is_unavailable;
  bool is_enabled;
  char name[256];
  size_t memory_size;
  size_t allocated_memory_size;
} LuminaryDeviceInfo;

/* ---- outputs (reference structs.h:87-121) ------------------------------------------------------------------ */
typedef struct LuminaryOutputProperties {
  bool enabled;
  uint32_t width;
  uint32_t height;
} LuminaryOutputProperties;

#define LUMINARY_OUTPUT_HANDLE_INVALID 0xFFFFFFFF
typedef uint32_t LuminaryOutputHandle;
typedef uint32_t LuminaryOutputPromiseHandle;

typedef struct LuminaryOutputRequestProperties {
  uint32_t sample_count;
  uint32_t width;
  uint32_t height;
} LuminaryOutputRequestProperties;

typedef struct LuminaryImage {
  uint8_t* buffer; /* LuminaryARGB8 pixels, owned by the host until the output is released */
  uint32_t width;
  uint32_t height;
  size_t ld; /* row pitch in pixels */
  struct {
    float time; /* cumulative GPU seconds spent on the samples of this output */
    uint32_t sample_count;
  } meta_data;
} LuminaryImage;

/* ---- camera (reference structs.h:127-211) ------------------------------------------------------------------ */
typedef enum LuminaryFilter {
  LUMINARY_FILTER_NONE = 0,
  LUMINARY_FILTER_GRAY = 1,
  LUMINARY_FILTER_SEPIA = 2,
  LUMINARY_FILTER_GAMEBOY = 3,
  LUMINARY_FILTER_2BITGRAY = 4,
  LUMINARY_FILTER_CRT = 5,
  LUMINARY_FILTER_BLACKWHITE = 6,
  LUMINARY_FILTER_COUNT
} LuminaryFilter;

typedef enum LuminaryToneMap {
  LUMINARY_TONEMAP_NONE = 0,
  LUMINARY_TONEMAP_ACES = 1,
  LUMINARY_TONEMAP_REINHARD = 2,
  LUMINARY_TONEMAP_UNCHARTED2 = 3,
  LUMINARY_TONEMAP_AGX = 4,
  LUMINARY_TONEMAP_AGX_PUNCHY = 5,
  LUMINARY_TONEMAP_AGX_CUSTOM = 6,
  LUMINARY_TONEMAP_COUNT
} LuminaryToneMap;

typedef enum LuminaryApertureShape { LUMINARY_APERTURE_ROUND = 0, LUMINARY_APERTURE_BLADED = 1, LUMINARY_APERTURE_COUNT } LuminaryApertureShape;

typedef struct LuminaryCamera {
  LuminaryVec3 pos;
  LuminaryVec3 rotation; /* Euler angles, radians */
  LuminaryApertureShape aperture_shape;
  uint32_t aperture_blade_count;
  float exposure; /* exponential scale: the image is multiplied by expf(exposure) */
  LuminaryToneMap tonemap;
  float agx_custom_slope;
  float agx_custom_power;
  float agx_custom_saturation;
  LuminaryFilter filter;
  bool use_local_error_minimization;
  float bloom_blend;
  bool dithering;
  bool purkinje;
  float purkinje_kappa1;
  float purkinje_kappa2;
  float wasd_speed;
  float mouse_speed;
  bool smooth_movement;
  float smoothing_factor;
  float russian_roulette_threshold;
  bool use_color_correction;
  LuminaryRGBF color_correction;
  float film_grain;
  float camera_scale;
  float object_distance;
  bool use_physical_camera;
  struct {
    float fov;
    float aperture_size;
  } thin_lens;
  struct {
    bool allow_reflections;
    bool use_spectral_rendering;
    float focal_length;
    float front_focal_point;
    float back_focal_point;
    float front_principal_point;
    float back_principal_point;
    float aperture_point;
    float aperture_diameter;
    float exit_pupil_point;
    float exit_pupil_diameter;
    float image_plane_distance;
    float sensor_width;
  } physical;
} LuminaryCamera;

/* ---- sky (reference structs.h:253-292); only mode and constant_color reach the path ------------------------- */
typedef enum LuminarySkyMode {
  LUMINARY_SKY_MODE_DEFAULT = 0,
  LUMINARY_SKY_MODE_HDRI = 1,
  LUMINARY_SKY_MODE_CONSTANT_COLOR = 2,
  LUMINARY_SKY_MODE_COUNT
} LuminarySkyMode;

typedef struct LuminarySky {
  LuminaryVec3 geometry_offset;
  float azimuth;
  float altitude;
  float moon_azimuth;
  float moon_altitude;
  float moon_tex_offset;
  float sun_strength;
  float base_density;
  bool ozone_absorption;
  uint32_t steps;
  uint32_t stars_count;
  uint32_t stars_seed;
  float stars_intensity;
  float rayleigh_density;
  float mie_density;
  float ozone_density;
  float rayleigh_falloff;
  float mie_falloff;
  float mie_diameter;
  float ground_visibility;
  float ozone_layer_thickness;
  float multiscattering_factor;
  uint32_t hdri_dim;
  uint32_t hdri_samples;
  bool aerial_perspective;
  LuminaryRGBF constant_color;
  LuminarySkyMode mode;
} LuminarySky;

/* ---- materials and instances (reference structs.h:352-391) ------------------------------------------------- */
typedef enum LuminaryMaterialBaseSubstrate {
  LUMINARY_MATERIAL_BASE_SUBSTRATE_OPAQUE,
  LUMINARY_MATERIAL_BASE_SUBSTRATE_TRANSLUCENT,
  LUMINARY_MATERIAL_BASE_SUBSTRATE_COUNT
} LuminaryMaterialBaseSubstrate;

typedef struct LuminaryMaterial {
  uint32_t id;
  LuminaryMaterialBaseSubstrate base_substrate;
  LuminaryRGBAF albedo;
  LuminaryRGBF emission;
  float emission_scale;
  float roughness;
  float roughness_clamp;
  float refraction_index;
  bool emission_active;
  bool thin_walled;
  bool metallic;
  bool colored_transparency;
  bool roughness_as_smoothness;
  bool normal_map_is_compressed;
  bool bidirectional_emission;
  uint16_t albedo_tex; /* 0xFFFF = none; ids index the textures in load order of the *.obj files (map_* statements) */
  uint16_t luminance_tex;
  uint16_t roughness_tex;
  uint16_t metallic_tex;
  uint16_t normal_tex;
} LuminaryMaterial;

typedef struct LuminaryInstance {
  uint32_t id;
  uint32_t mesh_id;
  LuminaryVec3 position;
  LuminaryVec3 rotation;
  LuminaryVec3 scale;
} LuminaryInstance;

/* entities outside the path: opaque, their accessors answer LUMINARY_ERROR_NOT_IMPLEMENTED */
typedef struct LuminaryOcean LuminaryOcean;
typedef struct LuminaryCloud LuminaryCloud;
typedef struct LuminaryFog LuminaryFog;
typedef struct LuminaryParticles LuminaryParticles;
typedef struct LuminaryPixelQueryResult LuminaryPixelQueryResult;

/* ---- host (reference host.h:29-129) ------------------------------------------------------------------------ */
typedef struct LuminaryHost LuminaryHost;

LUMINARY_API LuminaryResult luminary_host_create(LuminaryHost** host, LuminaryHostCreateInfo info);
LUMINARY_API LuminaryResult luminary_host_destroy(LuminaryHost** host);

LUMINARY_API LuminaryResult luminary_host_start_new_render(LuminaryHost* host);

LUMINARY_API LuminaryResult luminary_host_get_device_count(LuminaryHost* host, uint32_t* device_count);
LUMINARY_API LuminaryResult luminary_host_get_device_info(LuminaryHost* host, uint32_t device_id, LuminaryDeviceInfo* info);
LUMINARY_API LuminaryResult luminary_host_set_device_enable(LuminaryHost* host, uint32_t device_id, bool enable);

LUMINARY_API LuminaryResult luminary_host_load_lum_file(LuminaryHost* host, LuminaryPath* path);
LUMINARY_API LuminaryResult luminary_host_load_obj_file(LuminaryHost* host, LuminaryPath* path);

LUMINARY_API LuminaryResult luminary_host_get_current_sample_time(LuminaryHost* host, double* time);

LUMINARY_API LuminaryResult luminary_host_get_num_queue_workers(const LuminaryHost* host, uint32_t* num_queue_workers);
LUMINARY_API LuminaryResult luminary_host_get_queue_worker_name(const LuminaryHost* host, uint32_t queue_worker_id, const char** string);
LUMINARY_API LuminaryResult luminary_host_get_queue_worker_string(const LuminaryHost* host, uint32_t queue_worker_id, const char** string);
LUMINARY_API LuminaryResult luminary_host_get_queue_worker_time(const LuminaryHost* host, uint32_t queue_worker_id, double* time);

LUMINARY_API LuminaryResult luminary_host_set_output_properties(LuminaryHost* host, LuminaryOutputProperties properties);
LUMINARY_API LuminaryResult
  luminary_host_request_output(LuminaryHost* host, LuminaryOutputRequestProperties properties, LuminaryOutputPromiseHandle* handle);
/* writes LUMINARY_OUTPUT_HANDLE_INVALID while the requested sample count has not been reached */
LUMINARY_API LuminaryResult
  luminary_host_try_await_output(LuminaryHost* host, LuminaryOutputPromiseHandle handle, LuminaryOutputHandle* output_handle);
/* most recent finished output; every acquired handle must be released */
LUMINARY_API LuminaryResult luminary_host_acquire_output(LuminaryHost* host, LuminaryOutputHandle* output_handle);
LUMINARY_API LuminaryResult luminary_host_get_image(LuminaryHost* host, LuminaryOutputHandle output_handle, LuminaryImage* image);
LUMINARY_API LuminaryResult luminary_host_release_output(LuminaryHost* host, LuminaryOutputHandle output_handle);

LUMINARY_API LuminaryResult luminary_host_get_pixel_info(LuminaryHost* host, uint16_t x, uint16_t y, LuminaryPixelQueryResult* result);

LUMINARY_API LuminaryResult luminary_host_get_settings(LuminaryHost* host, LuminaryRendererSettings* settings);
LUMINARY_API LuminaryResult luminary_host_set_settings(LuminaryHost* host, const LuminaryRendererSettings* settings);
LUMINARY_API LuminaryResult luminary_host_get_camera(LuminaryHost* host, LuminaryCamera* camera);
LUMINARY_API LuminaryResult luminary_host_set_camera(LuminaryHost* host, const LuminaryCamera* camera);
LUMINARY_API LuminaryResult luminary_host_get_sky(LuminaryHost* host, LuminarySky* sky);
LUMINARY_API LuminaryResult luminary_host_set_sky(LuminaryHost* host, const LuminarySky* sky);
LUMINARY_API LuminaryResult luminary_host_get_ocean(LuminaryHost* host, LuminaryOcean* ocean);
LUMINARY_API LuminaryResult luminary_host_set_ocean(LuminaryHost* host, const LuminaryOcean* ocean);
LUMINARY_API LuminaryResult luminary_host_get_cloud(LuminaryHost* host, LuminaryCloud* cloud);
LUMINARY_API LuminaryResult luminary_host_set_cloud(LuminaryHost* host, const LuminaryCloud* cloud);
LUMINARY_API LuminaryResult luminary_host_get_fog(LuminaryHost* host, LuminaryFog* fog);
LUMINARY_API LuminaryResult luminary_host_set_fog(LuminaryHost* host, const LuminaryFog* fog);
LUMINARY_API LuminaryResult luminary_host_get_particles(LuminaryHost* host, LuminaryParticles* particles);
LUMINARY_API LuminaryResult luminary_host_set_particles(LuminaryHost* host, const LuminaryParticles* particles);

LUMINARY_API LuminaryResult luminary_host_get_material(LuminaryHost* host, uint16_t id, LuminaryMaterial* material);
LUMINARY_API LuminaryResult luminary_host_set_material(LuminaryHost* host, uint16_t id, const LuminaryMaterial* material);
LUMINARY_API LuminaryResult luminary_host_get_instance(LuminaryHost* host, uint32_t id, LuminaryInstance* instance);
LUMINARY_API LuminaryResult luminary_host_set_instance(LuminaryHost* host, const LuminaryInstance* instance);
LUMINARY_API LuminaryResult luminary_host_new_instance(LuminaryHost* host, LuminaryInstance* instance);

LUMINARY_API LuminaryResult luminary_host_get_num_meshes(LuminaryHost* host, uint32_t* num_meshes);
LUMINARY_API LuminaryResult luminary_host_get_num_materials(LuminaryHost* host, uint32_t* num_materials);
LUMINARY_API LuminaryResult luminary_host_get_num_instances(LuminaryHost* host, uint32_t* num_instances);

LUMINARY_API LuminaryResult luminary_host_save_png(LuminaryHost* host, LuminaryOutputHandle handle, LuminaryPath* path);
LUMINARY_API LuminaryResult luminary_host_request_sky_hdri_build(LuminaryHost* host);

/* ---- additions of this implementation (not in the reference) ------------------------------------------------ */
/* human readable text of the calling thread's most recent failure inside this library */
LUMINARY_API const char* luminary_b200_last_error(void);
/* rays traced (closest-hit + shadow + emitter enumeration) by all devices since the last start_new_render */
LUMINARY_API LuminaryResult luminary_b200_host_get_ray_count(LuminaryHost* host, uint64_t* rays);
/* blocks until the render worker has nothing left to do (all requested outputs produced or an error occurred) */
LUMINARY_API LuminaryResult luminary_b200_host_wait_idle(LuminaryHost* host);

#ifdef __cplusplus
}
#endif

#endif /* LUMINARY_B200_PUBLIC_API_H */
