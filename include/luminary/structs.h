/*
 * luminary/structs.h - the value types of the API: settings, camera, sky, materials, instances, outputs (reference structs.h:26-391)
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/structs.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LUMINARY_API_STRUCTS_H
#define LUMINARY_API_STRUCTS_H

#include <luminary/api_utils.h>
#include <stddef.h>

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

/* ---- pixel queries (reference structs.h:104-110), answered from a primary-ray trace of the requested pixel -------------- */
typedef struct LuminaryPixelQueryResult {
  bool pixel_query_is_valid;
  uint32_t instance_id;
  uint16_t material_id;
  float depth;
  LuminaryVec3 rel_hit_pos;
} LuminaryPixelQueryResult;

/* ---- entities outside the path (reference structs.h:207-346): real layouts, so that applications that read or copy them compile
 *      and run; get returns the reference's defaults with active = false, set accepts anything that keeps the entity inactive and
 *      answers LUMINARY_ERROR_NOT_IMPLEMENTED when asked to activate it -------------------------------------------------------- */
typedef enum LuminaryJerlovWaterType {
  LUMINARY_JERLOV_WATER_TYPE_I   = 0,
  LUMINARY_JERLOV_WATER_TYPE_IA  = 1,
  LUMINARY_JERLOV_WATER_TYPE_IB  = 2,
  LUMINARY_JERLOV_WATER_TYPE_II  = 3,
  LUMINARY_JERLOV_WATER_TYPE_III = 4,
  LUMINARY_JERLOV_WATER_TYPE_1C  = 5,
  LUMINARY_JERLOV_WATER_TYPE_3C  = 6,
  LUMINARY_JERLOV_WATER_TYPE_5C  = 7,
  LUMINARY_JERLOV_WATER_TYPE_7C  = 8,
  LUMINARY_JERLOV_WATER_TYPE_9C  = 9,
  LUMINARY_JERLOV_WATER_TYPE_COUNT
} LuminaryJerlovWaterType;

typedef struct LuminaryOcean {
  bool active;
  float height;
  float amplitude;
  float frequency;
  float refractive_index;
  LuminaryJerlovWaterType water_type;
  bool caustics_active;
  uint32_t caustics_ris_sample_count;
  float caustics_domain_scale;
  bool multiscattering;
  bool triangle_light_contribution;
} LuminaryOcean;

typedef struct LuminaryCloudLayer {
  bool active;
  float height_max;
  float height_min;
  float coverage;
  float coverage_min;
  float type;
  float type_min;
  float wind_speed;
  float wind_angle;
} LuminaryCloudLayer;

typedef struct LuminaryCloud {
  bool active;
  bool initialized;
  bool atmosphere_scattering;
  LuminaryCloudLayer low;
  LuminaryCloudLayer mid;
  LuminaryCloudLayer top;
  float offset_x;
  float offset_z;
  float density;
  uint32_t seed;
  float droplet_diameter;
  uint32_t steps;
  uint32_t shadow_steps;
  float noise_shape_scale;
  float noise_detail_scale;
  float noise_weather_scale;
  float mipmap_bias;
  uint32_t octaves;
} LuminaryCloud;

typedef struct LuminaryFog {
  bool active;
  float density;
  float droplet_diameter;
  float height;
  float dist;
} LuminaryFog;

typedef struct LuminaryParticles {
  bool active;
  uint32_t seed;
  uint32_t count;
  LuminaryRGBF albedo;
  float speed;
  float direction_altitude;
  float direction_azimuth;
  float phase_diameter;
  float scale;
  float size;
  float size_variation;
} LuminaryParticles;

#endif /* LUMINARY_API_STRUCTS_H */
