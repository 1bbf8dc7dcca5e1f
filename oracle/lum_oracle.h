/*
 * lum_oracle.h - CPU oracle for the per-bounce path-tracing hot path.
 *
 * TEST INFRASTRUCTURE ONLY. This library is a plain-C restatement of the
 * reference renderer's device functions (MilchRatchet/Luminary, file:line cited
 * at each function). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it. The product
 * (luminary_b200/) never links, imports or calls anything in this directory.
 *
 * Parity pinning: the reference ships no tests, golden vectors or CPU render
 * path, and its closest-hit arithmetic lives in the
 * closed-source OptiX driver. Integer paths (PathID, Squares/Sobol/blue-noise
 * RNG, packing) are pinned by known-answer vectors derived by hand from the
 * published algorithms and by the reference's own data file (bluenoise_2D.bin);
 * floating-point shading is "parity unpinned" against a running reference and
 * is checked statistically (tests state the tolerance).
 *
 * All arithmetic is IEEE-754 binary32, compiled with -ffp-contract=off so that
 * no multiply-add is fused unless fmaf() is written out. The CUDA product's
 * geometry kernels are compiled with -fmad=false for the same reason; this is
 * what makes closest-hit ids bit-exact between the two.
 */
#ifndef LUM_ORACLE_H
#define LUM_ORACLE_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, z; } OrcVec3;
typedef struct { float r, g, b; } OrcRGB;
typedef struct { float x, y, z, w; } OrcQuat;
typedef struct { uint16_t x, y, z, w; } OrcQuat16;
typedef struct { uint16_t x, y, z; } OrcPathID;
typedef struct { uint32_t x, y; } OrcUint2;
typedef struct { float x, y; } OrcFloat2;

#define ORC_PI 3.141592653589f
#define ORC_EPS 1.1920929e-07f /* FLT_EPSILON, the reference's `eps` (cuda/utils.cuh:41-43) */
#define ORC_FLT_MAX 3.402823466e+38f

#define ORC_HIT_SKY 0xFFFFFFFEu     /* HIT_TYPE_SKY, cuda/utils.cuh:52 */
#define ORC_HIT_INVALID 0xFFFFFFFFu /* HIT_TYPE_INVALID */
#define ORC_LIGHT_ID_INVALID 0xFFFFFFFFu

/* StateFlag, cuda/utils.cuh:113-120 */
enum {
  ORC_STATE_DELTA_PATH        = 0x01,
  ORC_STATE_CAMERA_DIRECTION  = 0x02,
  ORC_STATE_VOLUME_SCATTERED  = 0x04,
  ORC_STATE_ALLOW_EMISSION    = 0x08,
  ORC_STATE_ALLOW_AMBIENT     = 0x10,
  ORC_STATE_USE_IGNORE_HANDLE = 0x20
};

/* ---- random targets: enum RandomTarget, cuda/random.cuh:24-66 (each allocation takes size*sets+1 enumerators; values checked by compiling the macro) ---- */
enum {
  ORC_RT_LENS                    = 33,
  ORC_RT_LENS_BLADE              = 35,
  ORC_RT_BSDF_REFLECTION         = 39,  /* + set id (3 sets) */
  ORC_RT_BSDF_DIFFUSE            = 43,
  ORC_RT_BSDF_REFRACTION         = 47,
  ORC_RT_BSDF_RESAMPLING         = 51,
  ORC_RT_BSDF_OPACITY            = 55,
  ORC_RT_RUSSIAN_ROULETTE        = 61,
  ORC_RT_CAMERA_JITTER           = 63,
  ORC_RT_CAMERA_TIME             = 65,
  ORC_RT_SKY_STEP_OFFSET         = 77,
  ORC_RT_SKY_INSCATTERING_STEP   = 79,
  ORC_RT_LIGHT_SUN_BSDF          = 346, /* + set id (2 sets; the surface uses set 0) */
  ORC_RT_LIGHT_SUN_BSDF_METHOD   = 349,
  ORC_RT_LIGHT_SUN_RAY           = 352,
  ORC_RT_LIGHT_SUN_RESAMPLING    = 355,
  ORC_RT_LIGHT_GEO_RAY           = 367, /* 8 lanes, + 8 * set id */
  ORC_RT_LIGHT_GEO_RESAMPLING    = 384,
  ORC_RT_LIGHT_GEO_TREE_PREPASS  = 387, /* 8 lanes */
  ORC_RT_LIGHT_GEO_TREE_POSTPASS = 404, /* 8 lanes */
  ORC_RT_LIGHT_BSDF_CHOICE       = 569,
  ORC_RT_LIGHT_BSDF_DIRECTION    = 571,
  ORC_RT_LIGHT_BSDF_TRACE        = 573,
  ORC_RT_LIGHT_BSDF_RR           = 575,
  ORC_RT_COUNT                   = 577
};

/* ---- scene description (host SoA as handed to device_add_mesh, reference mesh.h:8-20) ---- */
typedef struct {
  uint32_t num_tris;
  const float* vertex;      /* 9 floats per triangle (v0,v1,v2) */
  const float* normal;      /* 9 floats per triangle */
  const float* uv;          /* 6 floats per triangle */
  const uint16_t* material; /* 1 per triangle */
} OrcMesh;

/* DeviceTransform, device_structs.h:292-297 */
typedef struct {
  OrcVec3 translation;
  OrcVec3 scale;
  OrcQuat16 rotation;
} OrcTransform;

typedef struct {
  uint32_t mesh_id;
  OrcTransform transform;
} OrcInstance;

/* DeviceMaterialCompressed, device_structs.h:232-254 (32 bytes) */
typedef struct {
  uint8_t flags;
  uint8_t roughness_clamp;
  uint16_t metallic_tex;
  uint16_t roughness;
  uint16_t refraction_index;
  uint16_t albedo_r, albedo_g, albedo_b, albedo_a;
  uint16_t emission_r, emission_g, emission_b, emission_scale;
  uint16_t albedo_tex, luminance_tex, roughness_tex, normal_tex;
} OrcMaterialPacked;

/* LuminaryMaterial subset, include/luminary/structs.h:360-381 */
typedef struct {
  uint32_t base_substrate; /* 0 opaque, 1 translucent */
  float albedo[4];
  float emission[3];
  float emission_scale;
  float roughness;
  float roughness_clamp;
  float refraction_index;
  bool emission_active, thin_walled, metallic, colored_transparency, roughness_as_smoothness, normal_map_is_compressed,
    bidirectional_emission;
} OrcMaterialDesc;

/* camera: DeviceCamera thin-lens subset, device_structs.h:38-83 */
typedef struct {
  OrcVec3 pos;
  OrcQuat rotation;
  float fov;             /* half width of the sensor at z = 1 */
  float aperture_size;   /* 0 => pinhole */
  float object_distance;
  float camera_scale;
  float russian_roulette_threshold;
  uint32_t aperture_shape; /* 0 round, 1 bladed */
  uint32_t aperture_blade_count;
} OrcCamera;

typedef struct {
  uint32_t width, height;
  uint32_t max_ray_depth;
  uint32_t sky_mode;       /* 0 default (procedural when the scene carries a sky, orc_scene_set_sky; else black), 1 HDRI (baked table of
                            * orc_scene_build_sky_hdri; marches like 0 until it is built), 2 constant colour */
  OrcRGB sky_constant_color;
} OrcSettings;

/* ------------------------------------------------------------------ */
/* orc_core.c : integer paths, RNG, packing, camera                    */
/* ------------------------------------------------------------------ */
OrcPathID orc_path_id_get(uint32_t x, uint32_t y, uint32_t sample_id);
void orc_path_id_pixel(OrcPathID id, uint32_t* x, uint32_t* y);
uint32_t orc_path_id_sample(OrcPathID id);

uint32_t orc_squares32(uint32_t key, uint32_t counter);
uint16_t orc_squares16(uint32_t key, uint32_t counter);
OrcUint2 orc_sobol(uint32_t offset, uint32_t dimension);
void orc_set_bluenoise(const uint32_t* table_256x256);
OrcUint2 orc_random_2d_base(uint32_t target, uint32_t px, uint32_t py, uint32_t sequence_id, uint32_t depth);
float orc_u32_to_float(uint32_t v);
float orc_u16_to_float(uint16_t v);
OrcFloat2 orc_random_2d(uint32_t target, OrcPathID id, uint32_t depth);
float orc_random_1d(uint32_t target, OrcPathID id, uint32_t depth);
float orc_random_saturate(float r);

uint32_t orc_pack_normal_host(OrcVec3 n);   /* device_packing.c:6-31 (double) */
uint32_t orc_pack_normal(OrcVec3 n);        /* cuda/math.cuh:1732-1758 (float)  */
OrcVec3 orc_unpack_normal(uint32_t p);      /* cuda/math.cuh:1716-1730 */
uint32_t orc_pack_uv(float u, float v);     /* device_packing.c:36-43 */
OrcFloat2 orc_unpack_uv(uint32_t p);        /* cuda/math.cuh:1706-1713 */
OrcUint2 orc_record_pack(OrcRGB c);         /* cuda/math.cuh:1609-1619 */
OrcRGB orc_record_unpack(OrcUint2 p);       /* cuda/math.cuh:1595-1607 */
OrcUint2 orc_ray_pack(OrcVec3 ray);         /* cuda/math.cuh:1637-1664 */
OrcVec3 orc_ray_unpack(OrcUint2 p);         /* cuda/math.cuh:1621-1635 */
uint32_t orc_ior_compress(float ior);       /* cuda/math.cuh:1762-1764 */
float orc_ior_decompress(uint32_t c);       /* cuda/math.cuh:1766-1768 */
OrcQuat orc_euler_to_quat(OrcVec3 rot);     /* host_math.c:6-21 */
OrcQuat16 orc_quat_pack(OrcQuat q);         /* device_structs.c:388-399 */
OrcVec3 orc_quat_apply(OrcQuat q, OrcVec3 v); /* cuda/math.cuh:411-427 */
OrcVec3 orc_transform_apply(const OrcTransform* t, OrcVec3 v);          /* cuda/math.cuh:459-491 */
OrcVec3 orc_transform_apply_inv(const OrcTransform* t, OrcVec3 v);
OrcVec3 orc_transform_apply_rotation(const OrcTransform* t, OrcVec3 v);
OrcVec3 orc_transform_apply_rotation_inv(const OrcTransform* t, OrcVec3 v);
OrcVec3 orc_transform_apply_relative(const OrcTransform* t, OrcVec3 v);
void orc_material_pack(const OrcMaterialDesc* m, OrcMaterialPacked* out); /* device_structs.c:270-330 */

void orc_camera_sample(const OrcCamera* cam, const OrcSettings* s, OrcPathID id, OrcVec3* origin, OrcVec3* dir);

/* ------------------------------------------------------------------ */
/* orc_trace.c : world-space flattening, BVH2, closest hit             */
/* ------------------------------------------------------------------ */
typedef struct OrcScene OrcScene;

OrcScene* orc_scene_create(
  const OrcMesh* meshes, uint32_t num_meshes, const OrcInstance* instances, uint32_t num_instances, const OrcMaterialPacked* materials,
  uint32_t num_materials);
void orc_scene_destroy(OrcScene* s);
uint32_t orc_scene_num_prims(const OrcScene* s);
/* world-space vertices of flattened primitive p (instance-major order), 9 floats */
const float* orc_scene_world_tris(const OrcScene* s);
void orc_scene_prim_handle(const OrcScene* s, uint32_t prim, uint32_t* instance_id, uint32_t* tri_id);

/* Moeller-Trumbore exactly as cuda/math.cuh:1337-1358. Returns FLT_MAX on miss. */
float orc_tri_mt(const float* v9, OrcVec3 origin, OrcVec3 ray, float* u, float* v);
/* Watertight test (Woop, Benthin, Wald 2013) in the operation order the product uses. Returns false on miss. */
bool orc_tri_watertight(const float* v9, OrcVec3 origin, OrcVec3 ray, float* t, float* u, float* v);

typedef struct {
  uint32_t prim; /* flattened primitive index or ORC_HIT_SKY */
  float t, u, v;
} OrcHit;

/* ------------------------------------------------------------------ */
/* orc_texture.c : material textures (`Texture`, reference texture.h:21-40; device_texture.c; cuda/texture_utils.cuh) */
/* ------------------------------------------------------------------ */
enum { ORC_TEX_FP32 = 0, ORC_TEX_U8 = 1, ORC_TEX_U16 = 2 }; /* TextureDataType, texture.h:7 */
typedef struct {
  uint32_t width, height;
  uint32_t pitch;          /* bytes per row */
  uint32_t type;           /* ORC_TEX_* */
  uint32_t num_components; /* 1, 2 or 4 */
  uint32_t wrap_u, wrap_v; /* TextureWrappingMode: 0 wrap, 1 clamp, 2 mirror, 3 border */
  uint32_t filter;         /* 0 point, 1 linear */
  float gamma;
  const void* data;        /* NULL = invalid texture (TEXTURE_OBJECT_INVALID): loads return their default */
} OrcTexture;

/* copies the descriptors; the texel data stays owned by the caller and must outlive the scene */
void orc_scene_set_textures(OrcScene* s, const OrcTexture* textures, uint32_t count);
/* tex2D<float4> on a normalised-coordinate texture object */
void orc_texture_fetch(const OrcTexture* t, float u, float v, float out[4]);
bool orc_texture_valid(const OrcScene* s, uint16_t tex);
void orc_texture_load(const OrcScene* s, uint16_t tex, float u, float v, bool flip_v, bool apply_gamma, const float def[4], float out[4]);
OrcFloat2 orc_prim_tex_coords(const OrcScene* s, uint32_t prim, float cu, float cv);
bool orc_alpha_cutout(const OrcScene* s, uint32_t prim, float bu, float bv);
void orc_shadow_albedo(const OrcScene* s, uint32_t prim, float bu, float bv, float rgba[4]);
/* one level of the mip chain (device_texture.c:128-245, cuda/mipmap.cuh): dst = (width >> 1) x (height >> 1) texels, 4 components */
void orc_texture_next_mip(const OrcTexture* src, void* dst);
/* light_compute_intensity, cuda/light.cuh:234-262: integrated luminance-texture intensity of one mesh triangle */
float orc_light_intensity(const OrcScene* s, uint32_t mesh_id, uint32_t tri_id);

/* closest hit through the BVH2; ignore_prim = 0xFFFFFFFF for none; counters may be NULL */
OrcHit orc_closest_hit(const OrcScene* s, OrcVec3 origin, OrcVec3 ray, float tmin, float tmax, uint32_t ignore_prim, uint64_t* nodes_visited,
                       uint64_t* tris_tested);
OrcHit orc_closest_hit_bruteforce(const OrcScene* s, OrcVec3 origin, OrcVec3 ray, float tmin, float tmax, uint32_t ignore_prim, int use_mt);

/* Config-1 style batch: primary rays of one sample pass. Outputs are width*height arrays. Returns seconds spent tracing. */
double orc_trace_primary(
  const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t sample_id, uint32_t* out_instance, uint32_t* out_tri, float* out_t,
  float* out_u, float* out_v, int num_threads, uint64_t* nodes_visited, uint64_t* tris_tested);
/* generic batch over explicit rays (origin/dir arrays of 3 floats each) */
double orc_trace_rays(
  const OrcScene* s, const float* origins, const float* dirs, uint32_t n, uint32_t* out_prim, float* out_t, float* out_u, float* out_v,
  int num_threads);

/* ------------------------------------------------------------------ */
/* orc_light.c / orc_shade.c : light tree, BSDF, path tracing          */
/* ------------------------------------------------------------------ */
typedef struct {
  const void* root;     /* DeviceLightTreeRootHeader + sections */
  const void* nodes;    /* DeviceLightTreeNode[] */
  const uint32_t* tri_handle_map; /* pairs (instance_id, tri_id) per light id */
  uint32_t num_lights;
} OrcLightTree;

void orc_scene_set_light_tree(OrcScene* s, const OrcLightTree* tree);
/* LUTs: R16 unorm 32x32 (conductor, glossy), 32x32x32 (dielectric, dielectric_inv) */
void orc_scene_set_bsdf_luts(OrcScene* s, const uint16_t* conductor, const uint16_t* glossy, const uint16_t* dielectric, const uint16_t* dielectric_inv);
void orc_bsdf_lut_generate(uint16_t* conductor, uint16_t* glossy, uint16_t* dielectric, uint16_t* dielectric_inv, uint32_t iterations, int num_threads,
                           int with_dielectric);

void orc_bsdf_lut_dielectric_texel(uint32_t id, uint32_t iterations, uint16_t* out, uint16_t* out_inv);

/* output chain (orc_output.c): accumulation planes -> LuminaryARGB8 {b, g, r, a}; bluenoise_1d NULL = dithering off */
void orc_output_argb8(const float* planes, uint32_t width, uint32_t height, uint32_t sample_count, float exposure, uint32_t tonemap,
                      float agx_slope, float agx_power, float agx_saturation, const uint16_t* bluenoise_1d, uint8_t* dst);
/* + Purkinje shift (cuda/purkinje.cuh) and supersampling: width / height are the rendered resolution */
void orc_output_argb8_ex(const float* planes, uint32_t width, uint32_t height, uint32_t sample_count, float exposure, uint32_t tonemap,
                         float agx_slope, float agx_power, float agx_saturation, const uint16_t* bluenoise_1d, int use_purkinje, float kappa1,
                         float kappa2, uint32_t supersampling, uint8_t* dst);

/* the same chain with every camera parameter of tonemap_apply / convert_RGBF_to_ARGB8 (mirrors Lumb200OutputParams) */
typedef struct {
  float exposure;
  uint32_t tonemap;
  float agx_slope, agx_power, agx_saturation;
  uint32_t dithering, purkinje;
  float purkinje_kappa1, purkinje_kappa2;
  uint32_t supersampling;
  uint32_t filter; /* LuminaryFilter */
  uint32_t use_color_correction;
  float color_correction[3];
  float film_grain;
} OrcOutputParams;
void orc_output_argb8_full(const float* planes, uint32_t width, uint32_t height, uint32_t sample_count, const OrcOutputParams* op,
                           const uint16_t* bluenoise_1d, uint8_t* dst);

/* bloom (device/device_post.c:62-140, cuda/post_common.cuh:71-143): in place on three planes of width * height mean radiance */
void orc_bloom_apply(float* rgb, uint32_t width, uint32_t height, float blend);

/* ------------------------------------------------------------------ */
/* orc_sky.c : procedural atmosphere (cuda/sky.cuh, cuda/sky_utils.cuh, device_sky.c)                                    */
/* ------------------------------------------------------------------ */
typedef struct { /* the fields of `Sky` (structs.h:262-292) that reach the path */
  OrcVec3 geometry_offset;
  float azimuth, altitude, moon_azimuth, moon_altitude, moon_tex_offset;
  float sun_strength, base_density;
  float rayleigh_density, mie_density, ozone_density, rayleigh_falloff, mie_falloff, mie_diameter, ground_visibility, ozone_layer_thickness,
    multiscattering_factor;
  float stars_intensity;
  uint32_t steps, ozone_absorption, stars_count, stars_seed;
  uint32_t aerial_perspective; /* sky_process_inscattering_events between the trace and the sort of every bounce (kernels.cuh:356-389) */
} OrcSkyParams;
void orc_sky_params_default(OrcSkyParams* p); /* sky_get_default, sky.c:6-42 */
/* attaches (p != NULL) or removes the procedural sky: sun / moon positions, star catalogue, and - when a medium parameter changed -
 * the transmittance and multiscattering LUTs (about a second of OpenMP work) */
void orc_scene_set_sky(OrcScene* s, const OrcSkyParams* p, int num_threads);
/* the four LUTs: tm_* = 256 x 64 float4, ms_* = 32 x 32 float4 */
void orc_scene_sky_luts(const OrcScene* s, const float** tm_low, const float** tm_high, const float** ms_low, const float** ms_high);
/* replaces the LUTs (tests: shade with the tables of the implementation under test, so that LUT error does not enter twice) */
void orc_scene_set_sky_luts(OrcScene* s, const float* tm_low, const float* tm_high, const float* ms_low, const float* ms_high);
void orc_scene_sky_info(const OrcScene* s, float sun_pos[3], float moon_pos[3], const float** stars, const uint32_t** stars_offsets,
                        uint32_t* stars_count);
/* sky_color_main (DEFAULT mode) of explicit rays: world-space origins, include_sun = state & (CAMERA_DIRECTION | ALLOW_EMISSION),
 * random_offsets = random_1D(RANDOM_TARGET_SKY_STEP_OFFSET) of each path */
void orc_sky_colors(const OrcScene* s, uint32_t n, const float* origins_world, const float* rays, const uint32_t* include_sun,
                    const float* random_offsets, float* rgb, int num_threads);
/* HDRI mode (sky mode 1): bakes the sky seen from origin_world into a dim x dim latitude / longitude table of float4 with
 * sample_count jittered samples per texel (sky_compute_hdri, sky_hdri.cuh:60-158); sky_color_main then reads the table
 * (point filter) and adds the sun's disc. `mode` selects the march (0) or the table (1). */
void orc_scene_build_sky_hdri(OrcScene* s, const float origin_world[3], uint32_t dim, uint32_t sample_count, int num_threads);
void orc_scene_sky_hdri(const OrcScene* s, const float** color, uint32_t* dim);
/* the moon's surface textures of device_load_embedded_data (moon_albedo.png / moon_normal.png as png_load delivers them); NULL = absent */
void orc_scene_set_moon_textures(OrcScene* s, const OrcTexture* albedo, const OrcTexture* normal);
void orc_scene_set_sky_hdri(OrcScene* s, const float* color, uint32_t dim);
/* aerial perspective: sky_trace_inscattering (sky.cuh:517-532) of explicit segments origin + [0, t] * ray (world space, metres) */
void orc_sky_inscatter_segments(const OrcScene* s, uint32_t n, const float* origins_world, const float* rays, const float* t, uint32_t depth,
                                const float* random_steps, const float* random_offsets, float* inscattering, float* transmittance);
void orc_sky_colors_mode(const OrcScene* s, uint32_t mode, uint32_t n, const float* origins_world, const float* rays, const uint32_t* include_sun,
                         const float* random_offsets, float* rgb, int num_threads);

/* Per-vertex view of geometry_process_tasks (cuda/geometry.cuh:11-180): what the reference kernel writes for ONE task, before
 * any shadow ray. Used to pin the restatement against the reference's own kernel (oracle/_ref/librefdev.so). */
typedef struct {
  OrcPathID path_id;
  uint16_t state;
  OrcVec3 origin, ray;  /* DeviceTask.origin / .ray BEFORE the hit distance is added */
  uint32_t prim;        /* flattened primitive index of the hit */
  float t;              /* DeviceTaskTrace.depth */
  OrcUint2 record;      /* DeviceTaskThroughput.record */
  uint32_t medium_ior;  /* DeviceTaskMediumStack.ior */
} OrcVertexIn;

typedef struct {
  uint32_t geo_light_id; OrcRGB geo_color; OrcVec3 geo_ray; float geo_dist;        /* DeviceTaskDirectLightGeo */
  OrcRGB bsdf_weight; OrcVec3 bsdf_ray; float bsdf_root_sum; float bsdf_prob;      /* DeviceTaskDirectLightBSDF */
  OrcUint2 amb_color; OrcUint2 amb_ray; uint32_t amb_valid;                        /* DeviceTaskDirectLightAmbient */
  OrcRGB emission;                                                                 /* emission x record added to the result */
  uint32_t bounce_alive; uint32_t bounce_state; OrcVec3 bounce_origin; OrcVec3 bounce_ray; OrcUint2 bounce_record; uint32_t bounce_medium_ior;
  OrcRGB bounce_weight; OrcVec3 normal; OrcVec3 hit_point; uint32_t is_transparent_pass; /* diagnostics */
  OrcUint2 sun_color; OrcUint2 sun_ray;                                            /* DeviceTaskDirectLightSun */
} OrcVertexOut;

void orc_shade_vertices(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t n, uint32_t depth, const OrcVertexIn* in,
                        OrcVertexOut* out, int num_threads);

void orc_path_vertices(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t sample_id, uint32_t iter, OrcVertexIn* out,
                       uint8_t* valid, int num_threads);
/* One NEE shadow segment of a path vertex as the product queues it: slot 0 light-tree light, 1 BSDF-sampled light, 2 ambient, 3 sun. */
#define ORC_NEE_SLOTS 4
typedef struct {
  uint32_t valid;       /* the segment carries a non-zero contribution */
  OrcVec3 ray;
  float dist;           /* ORC_FLT_MAX for the ambient segment */
  OrcRGB color;         /* unshadowed contribution x path throughput */
  uint32_t target_prim; /* flattened primitive of the emitter, 0xFFFFFFFF for none */
  OrcRGB visibility;    /* transmittance along the segment */
  uint32_t enum_hits;   /* slot 1 only: emitters counted by the enumeration ray */
} OrcNeeSegment;
void orc_nee_segments(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t n, uint32_t depth, const OrcVertexIn* in,
                      OrcNeeSegment* out /* ORC_NEE_SLOTS per vertex */, int num_threads);
void orc_shadow_rays(const OrcScene* s, uint32_t n, const float* origins, const float* dirs, const float* limits, const uint32_t* ignore_prims,
                     const uint32_t* target_prims, float* visibility, int num_threads);
size_t orc_sizeof_vertex_in(void);
size_t orc_sizeof_vertex_out(void);

typedef struct {
  uint64_t closest_rays;
  uint64_t shadow_rays;
  uint64_t light_enum_rays;
} OrcRayCounts;

/* Renders sample ids [first_sample, first_sample + num_samples) for every pixel and ADDS the results into the four
 * planes (sum R, sum G, sum B, sum luminance(colour^2)), each width*height floats. Returns seconds. */
double orc_render(
  const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t first_sample, uint32_t num_samples, float* planes, int num_threads,
  OrcRayCounts* counts);
/* Debug shading modes (LuminaryShadingMode, structs.h; settings.shading_mode != DEFAULT): one bounce, the hit / miss is painted by
 * geometry_process_tasks_debug (cuda/geometry.cuh:182-246) / sky_process_tasks_debug (cuda/sky.cuh:635-668). Adds into the planes like orc_render. */
#define ORC_SHADING_MODE_ALBEDO 1
#define ORC_SHADING_MODE_DEPTH 2
#define ORC_SHADING_MODE_NORMAL 3
#define ORC_SHADING_MODE_IDENTIFICATION 4
#define ORC_SHADING_MODE_LIGHTS 5
double orc_render_debug(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t shading_mode, uint32_t first_sample,
                        uint32_t num_samples, float* planes, int num_threads);
/* same, restricted to the pixel rectangle [x0,x1) x [y0,y1) (bounded CPU baseline sample) */
double orc_render_region(
  const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t first_sample, uint32_t num_samples, uint32_t x0, uint32_t y0,
  uint32_t x1, uint32_t y1, float* planes, int num_threads, OrcRayCounts* counts);

/* ------------------------------------------------------------------ */
/* orc_adaptive.c : adaptive sampling (cuda/adaptive_sampling.cuh, device/device_adaptive_sampler.c)                    */
/* ------------------------------------------------------------------ */
#define ORC_ADAPTIVE_STAGES 4 /* ADAPTIVE_SAMPLER_NUM_STAGES, device_utils.h:331 */
typedef struct {
  uint32_t max_sampling_rate; /* 1..256 */
  uint32_t avg_sampling_rate;
  uint32_t update_interval;   /* executions of stage s before stage s + 1 is built: update_interval << s */
  float exposure;             /* linear exposure when exposure-aware, else 0 */
  uint32_t tonemap;
  float agx_slope, agx_power, agx_saturation;
} OrcAdaptiveParams;

uint32_t orc_adaptive_stage_count(uint32_t word, uint32_t stage);
uint32_t orc_adaptive_block_samples(uint32_t word, const uint32_t executions[ORC_ADAPTIVE_STAGES + 1]);
float orc_adaptive_block_variance(const float* planes, uint32_t width, uint32_t height, const uint32_t* words,
                                  const uint32_t executions[ORC_ADAPTIVE_STAGES + 1], const OrcAdaptiveParams* p, float* block_variance);
void orc_adaptive_stage_counts(const float* block_variance, float sum_variance, uint32_t num_blocks, uint32_t stage, const OrcAdaptiveParams* p,
                               uint32_t* words);
/* Runs `num_executions` executions of the adaptive schedule from the state (words, executions, stage) and ADDS into the four
 * planes; words = ceil(w / 4) * ceil(h / 4) entries, executions = 5 counters, *stage in 0..4. Returns the number of paths traced. */
uint64_t orc_render_adaptive(const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, const OrcAdaptiveParams* p, uint32_t num_executions,
                             float* planes, uint32_t* words, uint32_t* executions, uint32_t* stage, int num_threads, OrcRayCounts* counts);
/* accumulation_generate_result (beauty): mean = first moment / the pixel's own sample count -> 3 planes */
void orc_adaptive_resolve(const float* planes, uint32_t width, uint32_t height, const uint32_t* words, const uint32_t* executions, float* rgb);
/* all output modes (0 beauty, 1 variance, 2 error, 3 sample distribution) + local error minimisation; words NULL = uniform_count */
void orc_resolve(const float* planes, uint32_t width, uint32_t height, const uint32_t* words, const uint32_t* executions, uint32_t uniform_count,
                 uint32_t mode, int local_error_minimization, uint32_t stage, const OrcAdaptiveParams* p, float* rgb);

#ifdef __cplusplus
}
#endif

#endif /* LUM_ORACLE_H */
