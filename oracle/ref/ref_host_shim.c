/* ref_host_shim.c - flat C entry points around the REFERENCE's own host-side C functions.
 *
 * TEST INFRASTRUCTURE ONLY (same rule as oracle/lum_oracle.h): only tests/, __graft_entry__.smoke() and the cpu_baseline
 * leg of bench.py may load the library this file is linked into. The product never does.
 *
 * This file contains no reference code. It is compiled together with the reference's unmodified C sources, taken from
 * where they lie under /root/reference (oracle/ref/Makefile), into oracle/_ref/libref_host.so:
 *   device/device_structs.c   scene entity -> device struct packers (settings, camera, sky, material, vertices, transforms)
 *   device/device_packing.c   normal / uv packing
 *   device/device_light.c     light tree build (binned SAH, collapse, quantisation, finalise)
 *   device/device_sky.c       star catalogue generation (the LUT / HDRI halves need a device and are stubbed out)
 *   host_math.c, camera.c, settings.c, sky.c, material.c, mesh.c, array.c, hashmap.c, host_memory.c, log.c, error.c
 * so that the repo's restatements (oracle/orc_core.c packers, luminary_b200/csrc/host/light_tree.c) can be pinned against the
 * running reference instead of against a second reading of its source.
 *
 * Inputs use the repo's C-ABI structs (include/lumb200.h); they are mapped field by field onto the reference's public
 * structs (include/luminary/structs.h) starting from the reference's own *_get_default values.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "camera.h"
#include "device/device_light.h"
#include "device/device_packing.h"
#include "device/device_sky.h"
#include "device/device_structs.h"
#include "internal_error.h"
#include "material.h"
#include "mesh.h"
#include "settings.h"
#include "sky.h"
#include "utils.h"

#include "../../include/lumb200.h"

#define REF_TRY(expr)                                                              \
  do {                                                                             \
    const LuminaryResult _r = (expr);                                              \
    if (_r != LUMINARY_SUCCESS) {                                                  \
      fprintf(stderr, "oracle/_ref: %s failed: %s\n", #expr, luminary_result_to_string(_r)); \
      return (int) (_r & 0xFFFF) | 0x10000;                                        \
    }                                                                              \
  } while (0)

static void material_from_abi(const Lumb200Material* m, uint32_t id, Material* out) {
  material_get_default(out);
  out->id                       = id;
  out->base_substrate           = (MaterialBaseSubstrate) m->base_substrate;
  out->albedo                   = (RGBAF) {.r = m->albedo[0], .g = m->albedo[1], .b = m->albedo[2], .a = m->albedo[3]};
  out->emission                 = (RGBF) {.r = m->emission[0], .g = m->emission[1], .b = m->emission[2]};
  out->emission_scale           = m->emission_scale;
  out->roughness                = m->roughness;
  out->roughness_clamp          = m->roughness_clamp;
  out->refraction_index         = m->refraction_index;
  out->emission_active          = m->emission_active != 0;
  out->thin_walled              = m->thin_walled != 0;
  out->metallic                 = m->metallic != 0;
  out->colored_transparency     = m->colored_transparency != 0;
  out->roughness_as_smoothness  = m->roughness_as_smoothness != 0;
  out->normal_map_is_compressed = m->normal_map_is_compressed != 0;
  out->bidirectional_emission   = m->bidirectional_emission != 0;
  out->albedo_tex               = m->albedo_tex;
  out->luminance_tex            = m->luminance_tex;
  out->roughness_tex            = m->roughness_tex;
  out->metallic_tex             = m->metallic_tex;
  out->normal_tex               = m->normal_tex;
}

static void instance_from_abi(const Lumb200Instance* in, uint32_t id, MeshInstance* out) {
  memset(out, 0, sizeof(*out));
  out->id          = id;
  out->mesh_id     = in->mesh_id;
  out->translation = (vec3) {.x = in->translation[0], .y = in->translation[1], .z = in->translation[2]};
  out->rotation    = (vec3) {.x = in->rotation[0], .y = in->rotation[1], .z = in->rotation[2]};
  out->scale       = (vec3) {.x = in->scale[0], .y = in->scale[1], .z = in->scale[2]};
  out->active      = in->active != 0;
}

void refhost_camera_from_abi(const Lumb200Camera* in, Camera* out) {
  camera_get_default(out);
  out->pos                        = (vec3) {.x = in->pos[0], .y = in->pos[1], .z = in->pos[2]};
  out->rotation                   = (vec3) {.x = in->rotation[0], .y = in->rotation[1], .z = in->rotation[2]};
  out->thin_lens.fov              = in->fov;
  out->thin_lens.aperture_size    = in->aperture_size;
  out->object_distance            = in->object_distance;
  out->camera_scale               = in->camera_scale;
  out->russian_roulette_threshold = in->russian_roulette_threshold;
  out->aperture_shape             = (ApertureShape) in->aperture_shape;
  out->aperture_blade_count       = in->aperture_blade_count;
  out->use_physical_camera        = false;
  out->physical.use_spectral_rendering = false;
}

/* device_struct_material_convert, device_structs.c:263-313 -> 32 bytes */
int refhost_material_convert(const Lumb200Material* m, void* out32) {
  Material mat;
  material_from_abi(m, 0, &mat);
  DeviceMaterialCompressed dm;
  memset(&dm, 0, sizeof(dm));
  REF_TRY(device_struct_material_convert(&mat, &dm));
  memcpy(out32, &dm, sizeof(dm));
  return 0;
}

/* device_struct_camera_convert, device_structs.c:40-85 -> DeviceCamera (108 bytes) */
int refhost_camera_convert(const Lumb200Camera* c, void* out, size_t out_size) {
  Camera cam;
  refhost_camera_from_abi(c, &cam);
  DeviceCamera dc;
  memset(&dc, 0, sizeof(dc));
  REF_TRY(device_struct_camera_convert(&cam, &dc));
  if (out_size < sizeof(dc))
    return 1;
  memcpy(out, &dc, sizeof(dc));
  return 0;
}

size_t refhost_sizeof_device_camera(void) { return sizeof(DeviceCamera); }

/* device_struct_instance_transform_convert, device_structs.c:401-413 -> DeviceTransform (32 bytes) */
int refhost_instance_transform_convert(const Lumb200Instance* in, void* out32) {
  MeshInstance mi;
  instance_from_abi(in, 0, &mi);
  DeviceTransform t;
  memset(&t, 0, sizeof(t));
  REF_TRY(device_struct_instance_transform_convert(&mi, &t));
  memcpy(out32, &t, sizeof(t));
  return 0;
}

/* device_struct_vertex_convert / device_struct_triangle_texture_convert, device_structs.c:351-374:
 * vertices_out = 3 x 16 bytes per triangle, textris_out = 16 bytes per triangle. */
int refhost_mesh_convert(const Lumb200Mesh* mesh, void* vertices_out, void* textris_out) {
  TriangleGeomData data;
  data.vertex_buffer      = (float*) mesh->vertex_buffer;
  data.normal_buffer      = (float*) mesh->normal_buffer;
  data.uv_buffer          = (float*) mesh->uv_buffer;
  data.material_id_buffer = (uint16_t*) mesh->material_id_buffer;
  data.triangle_count     = mesh->triangle_count;
  DeviceTriangleVertex* v  = (DeviceTriangleVertex*) vertices_out;
  DeviceTriangleTexture* t = (DeviceTriangleTexture*) textris_out;
  for (uint32_t i = 0; i < mesh->triangle_count; i++) {
    for (uint32_t k = 0; k < 3; k++)
      REF_TRY(device_struct_vertex_convert(&data, 3 * i + k, v + 3 * i + k));
    REF_TRY(device_struct_triangle_texture_convert(&data, i, t + i));
  }
  return 0;
}

uint32_t refhost_pack_normal(float x, float y, float z) { return device_pack_normal((vec3) {.x = x, .y = y, .z = z}); }
uint32_t refhost_pack_uv(float u, float v) { return device_pack_uv((UV) {.u = u, .v = v}); }

/* light_tree_create / update_cache_* / build, device_light.c:1701-1900, 2236-2268. Outputs are malloc'ed copies. */
int refhost_light_tree_build(const Lumb200Mesh* meshes, uint32_t num_meshes, const Lumb200Instance* instances, uint32_t num_instances,
                             const Lumb200Material* materials, uint32_t num_materials, Lumb200LightTreeBuffers* out, void** bvh_vertices,
                             size_t* bvh_vertices_size) {
  memset(out, 0, sizeof(*out));
  LightTree* tree;
  REF_TRY(light_tree_create(&tree));

  for (uint32_t i = 0; i < num_materials; i++) {
    Material mat;
    material_from_abi(materials + i, i, &mat);
    REF_TRY(light_tree_update_cache_material(tree, &mat));
  }
  for (uint32_t i = 0; i < num_meshes; i++) {
    const uint32_t n = meshes[i].triangle_count;
    /* the reference loads 4 floats per vertex (device_light.c:1673-1675): pad the copy */
    float* padded = (float*) calloc((size_t) n * 9 + 4, sizeof(float));
    memcpy(padded, meshes[i].vertex_buffer, (size_t) n * 9 * sizeof(float));
    Mesh mesh;
    memset(&mesh, 0, sizeof(mesh));
    mesh.id                      = i;
    mesh.data.vertex_buffer      = padded;
    mesh.data.normal_buffer      = (float*) meshes[i].normal_buffer;
    mesh.data.uv_buffer          = (float*) meshes[i].uv_buffer;
    mesh.data.material_id_buffer = (uint16_t*) meshes[i].material_id_buffer;
    mesh.data.triangle_count     = n;
    const LuminaryResult r       = light_tree_update_cache_mesh(tree, &mesh);
    free(padded);
    REF_TRY(r);
  }
  for (uint32_t i = 0; i < num_instances; i++) {
    MeshInstance mi;
    instance_from_abi(instances + i, i, &mi);
    REF_TRY(light_tree_update_cache_instance(tree, &mi));
  }

  static uint64_t fake_device[8192]; /* only NULL-checked when no emitter is textured (device_light.c:1952-1960) */
  REF_TRY(light_tree_build(tree, (Device*) fake_device));

  out->num_lights = tree->light_count;
  out->root_size  = tree->root_size;
  out->nodes_size = tree->nodes_size;
  if (tree->root_size) {
    out->root_data = malloc(tree->root_size);
    memcpy(out->root_data, tree->root_data, tree->root_size);
  }
  if (tree->nodes_size) {
    out->nodes_data = malloc(tree->nodes_size);
    memcpy(out->nodes_data, tree->nodes_data, tree->nodes_size);
  }
  if (tree->light_count) {
    out->tri_handle_map = (uint32_t*) malloc(sizeof(TriangleHandle) * tree->light_count);
    memcpy(out->tri_handle_map, tree->tri_handle_map_data, sizeof(TriangleHandle) * tree->light_count);
    if (bvh_vertices) {
      const size_t sz = sizeof(float) * 4 * 3 * tree->light_count;
      *bvh_vertices   = malloc(sz);
      memcpy(*bvh_vertices, tree->bvh_vertex_buffer_data, sz);
      *bvh_vertices_size = sz;
    }
  }
  REF_TRY(light_tree_destroy(&tree));
  return 0;
}

/* device_struct_settings_convert, device_structs.c:11-38 -> DeviceRendererSettings (16 bytes). supersampling 0, full-frame region
 * (what the benchmark drivers set through luminary_host_set_settings). */
int refhost_settings_convert(uint32_t width, uint32_t height, uint32_t max_ray_depth, void* out16) {
  RendererSettings s;
  REF_TRY(settings_get_default(&s));
  s.width         = width;
  s.height        = height;
  s.max_ray_depth = max_ray_depth;
  s.supersampling = 0;
  s.undersampling = 0;
  s.region_x      = 0.0f;
  s.region_y      = 0.0f;
  s.region_width  = 1.0f;
  s.region_height = 1.0f;
  DeviceRendererSettings ds;
  memset(&ds, 0, sizeof(ds));
  REF_TRY(device_struct_settings_convert(&s, &ds));
  memcpy(out16, &ds, sizeof(ds));
  return 0;
}

/* device_struct_sky_convert, device_structs.c:107-180 -> DeviceSky (104 bytes) */
int refhost_sky_convert(uint32_t mode, const float* constant_color, void* out, size_t out_size) {
  Sky sky;
  REF_TRY(sky_get_default(&sky));
  sky.mode           = (LuminarySkyMode) mode;
  sky.constant_color = (RGBF) {.r = constant_color[0], .g = constant_color[1], .b = constant_color[2]};
  DeviceSky ds;
  memset(&ds, 0, sizeof(ds));
  REF_TRY(device_struct_sky_convert(&sky, &ds));
  if (out_size < sizeof(ds))
    return 1;
  memcpy(out, &ds, sizeof(ds));
  return 0;
}

size_t refhost_sizeof_device_sky(void) { return sizeof(DeviceSky); }

/* The same, with every field of `Sky` that reaches the path set by the caller (layout = OrcSkyParams of oracle/lum_oracle.h). */
typedef struct {
  float geometry_offset[3];
  float azimuth, altitude, moon_azimuth, moon_altitude, moon_tex_offset;
  float sun_strength, base_density;
  float rayleigh_density, mie_density, ozone_density, rayleigh_falloff, mie_falloff, mie_diameter, ground_visibility, ozone_layer_thickness,
    multiscattering_factor;
  float stars_intensity;
  uint32_t steps, ozone_absorption, stars_count, stars_seed;
  uint32_t aerial_perspective;
} RefSkyParams;

static void sky_from_params(const RefSkyParams* p, uint32_t mode, const float* constant_color, Sky* sky) {
  sky_get_default(sky);
  sky->mode                   = (LuminarySkyMode) mode;
  sky->constant_color         = (RGBF) {.r = constant_color[0], .g = constant_color[1], .b = constant_color[2]};
  sky->geometry_offset        = (vec3) {.x = p->geometry_offset[0], .y = p->geometry_offset[1], .z = p->geometry_offset[2]};
  sky->azimuth                = p->azimuth;
  sky->altitude               = p->altitude;
  sky->moon_azimuth           = p->moon_azimuth;
  sky->moon_altitude          = p->moon_altitude;
  sky->moon_tex_offset        = p->moon_tex_offset;
  sky->sun_strength           = p->sun_strength;
  sky->base_density           = p->base_density;
  sky->rayleigh_density       = p->rayleigh_density;
  sky->mie_density            = p->mie_density;
  sky->ozone_density          = p->ozone_density;
  sky->rayleigh_falloff       = p->rayleigh_falloff;
  sky->mie_falloff            = p->mie_falloff;
  sky->mie_diameter           = p->mie_diameter;
  sky->ground_visibility      = p->ground_visibility;
  sky->ozone_layer_thickness  = p->ozone_layer_thickness;
  sky->multiscattering_factor = p->multiscattering_factor;
  sky->stars_intensity        = p->stars_intensity;
  sky->steps                  = p->steps;
  sky->ozone_absorption       = p->ozone_absorption != 0;
  sky->stars_count            = p->stars_count;
  sky->stars_seed             = p->stars_seed;
  sky->aerial_perspective     = p->aerial_perspective != 0;
}

int refhost_sky_convert_params(const RefSkyParams* p, uint32_t mode, const float* constant_color, void* out, size_t out_size) {
  Sky sky;
  sky_from_params(p, mode, constant_color, &sky);
  DeviceSky ds;
  memset(&ds, 0, sizeof(ds));
  REF_TRY(device_struct_sky_convert(&sky, &ds));
  if (out_size < sizeof(ds))
    return 1;
  memcpy(out, &ds, sizeof(ds));
  return 0;
}

/* sky_get_default (sky.c:6-42) in the RefSkyParams layout */
void refhost_sky_default_params(RefSkyParams* p) {
  Sky sky;
  sky_get_default(&sky);
  p->geometry_offset[0] = sky.geometry_offset.x, p->geometry_offset[1] = sky.geometry_offset.y, p->geometry_offset[2] = sky.geometry_offset.z;
  p->azimuth = sky.azimuth, p->altitude = sky.altitude, p->moon_azimuth = sky.moon_azimuth, p->moon_altitude = sky.moon_altitude;
  p->moon_tex_offset = sky.moon_tex_offset, p->sun_strength = sky.sun_strength, p->base_density = sky.base_density;
  p->rayleigh_density = sky.rayleigh_density, p->mie_density = sky.mie_density, p->ozone_density = sky.ozone_density;
  p->rayleigh_falloff = sky.rayleigh_falloff, p->mie_falloff = sky.mie_falloff, p->mie_diameter = sky.mie_diameter;
  p->ground_visibility = sky.ground_visibility, p->ozone_layer_thickness = sky.ozone_layer_thickness;
  p->multiscattering_factor = sky.multiscattering_factor, p->stars_intensity = sky.stars_intensity;
  p->steps = sky.steps, p->ozone_absorption = sky.ozone_absorption ? 1u : 0u, p->stars_count = sky.stars_count, p->stars_seed = sky.stars_seed;
  p->aerial_perspective = sky.aerial_perspective ? 1u : 0u;
}

/* sky_stars_create + sky_stars_update (device_sky.c:470-572): the star catalogue of (seed, count), 4 floats per star
 * (altitude, azimuth, radius, intensity) sorted by grid cell, and the STARS_GRID_LD x 32 + 1 cell offsets. */
int refhost_stars_generate(uint32_t seed, uint32_t count, float* stars_out, uint32_t* offsets_out) {
  SkyStars* stars;
  REF_TRY(sky_stars_create(&stars));
  Sky sky;
  sky_get_default(&sky);
  sky.stars_seed  = seed;
  sky.stars_count = count;
  stars->seed     = ~seed; /* force generation whatever the defaults are */
  stars->count    = ~count;
  REF_TRY(sky_stars_update(stars, &sky));
  _Static_assert(sizeof(Star) == 16, "Star layout");
  memcpy(stars_out, stars->data, sizeof(Star) * (size_t) count);
  memcpy(offsets_out, stars->offsets, sizeof(uint32_t) * (STARS_GRID_LD * 32 + 1));
  REF_TRY(sky_stars_destroy(&stars));
  return 0;
}

void refhost_free(void* p) { free(p); }
