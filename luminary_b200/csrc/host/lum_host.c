/* lum_host.c - implementation of Luminary's public host API (include/luminary/luminary.h) on top of the C ABI of the
 * B200 device path (include/lumb200.h). Plain C11 + pthreads, like the reference's host layer.
 *
 * Mirrors src/luminary/host/host.c of the reference:
 *   - the application thread only edits the caller-side scene and enqueues (host.c:239-286): setters copy PODs under
 *     a mutex, nothing blocks on the GPU;
 *   - one worker thread owns all devices (device_manager.c:828-832). On luminary_host_start_new_render it snapshots
 *     the scene, uploads it to every enabled device (device_manager.c:281-513), builds the light tree on the CPU
 *     (device_light.c:2236) and the acceleration structures on the device, and then renders sample passes;
 *   - sample ids are partitioned over the devices (device_adaptive_sampler.c:58-71): device g of G renders ids
 *     g, g + G, ...; at every requested output the secondary devices' accumulation planes are added into the main
 *     device's (device_result_interface.c:107-299 - here a peer-to-peer copy over NVLink instead of pinned host memory);
 *   - outputs follow the promise protocol of host.h:73-86: request_output -> try_await_output -> get_image -> release.
 * Deviations are listed in include/luminary/luminary.h. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "lum_host_internal.h"

#define LUM_MAX_DEVICES 8
#define LUM_PASSES_PER_CHUNK 32u /* passes queued per device between two checks of the restart / shutdown flags */

/* ------------------------------------------------------------------------------------------------------------ */
/* errors and logging                                                                                           */
/* ------------------------------------------------------------------------------------------------------------ */
static _Thread_local char g_error[1024] = "";
static int g_log_level                  = 1; /* 0 quiet, 1 warnings + info, 2 everything; LUMINARY_B200_LOG overrides */

void lum_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  if (g_log_level >= 1)
    fprintf(stderr, "[luminary_b200][error] %s\n", g_error);
}

void lum_log(const char* level, const char* fmt, ...) {
  const int need = (!strcmp(level, "log")) ? 2 : 1;
  if (g_log_level < need)
    return;
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "[luminary_b200][%s] ", level);
  vfprintf(stderr, fmt, ap);
  fputc('\n', stderr);
  va_end(ap);
}

const char* luminary_b200_last_error(void) { return g_error; }

const char* luminary_result_to_string(LuminaryResult result) { /* reference error.c */
  switch (result & ~LUMINARY_ERROR_PROPAGATED) {
    case LUMINARY_SUCCESS: return "Success";
    case LUMINARY_ERROR_ARGUMENT_NULL: return "Argument was NULL";
    case LUMINARY_ERROR_NOT_IMPLEMENTED: return "Not implemented";
    case LUMINARY_ERROR_INVALID_API_ARGUMENT: return "Invalid API argument";
    case LUMINARY_ERROR_MEMORY_LEAK: return "Memory leak";
    case LUMINARY_ERROR_OUT_OF_MEMORY: return "Out of memory";
    case LUMINARY_ERROR_C_STD: return "C standard library error";
    case LUMINARY_ERROR_API_EXCEPTION: return "API exception";
    case LUMINARY_ERROR_CUDA: return "CUDA error";
    case LUMINARY_ERROR_OPTIX: return "OptiX error";
    case LUMINARY_ERROR_PREVIOUS_ERROR: return "Previous error";
    case LUMINARY_ERROR_DEBUG_ASSERT: return "Debug assertion";
    case LUMINARY_ERROR_MISSING_DATA: return "Missing data";
    case LUMINARY_ERROR_INVALID_DEVICE: return "Invalid device";
    default: return "Unknown error";
  }
}

static LuminaryResult from_device(Lumb200Result r) {
  if (r == LUMB200_SUCCESS)
    return LUMINARY_SUCCESS;
  lum_set_error("%s", lumb200_last_error());
  return (LuminaryResult) r | LUMINARY_ERROR_PROPAGATED; /* same numbering on both sides of the ABI */
}
#define DEV_TRY(expr) LUM_TRY(from_device(expr))

void luminary_init(void) {
  const char* e = getenv("LUMINARY_B200_LOG");
  if (e)
    g_log_level = atoi(e);
}

void luminary_shutdown(void) {}

/* ------------------------------------------------------------------------------------------------------------ */
/* paths                                                                                                        */
/* ------------------------------------------------------------------------------------------------------------ */
LuminaryResult luminary_path_create(LuminaryPath** path) {
  LUM_CHECK_NULL(path);
  *path = (LuminaryPath*) calloc(1, sizeof(LuminaryPath));
  if (!*path)
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "out of host memory");
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_path_set_from_string(LuminaryPath* path, const char* string) {
  LUM_CHECK_NULL(path);
  LUM_CHECK_NULL(string);
  free(path->string);
  path->string = strdup(string);
  return path->string ? LUMINARY_SUCCESS : LUMINARY_ERROR_OUT_OF_MEMORY;
}

LuminaryResult luminary_path_destroy(LuminaryPath** path) {
  LUM_CHECK_NULL(path);
  LUM_CHECK_NULL(*path);
  free((*path)->string);
  free(*path);
  *path = NULL;
  return LUMINARY_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* host object                                                                                                  */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  uint32_t sample_count, width, height;
  bool done;
  LuminaryOutputHandle output;
} OutputRequest;

typedef struct {
  uint8_t* buffer;
  uint32_t width, height;
  float time;
  uint32_t sample_count;
  int refs;
  bool valid;
} OutputSlot;

typedef struct {
  Lumb200Device* dev;
  uint32_t cuda_index;
  bool enabled;
  uint32_t meshes_uploaded;
  uint32_t textures_uploaded;
  bool data_loaded;
  char name[256];
  size_t memory;
  double gpu_seconds; /* cumulative over the current render, survives the resets of secondary devices */
  uint64_t rays;
} HostDevice;

struct LuminaryHost {
  pthread_mutex_t lock;
  pthread_cond_t wake, idle;
  pthread_t worker;
  bool worker_started, shutdown, busy;
  bool hdri_request; /* luminary_host_request_sky_hdri_build: re-bake at the next render even if the sky is unchanged (moved camera) */

  /* caller-side scene (reference: scene_caller) */
  LuminaryRendererSettings settings;
  LuminaryCamera camera;
  LuminarySky sky;
  LumHostMesh* meshes;
  uint32_t num_meshes;
  LuminaryMaterial* materials;
  uint32_t num_materials;
  LuminaryInstance* instances;
  uint32_t num_instances;
  LumHostTexture* textures; /* append-only, like the meshes (device_manager_add_textures, device_manager.c:1065) */
  uint32_t num_textures;
  /* entities outside the path: stored so that get returns what set stored, never active */
  LuminaryOcean ocean;
  LuminaryCloud cloud;
  LuminaryFog fog;
  LuminaryParticles particles;
  /* recurring outputs (luminary_host_set_output_properties / acquire_output, reference host_output_handler.c): while enabled the
   * render keeps going and a fresh output of the render resolution is produced after every chunk of passes */
  LuminaryOutputProperties output_properties;

  HostDevice devices[LUM_MAX_DEVICES];
  uint32_t num_devices;
  /* NCCL communicator over the enabled devices (lumb200_comm_create_all), rebuilt when the enabled set changes */
  Lumb200Comm* comms[LUM_MAX_DEVICES];
  uint32_t num_comms;
  uint32_t comm_mask; /* bit g: devices[g] is a member */

  uint32_t requested_generation, finished_generation;
  uint32_t parked_generation; /* != 0: the worker sits inside this (live) render with every requested output produced */
  LuminaryResult worker_error;
  uint32_t samples_done;
  double sample_time;
  uint64_t rays;

  OutputRequest* requests;
  uint32_t num_requests;
  OutputSlot* outputs;
  uint32_t num_outputs;
  LuminaryOutputHandle latest_output;

  const char* task;
  struct timespec task_start;
};

static double now_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static void set_task(LuminaryHost* h, const char* task) {
  pthread_mutex_lock(&h->lock);
  h->task = task;
  clock_gettime(CLOCK_MONOTONIC, &h->task_start);
  pthread_mutex_unlock(&h->lock);
}

static char* data_file_path(const char* name) {
  Dl_info info;
  static char path[4096];
  if (dladdr((void*) &data_file_path, &info) && info.dli_fname) {
    const char* slash = strrchr(info.dli_fname, '/');
    if (slash) {
      snprintf(path, sizeof(path), "%.*s/data/%s", (int) (slash - info.dli_fname), info.dli_fname, name);
      return path;
    }
  }
  snprintf(path, sizeof(path), "data/%s", name);
  return path;
}

static void* read_file(const char* path, size_t expect) {
  FILE* f = fopen(path, "rb");
  if (!f)
    return NULL;
  void* buf = malloc(expect);
  if (buf && fread(buf, 1, expect, f) != expect) {
    free(buf);
    buf = NULL;
  }
  fclose(f);
  return buf;
}

/* device_load_embedded_data (device_embedded_data.c): the two blue-noise masks shipped next to the library */
static LuminaryResult load_embedded_data(HostDevice* d) {
  if (d->data_loaded)
    return LUMINARY_SUCCESS;
  uint32_t* bn2 = (uint32_t*) read_file(data_file_path("bluenoise_2D.bin"), 256 * 256 * 4);
  if (!bn2)
    LUM_RETURN_ERROR(LUMINARY_ERROR_MISSING_DATA, "embedded file %s is missing", data_file_path("bluenoise_2D.bin"));
  Lumb200Result r = lumb200_device_load_bluenoise(d->dev, bn2, 256 * 256);
  free(bn2);
  DEV_TRY(r);
  uint16_t* bn1 = (uint16_t*) read_file(data_file_path("bluenoise_1D.bin"), 256 * 256 * 2);
  if (!bn1)
    LUM_RETURN_ERROR(LUMINARY_ERROR_MISSING_DATA, "embedded file %s is missing", data_file_path("bluenoise_1D.bin"));
  r = lumb200_device_load_bluenoise_1d(d->dev, bn1, 256 * 256);
  free(bn1);
  DEV_TRY(r);
  DEV_TRY(lumb200_device_build_bsdf_lut(d->dev));
  { /* device_embedded_data.c:62-100: the moon's surface through the PNG reader (RGBA8, wrap, linear); a missing file leaves a black disc */
    LumHostTexture moon[2];
    Lumb200Texture desc[2];
    const char* names[2] = {"moon_albedo.png", "moon_normal.png"};
    bool have[2]         = {false, false};
    memset(moon, 0, sizeof(moon));
    memset(desc, 0, sizeof(desc));
    for (int k = 0; k < 2; k++) {
      if (lum_png_read(data_file_path(names[k]), &moon[k]) != LUMINARY_SUCCESS || !moon[k].data) {
        lum_log("warn", "embedded file %s is missing or unreadable: the moon's surface stays black", names[k]);
        continue;
      }
      desc[k].width = moon[k].width, desc[k].height = moon[k].height, desc[k].pitch = moon[k].pitch;
      desc[k].type = moon[k].type, desc[k].num_components = moon[k].num_components;
      desc[k].wrap_mode_u = LUMB200_WRAP_WRAP, desc[k].wrap_mode_v = LUMB200_WRAP_WRAP;
      desc[k].filter = LUMB200_FILTER_LINEAR, desc[k].gamma = moon[k].gamma, desc[k].mipmap = 0, desc[k].data = moon[k].data;
      have[k] = true;
    }
    r = lumb200_device_load_moon_textures(d->dev, have[0] ? &desc[0] : NULL, have[1] ? &desc[1] : NULL);
    for (int k = 0; k < 2; k++)
      free(moon[k].data);
    DEV_TRY(r);
  }
  d->data_loaded = true;
  return LUMINARY_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* worker: scene upload + sample passes                                                                         */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  LuminaryRendererSettings settings;
  LuminaryCamera camera;
  LuminarySky sky;
  LumHostMesh* meshes;
  uint32_t num_meshes;
  LuminaryMaterial* materials;
  uint32_t num_materials;
  LuminaryInstance* instances;
  uint32_t num_instances;
  LumHostTexture* textures;
  uint32_t num_textures;
} SceneSnapshot;

static void snapshot_free(SceneSnapshot* s) {
  free(s->textures);
  free(s->meshes);
  free(s->materials);
  free(s->instances);
  memset(s, 0, sizeof(*s));
}

static void convert_material(const LuminaryMaterial* m, Lumb200Material* o) {
  memset(o, 0, sizeof(*o));
  o->base_substrate = (uint32_t) m->base_substrate;
  o->albedo[0] = m->albedo.r, o->albedo[1] = m->albedo.g, o->albedo[2] = m->albedo.b, o->albedo[3] = m->albedo.a;
  o->emission[0] = m->emission.r, o->emission[1] = m->emission.g, o->emission[2] = m->emission.b;
  o->emission_scale           = m->emission_scale;
  o->roughness                = m->roughness;
  o->roughness_clamp          = m->roughness_clamp;
  o->refraction_index         = m->refraction_index;
  o->emission_active          = m->emission_active;
  o->thin_walled              = m->thin_walled;
  o->metallic                 = m->metallic;
  o->colored_transparency     = m->colored_transparency;
  o->roughness_as_smoothness  = m->roughness_as_smoothness;
  o->normal_map_is_compressed = m->normal_map_is_compressed;
  o->bidirectional_emission   = m->bidirectional_emission;
  o->albedo_tex               = m->albedo_tex;
  o->luminance_tex            = m->luminance_tex;
  o->roughness_tex            = m->roughness_tex;
  o->metallic_tex             = m->metallic_tex;
  o->normal_tex               = m->normal_tex;
}

/* device_add_mesh / device_add_textures for everything the device has not seen yet (both lists are append-only) */
static LuminaryResult upload_meshes_and_textures(HostDevice* d, const SceneSnapshot* s, const Lumb200Mesh* meshes) {
  for (uint32_t k = d->meshes_uploaded; k < s->num_meshes; k++) {
    uint32_t id = 0;
    DEV_TRY(lumb200_device_add_mesh(d->dev, &meshes[k], &id));
    d->meshes_uploaded = k + 1;
  }
  if (d->textures_uploaded < s->num_textures) { /* device_add_textures, device.h:160 */
    const uint32_t first = d->textures_uploaded, n = s->num_textures - first;
    Lumb200Texture* tex  = (Lumb200Texture*) calloc(n, sizeof(Lumb200Texture));
    if (!tex)
      LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "out of host memory");
    for (uint32_t k = 0; k < n; k++) {
      const LumHostTexture* t = &s->textures[first + k];
      tex[k].width = t->width, tex[k].height = t->height, tex[k].pitch = t->pitch;
      tex[k].type = t->type, tex[k].num_components = t->num_components;
      tex[k].wrap_mode_u = LUMB200_WRAP_WRAP, tex[k].wrap_mode_v = LUMB200_WRAP_WRAP; /* texture_create, texture.c:80-84 */
      tex[k].filter = LUMB200_FILTER_LINEAR;
      tex[k].gamma  = t->gamma;
      tex[k].mipmap = 1; /* "Scene textures require mipmapping", host/wavefront.c:267-268 */
      tex[k].data   = t->data;
    }
    const Lumb200Result r = lumb200_device_add_textures(d->dev, tex, n);
    free(tex);
    DEV_TRY(r);
    d->textures_uploaded = s->num_textures;
  }
  return LUMINARY_SUCCESS;
}

static LuminaryResult upload_scene(LuminaryHost* h, const SceneSnapshot* s) {
  const uint32_t nm = s->num_materials ? s->num_materials : 1;
  const uint32_t ni = s->num_instances ? s->num_instances : 1;
  Lumb200Material* mats  = (Lumb200Material*) calloc(nm, sizeof(Lumb200Material));
  Lumb200Instance* insts = (Lumb200Instance*) calloc(ni, sizeof(Lumb200Instance));
  Lumb200Mesh* meshes    = (Lumb200Mesh*) calloc(s->num_meshes ? s->num_meshes : 1, sizeof(Lumb200Mesh));
  if (!mats || !insts || !meshes) {
    free(mats), free(insts), free(meshes);
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "out of host memory");
  }
  for (uint32_t k = 0; k < s->num_materials; k++)
    convert_material(&s->materials[k], &mats[k]);
  for (uint32_t k = 0; k < s->num_instances; k++) {
    const LuminaryInstance* in = &s->instances[k];
    insts[k].mesh_id           = in->mesh_id;
    insts[k].translation[0] = in->position.x, insts[k].translation[1] = in->position.y, insts[k].translation[2] = in->position.z;
    insts[k].rotation[0] = in->rotation.x, insts[k].rotation[1] = in->rotation.y, insts[k].rotation[2] = in->rotation.z;
    insts[k].scale[0] = in->scale.x, insts[k].scale[1] = in->scale.y, insts[k].scale[2] = in->scale.z;
    insts[k].active = in->mesh_id < s->num_meshes;
  }
  for (uint32_t k = 0; k < s->num_meshes; k++) {
    meshes[k].triangle_count     = s->meshes[k].triangle_count;
    meshes[k].vertex_buffer      = s->meshes[k].vertex_buffer;
    meshes[k].normal_buffer      = s->meshes[k].normal_buffer;
    meshes[k].uv_buffer          = s->meshes[k].uv_buffer;
    meshes[k].material_id_buffer = s->meshes[k].material_id_buffer;
  }

  /* luminance-textured emitters: their per-triangle intensity is integrated on the main device before the tree is built
   * (_light_tree_integrate, device_light.c:1952-2018); that device needs the meshes, textures and materials first */
  LuminaryResult result = LUMINARY_SUCCESS;
  float** tri_intensity = (float**) calloc(s->num_meshes ? s->num_meshes : 1, sizeof(float*));
  bool any_textured     = false;
  for (uint32_t k = 0; k < s->num_materials; k++)
    any_textured |= s->materials[k].emission_active && s->materials[k].luminance_tex != 0xFFFF;
  HostDevice* main_dev = NULL;
  for (uint32_t g = 0; g < h->num_devices && !main_dev; g++)
    if (h->devices[g].enabled && h->devices[g].dev)
      main_dev = &h->devices[g];
  if (!tri_intensity)
    result = LUMINARY_ERROR_OUT_OF_MEMORY;
  if (any_textured && main_dev && result == LUMINARY_SUCCESS) {
    set_task(h, "Integrating textured lights");
    result = upload_meshes_and_textures(main_dev, s, meshes);
    if (result == LUMINARY_SUCCESS)
      result = from_device(lumb200_device_update_materials(main_dev->dev, mats, s->num_materials));
    for (uint32_t m = 0; m < s->num_meshes && result == LUMINARY_SUCCESS; m++) {
      const uint32_t nt = s->meshes[m].triangle_count;
      uint32_t count    = 0;
      for (uint32_t t = 0; t < nt; t++) {
        const uint16_t mid = s->meshes[m].material_id_buffer[t];
        count += (mid < s->num_materials && s->materials[mid].emission_active && s->materials[mid].luminance_tex != 0xFFFF);
      }
      if (!count)
        continue;
      uint32_t* mesh_ids = (uint32_t*) malloc(sizeof(uint32_t) * count);
      uint32_t* tri_ids  = (uint32_t*) malloc(sizeof(uint32_t) * count);
      float* values      = (float*) malloc(sizeof(float) * count);
      tri_intensity[m]   = (float*) malloc(sizeof(float) * nt);
      if (!mesh_ids || !tri_ids || !values || !tri_intensity[m])
        result = LUMINARY_ERROR_OUT_OF_MEMORY;
      else {
        uint32_t n = 0;
        for (uint32_t t = 0; t < nt; t++) {
          const uint16_t mid  = s->meshes[m].material_id_buffer[t];
          tri_intensity[m][t] = 1.0f;
          if (mid < s->num_materials && s->materials[mid].emission_active && s->materials[mid].luminance_tex != 0xFFFF)
            mesh_ids[n] = m, tri_ids[n++] = t;
        }
        result = from_device(lumb200_device_compute_light_intensities(main_dev->dev, mesh_ids, tri_ids, count, values));
        for (uint32_t k = 0; k < count && result == LUMINARY_SUCCESS; k++)
          tri_intensity[m][tri_ids[k]] = values[k];
      }
      free(mesh_ids), free(tri_ids), free(values);
    }
  }

  /* light tree: built once on the CPU, uploaded to every device (device_manager.c:443-450) */
  set_task(h, "Building light tree");
  Lumb200LightTreeBuffers tree;
  memset(&tree, 0, sizeof(tree));
  if (result == LUMINARY_SUCCESS)
    result = from_device(lumb200_host_build_light_tree_textured(
      meshes, s->num_meshes, insts, s->num_instances, mats, s->num_materials, any_textured ? (const float* const*) tri_intensity : NULL, &tree));
  for (uint32_t m = 0; tri_intensity && m < s->num_meshes; m++)
    free(tri_intensity[m]);
  free(tri_intensity);

  /* internal resolution = width << supersampling (device_structs.c:21-22) */
  Lumb200Settings ds = {s->settings.width << s->settings.supersampling, s->settings.height << s->settings.supersampling, s->settings.max_ray_depth, 1};
  Lumb200Camera dc;
  memset(&dc, 0, sizeof(dc));
  dc.pos[0] = s->camera.pos.x, dc.pos[1] = s->camera.pos.y, dc.pos[2] = s->camera.pos.z;
  dc.rotation[0] = s->camera.rotation.x, dc.rotation[1] = s->camera.rotation.y, dc.rotation[2] = s->camera.rotation.z;
  dc.fov                        = s->camera.thin_lens.fov;
  dc.aperture_size              = s->camera.thin_lens.aperture_size;
  dc.object_distance            = s->camera.object_distance;
  dc.camera_scale               = s->camera.camera_scale;
  dc.russian_roulette_threshold = s->camera.russian_roulette_threshold;
  dc.aperture_shape             = (uint32_t) s->camera.aperture_shape;
  dc.aperture_blade_count       = s->camera.aperture_blade_count;
  /* device_struct_sky_convert (device_structs.c:107-172) happens inside the device library; the HDRI mode is refused there */
  Lumb200Sky dsky;
  lumb200_sky_default(&dsky);
  dsky.mode                   = (uint32_t) s->sky.mode;
  dsky.constant_color[0]      = s->sky.constant_color.r, dsky.constant_color[1] = s->sky.constant_color.g, dsky.constant_color[2] = s->sky.constant_color.b;
  dsky.geometry_offset[0]     = s->sky.geometry_offset.x, dsky.geometry_offset[1] = s->sky.geometry_offset.y, dsky.geometry_offset[2] = s->sky.geometry_offset.z;
  dsky.azimuth                = s->sky.azimuth;
  dsky.altitude               = s->sky.altitude;
  dsky.moon_azimuth           = s->sky.moon_azimuth;
  dsky.moon_altitude          = s->sky.moon_altitude;
  dsky.moon_tex_offset        = s->sky.moon_tex_offset;
  dsky.sun_strength           = s->sky.sun_strength;
  dsky.base_density           = s->sky.base_density;
  dsky.rayleigh_density       = s->sky.rayleigh_density;
  dsky.mie_density            = s->sky.mie_density;
  dsky.ozone_density          = s->sky.ozone_density;
  dsky.rayleigh_falloff       = s->sky.rayleigh_falloff;
  dsky.mie_falloff            = s->sky.mie_falloff;
  dsky.mie_diameter           = s->sky.mie_diameter;
  dsky.ground_visibility      = s->sky.ground_visibility;
  dsky.ozone_layer_thickness  = s->sky.ozone_layer_thickness;
  dsky.multiscattering_factor = s->sky.multiscattering_factor;
  dsky.stars_intensity        = s->sky.stars_intensity;
  dsky.steps                  = s->sky.steps;
  dsky.ozone_absorption       = s->sky.ozone_absorption ? 1u : 0u;
  dsky.aerial_perspective     = s->sky.aerial_perspective ? 1u : 0u;
  dsky.stars_count            = s->sky.stars_count;
  dsky.stars_seed             = s->sky.stars_seed;
  dsky.hdri_dim               = s->sky.hdri_dim;
  dsky.hdri_samples           = s->sky.hdri_samples;

  pthread_mutex_lock(&h->lock);
  const bool hdri_request = h->hdri_request;
  h->hdri_request         = false;
  pthread_mutex_unlock(&h->lock);

  set_task(h, "Updating scene");
  for (uint32_t g = 0; g < h->num_devices && result == LUMINARY_SUCCESS; g++) {
    HostDevice* d = &h->devices[g];
    if (!d->enabled || !d->dev)
      continue;
#define STEP(expr)                      \
  if (result == LUMINARY_SUCCESS)       \
    result = (expr);
    STEP(load_embedded_data(d));
    STEP(upload_meshes_and_textures(d, s, meshes));
    STEP(from_device(lumb200_device_update_materials(d->dev, mats, s->num_materials)));
    STEP(from_device(lumb200_device_update_instances(d->dev, insts, s->num_instances)));
    STEP(from_device(lumb200_device_update_settings(d->dev, &ds)));
    STEP(from_device(lumb200_device_set_shading_mode(d->dev, (uint32_t) s->settings.shading_mode)));
    STEP(from_device(lumb200_device_update_camera(d->dev, &dc)));
    STEP(from_device(lumb200_device_update_sky(d->dev, &dsky)));
    if (hdri_request && dsky.mode == 1) /* SCENE_DIRTY_FLAG_HDRI, device_manager.c:351-365; every device bakes the same table */
      STEP(from_device(lumb200_device_build_sky_hdri(d->dev)));
    {
      /* adaptive_sampler_setup (device_adaptive_sampler.c:29-56) with the values of device_manager.c: the sampler sees the
       * camera's linear exposure and tone map when it is exposure-aware */
      Lumb200AdaptiveSampling as;
      memset(&as, 0, sizeof(as));
      as.enable            = s->settings.enable_adaptive_sampling ? 1u : 0u;
      as.max_sampling_rate = s->settings.adaptive_sampling_max_sampling_rate;
      as.avg_sampling_rate = s->settings.adaptive_sampling_avg_sampling_rate;
      as.update_interval   = s->settings.adaptive_sampling_update_interval ? s->settings.adaptive_sampling_update_interval : 1;
      as.exposure_aware    = s->settings.adaptive_sampling_exposure_aware ? 1u : 0u;
      as.exposure          = expf(s->camera.exposure);
      as.tonemap           = (uint32_t) s->camera.tonemap;
      as.agx_slope         = s->camera.agx_custom_slope;
      as.agx_power         = s->camera.agx_custom_power;
      as.agx_saturation    = s->camera.agx_custom_saturation;
      as.output_mode       = (uint32_t) s->settings.adaptive_sampling_output_mode;
      STEP(from_device(lumb200_device_update_adaptive_sampling(d->dev, &as)));
    }
    if (result == LUMINARY_SUCCESS) {
      Lumb200LightTree lt = {tree.root_data, tree.root_size, tree.nodes_data, tree.nodes_size, tree.tri_handle_map, tree.num_lights};
      result              = from_device(lumb200_device_update_light_tree(d->dev, &lt));
    }
    STEP(from_device(lumb200_device_build_accel(d->dev)));
    STEP(from_device(lumb200_device_start_render(d->dev)));
#undef STEP
    d->gpu_seconds = 0.0;
    d->rays        = 0;
  }
  lumb200_host_free_light_tree(&tree);
  free(mats), free(insts), free(meshes);
  return result;
}

static uint32_t enabled_devices(LuminaryHost* h, HostDevice** list) {
  uint32_t n = 0;
  for (uint32_t g = 0; g < h->num_devices; g++)
    if (h->devices[g].enabled && h->devices[g].dev)
      list[n++] = &h->devices[g];
  return n;
}

/* folds a device's counters into the host-side totals (called before a secondary device is reset) */
static LuminaryResult harvest_stats(HostDevice* d) {
  Lumb200Stats st;
  DEV_TRY(lumb200_device_get_stats(d->dev, &st));
  d->gpu_seconds += st.render_seconds;
  d->rays += st.closest_rays + st.shadow_rays + st.light_rays;
  return LUMINARY_SUCCESS;
}

/* (re)builds the communicator when the set of enabled devices changed; without NCCL (library missing) the combine falls back to
 * peer copies */
static void ensure_comm(LuminaryHost* h, HostDevice** devs, uint32_t G) {
  uint32_t mask = 0;
  for (uint32_t g = 0; g < G; g++)
    mask |= 1u << (uint32_t) (devs[g] - h->devices);
  if (G > 1 && h->num_comms == G && h->comm_mask == mask)
    return;
  for (uint32_t k = 0; k < h->num_comms; k++)
    lumb200_comm_destroy(&h->comms[k]);
  h->num_comms = 0;
  h->comm_mask = 0;
  if (G < 2)
    return;
  Lumb200Device* list[LUM_MAX_DEVICES];
  for (uint32_t g = 0; g < G; g++)
    list[g] = devs[g]->dev;
  if (lumb200_comm_create_all(h->comms, list, G) == LUMB200_SUCCESS) {
    h->num_comms = G;
    h->comm_mask = mask;
  }
  else
    fprintf(stderr, "[luminary_b200] NCCL communicator unavailable (%s): combining devices with peer copies\n", lumb200_last_error());
}

/* device_handle_result_sharing (device/device.c:1587-1612): all samples of the secondaries onto the main device. One ncclReduce over
 * NVLink queued behind the sample passes of every device; the secondaries' planes are cleared afterwards (queued as well). */
static LuminaryResult combine_planes(LuminaryHost* h, HostDevice** devs, uint32_t G) {
  if (G < 2)
    return LUMINARY_SUCCESS;
  if (h->num_comms == G) {
    DEV_TRY(lumb200_comm_reduce_planes_all(h->comms, G, 0));
    for (uint32_t g = 1; g < G; g++)
      DEV_TRY(lumb200_device_clear_frame_planes(devs[g]->dev));
    for (uint32_t g = 0; g < G; g++)
      DEV_TRY(lumb200_device_sync(devs[g]->dev));
    return LUMINARY_SUCCESS;
  }
  for (uint32_t g = 1; g < G; g++) {
    DEV_TRY(lumb200_device_add_planes_from(devs[0]->dev, devs[g]->dev));
    DEV_TRY(lumb200_device_clear_frame_planes(devs[g]->dev));
  }
  return LUMINARY_SUCCESS;
}

static LuminaryResult produce_outputs(LuminaryHost* h, const SceneSnapshot* s, HostDevice** devs, uint32_t G, uint32_t done) {
  HostDevice* main_dev = devs[0];
  set_task(h, "Gathering results");
  LUM_TRY(combine_planes(h, devs, G));
  for (uint32_t g = 1; g < G; g++) {
    LUM_TRY(harvest_stats(devs[g]));
    if (!s->settings.enable_adaptive_sampling)
      DEV_TRY(lumb200_device_start_render(devs[g]->dev)); /* the planes of a secondary only hold samples not yet combined */
  }
  Lumb200Stats st;
  DEV_TRY(lumb200_device_get_stats(main_dev->dev, &st));
  double seconds = main_dev->gpu_seconds + st.render_seconds;
  uint64_t rays  = main_dev->rays + st.closest_rays + st.shadow_rays + st.light_rays;
  for (uint32_t g = 1; g < G; g++) {
    seconds = fmax(seconds, devs[g]->gpu_seconds);
    rays += devs[g]->rays;
  }

  Lumb200OutputParams op;
  op.exposure       = expf(s->camera.exposure); /* device_structs.c:77 */
  op.tonemap        = (uint32_t) s->camera.tonemap;
  op.agx_slope      = s->camera.agx_custom_slope;
  op.agx_power      = s->camera.agx_custom_power;
  op.agx_saturation = s->camera.agx_custom_saturation;
  op.dithering      = s->camera.dithering ? 1u : 0u;
  op.purkinje        = s->camera.purkinje ? 1u : 0u;
  op.purkinje_kappa1 = s->camera.purkinje_kappa1;
  op.purkinje_kappa2 = s->camera.purkinje_kappa2;
  op.supersampling   = s->settings.supersampling;
  op.bloom_blend     = s->camera.bloom_blend; /* device_post_update, device_post.c:187-208 */
  op.local_error_minimization = s->camera.use_local_error_minimization ? 1u : 0u;
  op.filter               = (uint32_t) s->camera.filter;
  op.use_color_correction = s->camera.use_color_correction ? 1u : 0u;
  op.color_correction[0]  = s->camera.color_correction.r; /* device_output.c:128: hue, saturation, value offsets */
  op.color_correction[1]  = s->camera.color_correction.g;
  op.color_correction[2]  = s->camera.color_correction.b;
  op.film_grain           = s->camera.film_grain;

  set_task(h, "Generating output");
  const size_t bytes = 4 * (size_t) s->settings.width * s->settings.height;
  uint8_t* image     = NULL;
  pthread_mutex_lock(&h->lock);
  h->samples_done = done;
  h->sample_time  = seconds;
  h->rays         = rays;
  bool wanted     = false;
  for (uint32_t k = 0; k < h->num_requests; k++)
    wanted |= !h->requests[k].done && h->requests[k].sample_count == done;
  wanted |= h->output_properties.enabled;
  pthread_mutex_unlock(&h->lock);
  if (!wanted)
    return LUMINARY_SUCCESS;
  image = (uint8_t*) malloc(bytes);
  if (!image)
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "out of host memory for an output of %zu bytes", bytes);
  const Lumb200Result r = lumb200_device_download_output_argb8(main_dev->dev, done, &op, image);
  if (r != LUMB200_SUCCESS) {
    free(image);
    return from_device(r);
  }

  pthread_mutex_lock(&h->lock);
  /* one output slot, shared by all requests of this sample count */
  uint32_t slot = h->num_outputs;
  for (uint32_t k = 0; k < h->num_outputs; k++)
    if (!h->outputs[k].valid) {
      slot = k;
      break;
    }
  if (slot == h->num_outputs) {
    h->outputs = (OutputSlot*) realloc(h->outputs, sizeof(OutputSlot) * (h->num_outputs + 1));
    memset(&h->outputs[h->num_outputs++], 0, sizeof(OutputSlot));
  }
  OutputSlot* o   = &h->outputs[slot];
  o->buffer       = image;
  o->width        = s->settings.width;
  o->height       = s->settings.height;
  o->time         = (float) seconds;
  o->sample_count = done;
  o->valid        = true;
  o->refs         = 1; /* the host's own reference: keeps the most recent output alive for acquire_output */
  if (h->latest_output != LUMINARY_OUTPUT_HANDLE_INVALID && h->latest_output < h->num_outputs) {
    OutputSlot* prev = &h->outputs[h->latest_output];
    if (prev->valid && --prev->refs == 0) {
      free(prev->buffer);
      memset(prev, 0, sizeof(*prev));
    }
  }
  h->latest_output = slot;
  for (uint32_t k = 0; k < h->num_requests; k++) {
    OutputRequest* q = &h->requests[k];
    if (!q->done && q->sample_count == done) {
      q->done   = true;
      q->output = slot;
      o->refs++; /* released by luminary_host_release_output */
    }
  }
  pthread_mutex_unlock(&h->lock);
  return LUMINARY_SUCCESS;
}

static LuminaryResult render_generation(LuminaryHost* h, uint32_t generation, const SceneSnapshot* s) {
  if (s->settings.width == 0 || s->settings.height == 0)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "render resolution is 0");
  HostDevice* devs[LUM_MAX_DEVICES];
  const uint32_t G = enabled_devices(h, devs);
  if (G == 0)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_DEVICE, "no enabled CUDA device");
  LUM_TRY(upload_scene(h, s));
  ensure_comm(h, devs, G);

  uint32_t done = 0;
  /* the shared adaptive schedule (adaptive_sampler_allocate_sample): stage, executions finished per stage, position in the stage */
  uint32_t as_stage = 0, as_in_stage = 0;
  uint32_t as_ex[LUMB200_ADAPTIVE_STAGES + 1] = {0, 0, 0, 0, 0};
  for (;;) {
    /* next requested sample count above what has been rendered; requests at or below it can no longer be met
     * exactly (reference: device_output.c:225-233 matches on the exact aggregated sample count) */
    pthread_mutex_lock(&h->lock);
    const bool stop = h->shutdown || h->requested_generation != generation;
    uint32_t target = 0xFFFFFFFFu;
    for (uint32_t k = 0; k < h->num_requests; k++) {
      const OutputRequest* q = &h->requests[k];
      if (!q->done && q->sample_count > done && q->sample_count < target) {
        if (q->width != s->settings.width || q->height != s->settings.height) {
          pthread_mutex_unlock(&h->lock);
          LUM_RETURN_ERROR(
            LUMINARY_ERROR_NOT_IMPLEMENTED, "output request %ux%u differs from the render resolution %ux%u (rescaled outputs are not supported)",
            q->width, q->height, s->settings.width, s->settings.height);
        }
        target = q->sample_count;
      }
    }
    if (!stop && h->output_properties.enabled && done < (1u << 20)) {
      /* recurring outputs: next stop after one more chunk of passes (or at the next requested count, whichever comes first) */
      if (h->output_properties.width != s->settings.width || h->output_properties.height != s->settings.height) {
        pthread_mutex_unlock(&h->lock);
        LUM_RETURN_ERROR(
          LUMINARY_ERROR_NOT_IMPLEMENTED, "recurring output %ux%u differs from the render resolution %ux%u (rescaled outputs are not supported)",
          h->output_properties.width, h->output_properties.height, s->settings.width, s->settings.height);
      }
      const uint32_t step = done < 8 ? 1u : LUM_PASSES_PER_CHUNK * G; /* the first previews arrive quickly */
      if (target > done + step)
        target = done + step;
    }
    if (!stop && target == 0xFFFFFFFFu) {
      /* every requested output exists: park inside the live render (the planes keep accumulating state) until a new
       * request, a new render or shutdown arrives */
      h->parked_generation = generation;
      h->busy              = false;
      h->task              = NULL;
      pthread_cond_broadcast(&h->idle);
      pthread_cond_wait(&h->wake, &h->lock);
      h->parked_generation = 0;
      h->busy              = true;
      pthread_mutex_unlock(&h->lock);
      continue;
    }
    pthread_mutex_unlock(&h->lock);
    if (stop)
      return LUMINARY_SUCCESS;

    set_task(h, "Rendering");
    while (done < target && s->settings.enable_adaptive_sampling) {
      /* adaptive sampling: "samples" are executions of the adaptive schedule, shared by all devices the way the reference's
       * devices share one AdaptiveSampler (device_adaptive_sampler.c:58-71, 330-420): the executions of a stage are dealt round
       * robin; at a stage boundary the planes are combined onto the main device (NCCL reduce), it builds the next stage and its
       * counts are broadcast, because every device must derive sample ids from identical counts. */
      const uint32_t interval = s->settings.adaptive_sampling_update_interval ? s->settings.adaptive_sampling_update_interval : 1;
      if (as_stage < 4 && as_ex[as_stage] >= (interval << as_stage)) {
        if (G > 1)
          LUM_TRY(combine_planes(h, devs, G));
        DEV_TRY(lumb200_device_set_adaptive_state(devs[0]->dev, as_stage, as_ex, NULL, 0));
        DEV_TRY(lumb200_device_build_adaptive_stage(devs[0]->dev));
        as_stage++;
        if (G > 1 && h->num_comms == G) {
          DEV_TRY(lumb200_comm_broadcast_adaptive_words_all(h->comms, G, 0));
          for (uint32_t g = 1; g < G; g++)
            DEV_TRY(lumb200_device_adopt_adaptive_stage(devs[g]->dev, as_stage, as_ex));
        }
        as_in_stage = 0;
      }
      const uint32_t g = (G > 1 && h->num_comms == G) ? as_in_stage % G : 0;
      DEV_TRY(lumb200_device_render_allocated_execution(devs[g]->dev, as_ex));
      as_ex[as_stage]++;
      as_in_stage++;
      done++;
      if ((done % LUM_PASSES_PER_CHUNK) == 0 || done == target) {
        for (uint32_t k = 0; k < G; k++)
          DEV_TRY(lumb200_device_sync(devs[k]->dev));
        pthread_mutex_lock(&h->lock);
        const bool interrupted = h->shutdown || h->requested_generation != generation;
        pthread_mutex_unlock(&h->lock);
        if (interrupted)
          return LUMINARY_SUCCESS;
      }
      if (done == target) /* the output needs the main device's sampler state to match the global schedule */
        DEV_TRY(lumb200_device_set_adaptive_state(devs[0]->dev, as_stage, as_ex, NULL, 0));
    }
    while (done < target) {
      const uint32_t end = (target - done > LUM_PASSES_PER_CHUNK * G) ? done + LUM_PASSES_PER_CHUNK * G : target;
      for (uint32_t g = 0; g < G; g++) {
        /* ids in [done, end) congruent to g modulo G */
        uint32_t first = done + ((g + G - (done % G)) % G);
        if (first >= end)
          continue;
        const uint32_t count = (end - first + G - 1) / G;
        DEV_TRY(lumb200_device_render_samples(devs[g]->dev, first, count, G));
      }
      for (uint32_t g = 0; g < G; g++)
        DEV_TRY(lumb200_device_sync(devs[g]->dev));
      done = end;
      pthread_mutex_lock(&h->lock);
      const bool interrupted = h->shutdown || h->requested_generation != generation;
      pthread_mutex_unlock(&h->lock);
      if (interrupted)
        return LUMINARY_SUCCESS;
    }
    LUM_TRY(produce_outputs(h, s, devs, (s->settings.enable_adaptive_sampling && h->num_comms != G) ? 1 : G, done));
  }
}

static void* worker_main(void* arg) {
  LuminaryHost* h = (LuminaryHost*) arg;
  pthread_mutex_lock(&h->lock);
  for (;;) {
    while (!h->shutdown && h->requested_generation == h->finished_generation)
      pthread_cond_wait(&h->wake, &h->lock);
    if (h->shutdown)
      break;
    const uint32_t generation = h->requested_generation;
    SceneSnapshot s;
    memset(&s, 0, sizeof(s));
    s.settings      = h->settings;
    s.camera        = h->camera;
    s.sky           = h->sky;
    s.num_meshes    = h->num_meshes;
    s.num_materials = h->num_materials;
    s.num_instances = h->num_instances;
    s.meshes        = (LumHostMesh*) malloc(sizeof(LumHostMesh) * (s.num_meshes ? s.num_meshes : 1));
    s.materials     = (LuminaryMaterial*) malloc(sizeof(LuminaryMaterial) * (s.num_materials ? s.num_materials : 1));
    s.instances     = (LuminaryInstance*) malloc(sizeof(LuminaryInstance) * (s.num_instances ? s.num_instances : 1));
    memcpy(s.meshes, h->meshes, sizeof(LumHostMesh) * s.num_meshes); /* triangle buffers are immutable once added */
    s.num_textures = h->num_textures;
    s.textures     = (LumHostTexture*) malloc(sizeof(LumHostTexture) * (s.num_textures ? s.num_textures : 1));
    memcpy(s.textures, h->textures, sizeof(LumHostTexture) * s.num_textures); /* so are texels */
    memcpy(s.materials, h->materials, sizeof(LuminaryMaterial) * s.num_materials);
    memcpy(s.instances, h->instances, sizeof(LuminaryInstance) * s.num_instances);
    h->busy         = true;
    h->samples_done = 0;
    h->sample_time  = 0.0;
    h->rays         = 0;
    pthread_mutex_unlock(&h->lock);

    const LuminaryResult r = render_generation(h, generation, &s);
    snapshot_free(&s);

    pthread_mutex_lock(&h->lock);
    h->busy = false;
    h->task = NULL;
    if (r != LUMINARY_SUCCESS)
      h->worker_error = r;
    /* a request that arrives while this generation was idle-exiting restarts the loop through requested_generation */
    if (h->requested_generation == generation)
      h->finished_generation = generation;
    pthread_cond_broadcast(&h->idle);
  }
  pthread_mutex_unlock(&h->lock);
  return NULL;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* API                                                                                                          */
/* ------------------------------------------------------------------------------------------------------------ */
LuminaryResult luminary_host_create(LuminaryHost** host, LuminaryHostCreateInfo info) {
  LUM_CHECK_NULL(host);
  *host           = NULL;
  LuminaryHost* h = (LuminaryHost*) calloc(1, sizeof(LuminaryHost));
  if (!h)
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "out of host memory");
  pthread_mutex_init(&h->lock, NULL);
  pthread_cond_init(&h->wake, NULL);
  pthread_cond_init(&h->idle, NULL);
  lum_settings_default(&h->settings);
  lum_camera_default(&h->camera);
  lum_sky_default(&h->sky);
  lum_inactive_entities_default(&h->ocean, &h->cloud, &h->fog, &h->particles);
  h->latest_output = LUMINARY_OUTPUT_HANDLE_INVALID;

  uint32_t count        = 0;
  const Lumb200Result r = lumb200_get_device_count(&count);
  if (r != LUMB200_SUCCESS) {
    free(h);
    return from_device(r);
  }
  if (count > LUM_MAX_DEVICES)
    count = LUM_MAX_DEVICES;
  h->num_devices = count;
  uint32_t usable = 0;
  for (uint32_t g = 0; g < count; g++) {
    HostDevice* d = &h->devices[g];
    d->cuda_index = g;
    d->enabled    = (info.device_mask >> g) & 1u;
    lumb200_get_device_properties(g, d->name, sizeof(d->name), &d->memory);
    if (d->enabled) {
      if (lumb200_device_create(&d->dev, g) != LUMB200_SUCCESS) {
        lum_log("warn", "CUDA device %u is unavailable: %s", g, lumb200_last_error());
        d->dev     = NULL;
        d->enabled = false;
      }
      else
        usable++;
    }
  }
  if (usable == 0) {
    free(h);
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_DEVICE, "no usable CUDA device (mask 0x%x, %u devices present)", info.device_mask, count);
  }
  if (pthread_create(&h->worker, NULL, worker_main, h) != 0) {
    free(h);
    LUM_RETURN_ERROR(LUMINARY_ERROR_C_STD, "failed to start the device worker thread");
  }
  h->worker_started = true;
  *host             = h;
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_destroy(LuminaryHost** host) {
  LUM_CHECK_NULL(host);
  LUM_CHECK_NULL(*host);
  LuminaryHost* h = *host;
  pthread_mutex_lock(&h->lock);
  h->shutdown = true;
  pthread_cond_broadcast(&h->wake);
  pthread_mutex_unlock(&h->lock);
  if (h->worker_started)
    pthread_join(h->worker, NULL);
  for (uint32_t c = 0; c < h->num_comms; c++)
    lumb200_comm_destroy(&h->comms[c]);
  h->num_comms = 0;
  for (uint32_t g = 0; g < h->num_devices; g++)
    if (h->devices[g].dev)
      lumb200_device_destroy(&h->devices[g].dev);
  for (uint32_t k = 0; k < h->num_meshes; k++)
    lum_host_mesh_free(&h->meshes[k]);
  for (uint32_t k = 0; k < h->num_textures; k++)
    lum_host_texture_free(&h->textures[k]);
  for (uint32_t k = 0; k < h->num_outputs; k++)
    free(h->outputs[k].buffer);
  free(h->meshes), free(h->materials), free(h->instances), free(h->requests), free(h->outputs), free(h->textures);
  pthread_mutex_destroy(&h->lock);
  pthread_cond_destroy(&h->wake);
  pthread_cond_destroy(&h->idle);
  free(h);
  *host = NULL;
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_start_new_render(LuminaryHost* h) {
  LUM_CHECK_NULL(h);
  pthread_mutex_lock(&h->lock);
  const LuminaryRendererSettings st = h->settings;
  const LuminaryCamera cam          = h->camera;
  pthread_mutex_unlock(&h->lock);
  if (st.supersampling > 2 || ((uint64_t) st.width << st.supersampling) > 16384 || ((uint64_t) st.height << st.supersampling) > 16384)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "supersampling %u of %ux%u exceeds the 16384 pixel limit per axis", st.supersampling,
                     st.width, st.height);
  if ((uint32_t) st.adaptive_sampling_output_mode > 3)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "Invalid adaptive sampling output mode.");
  if ((uint32_t) st.shading_mode >= LUMINARY_SHADING_MODE_COUNT)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "Invalid shading mode.");
  if (cam.use_physical_camera)
    LUM_RETURN_ERROR(LUMINARY_ERROR_NOT_IMPLEMENTED, "the physical camera model is not implemented by this path");
  if ((uint32_t) cam.filter >= LUMINARY_FILTER_COUNT)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "Invalid filter.");
  /* undersampling only applies to recurring (interactive) outputs - "output requests can never request undersampled outputs",
   * device.c:1300-1307 - and only to their first frames: recurring outputs here start at full resolution, requested outputs are unaffected */
  if (st.undersampling != 0 && h->output_properties.enabled)
    lum_log("info", "undersampling %u: recurring outputs start at full resolution", st.undersampling);
  pthread_mutex_lock(&h->lock);
  h->requested_generation++;
  h->worker_error = LUMINARY_SUCCESS;
  pthread_cond_broadcast(&h->wake);
  pthread_mutex_unlock(&h->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_b200_host_wait_idle(LuminaryHost* h) {
  LUM_CHECK_NULL(h);
  pthread_mutex_lock(&h->lock);
  while (!(h->parked_generation == h->requested_generation && h->requested_generation != 0)
         && (h->busy || h->requested_generation != h->finished_generation))
    pthread_cond_wait(&h->idle, &h->lock);
  const LuminaryResult r = h->worker_error;
  pthread_mutex_unlock(&h->lock);
  return r;
}

LuminaryResult luminary_b200_host_get_ray_count(LuminaryHost* h, uint64_t* rays) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(rays);
  pthread_mutex_lock(&h->lock);
  *rays = h->rays;
  pthread_mutex_unlock(&h->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_get_device_count(LuminaryHost* h, uint32_t* device_count) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(device_count);
  *device_count = h->num_devices;
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_get_device_info(LuminaryHost* h, uint32_t device_id, LuminaryDeviceInfo* info) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(info);
  if (device_id >= h->num_devices)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_DEVICE, "device %u does not exist", device_id);
  memset(info, 0, sizeof(*info));
  const HostDevice* d = &h->devices[device_id];
  HostDevice* list[LUM_MAX_DEVICES];
  const uint32_t n     = enabled_devices(h, list);
  info->is_main_device = n > 0 && list[0] == d;
  info->is_unavailable = d->dev == NULL && ((d->enabled == false) ? false : true);
  info->is_enabled     = d->enabled;
  snprintf(info->name, sizeof(info->name), "%s", d->name);
  info->memory_size = d->memory;
  if (d->dev) {
    Lumb200Stats st;
    if (lumb200_device_get_stats(d->dev, &st) == LUMB200_SUCCESS)
      info->allocated_memory_size = (size_t) st.device_bytes;
  }
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_set_device_enable(LuminaryHost* h, uint32_t device_id, bool enable) {
  LUM_CHECK_NULL(h);
  if (device_id >= h->num_devices)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_DEVICE, "device %u does not exist", device_id);
  pthread_mutex_lock(&h->lock);
  const bool busy = h->busy || h->parked_generation != 0 || h->requested_generation != h->finished_generation;
  pthread_mutex_unlock(&h->lock);
  if (busy)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "devices can only be enabled or disabled while no render is running");
  HostDevice* d = &h->devices[device_id];
  if (enable && !d->dev) {
    DEV_TRY(lumb200_device_create(&d->dev, d->cuda_index));
    d->meshes_uploaded   = 0;
    d->textures_uploaded = 0;
    d->data_loaded       = false;
  }
  d->enabled = enable;
  return LUMINARY_SUCCESS;
}

static LuminaryResult add_obj(LuminaryHost* h, const char* path, LumWavefrontArgs args) {
  pthread_mutex_lock(&h->lock);
  const uint32_t material_offset = h->num_materials;
  const uint32_t texture_offset  = h->num_textures;
  pthread_mutex_unlock(&h->lock);
  LumHostMesh mesh;
  bool has_mesh           = false;
  LuminaryMaterial* mats  = NULL;
  uint32_t num_mats       = 0;
  LumHostTexture* texs    = NULL;
  uint32_t num_texs       = 0;
  LUM_TRY(lum_wavefront_load(path, args, material_offset, texture_offset, &mesh, &has_mesh, &mats, &num_mats, &texs, &num_texs));
  if (!has_mesh) {
    free(mats);
    return LUMINARY_SUCCESS;
  }
  for (uint32_t k = 0; k < num_mats; k++) {
    if (material_offset + k > 0xFFFF) {
      lum_host_mesh_free(&mesh);
      free(mats);
      LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "more than 65536 materials");
    }
  }
  pthread_mutex_lock(&h->lock);
  h->materials = (LuminaryMaterial*) realloc(h->materials, sizeof(LuminaryMaterial) * (h->num_materials + num_mats));
  memcpy(h->materials + h->num_materials, mats, sizeof(LuminaryMaterial) * num_mats);
  h->num_materials += num_mats;
  h->meshes                  = (LumHostMesh*) realloc(h->meshes, sizeof(LumHostMesh) * (h->num_meshes + 1));
  h->meshes[h->num_meshes++] = mesh;
  if (num_texs) {
    h->textures = (LumHostTexture*) realloc(h->textures, sizeof(LumHostTexture) * (h->num_textures + num_texs));
    memcpy(h->textures + h->num_textures, texs, sizeof(LumHostTexture) * num_texs);
    h->num_textures += num_texs;
  }
  pthread_mutex_unlock(&h->lock);
  free(mats);
  free(texs);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_load_obj_file(LuminaryHost* h, LuminaryPath* path) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(path);
  LUM_CHECK_NULL(path->string);
  LumWavefrontArgs args;
  lum_wavefront_args_default(&args);
  return add_obj(h, path->string, args);
}

LuminaryResult luminary_host_load_lum_file(LuminaryHost* h, LuminaryPath* path) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(path);
  LUM_CHECK_NULL(path->string);
  LumFileContent content;
  lum_file_content_init(&content);
  LuminaryResult r = lum_file_read(path->string, &content);
  if (r != LUMINARY_SUCCESS) {
    lum_file_content_free(&content);
    return r | LUMINARY_ERROR_PROPAGATED;
  }
  const char* slash = strrchr(path->string, '/');
  for (uint32_t k = 0; k < content.num_mesh_files && r == LUMINARY_SUCCESS; k++) {
    char obj[4096];
    if (slash && content.mesh_files[k][0] != '/')
      snprintf(obj, sizeof(obj), "%.*s/%s", (int) (slash - path->string), path->string, content.mesh_files[k]);
    else
      snprintf(obj, sizeof(obj), "%s", content.mesh_files[k]);
    pthread_mutex_lock(&h->lock);
    const uint32_t mesh_id = h->num_meshes;
    pthread_mutex_unlock(&h->lock);
    r = add_obj(h, obj, content.wavefront_args);
    if (r == LUMINARY_SUCCESS) {
      /* legacy behaviour: one untransformed instance per MESHFILE (lum_v4.c:33-41) */
      LuminaryInstance inst;
      memset(&inst, 0, sizeof(inst));
      inst.mesh_id = mesh_id;
      inst.scale.x = inst.scale.y = inst.scale.z = 1.0f;
      pthread_mutex_lock(&h->lock);
      inst.id                          = h->num_instances;
      h->instances                     = (LuminaryInstance*) realloc(h->instances, sizeof(LuminaryInstance) * (h->num_instances + 1));
      h->instances[h->num_instances++] = inst;
      pthread_mutex_unlock(&h->lock);
    }
  }
  if (r == LUMINARY_SUCCESS) {
    /* a v4 file cannot express the sampling settings: keep the host's (luminary_host_set_settings semantics) */
    pthread_mutex_lock(&h->lock);
    h->settings.width         = content.settings.width;
    h->settings.height        = content.settings.height;
    h->settings.max_ray_depth = content.settings.max_ray_depth;
    h->camera                 = content.camera;
    h->sky                    = content.sky;
    pthread_mutex_unlock(&h->lock);
  }
  lum_file_content_free(&content);
  return r;
}

LuminaryResult luminary_host_get_current_sample_time(LuminaryHost* h, double* time) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(time);
  pthread_mutex_lock(&h->lock);
  *time = h->sample_time;
  pthread_mutex_unlock(&h->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_get_num_queue_workers(const LuminaryHost* h, uint32_t* n) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(n);
  *n = 1;
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_get_queue_worker_name(const LuminaryHost* h, uint32_t id, const char** string) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(string);
  *string = (id == 0 && h->worker_started) ? "Device Manager" : NULL;
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_get_queue_worker_string(const LuminaryHost* h, uint32_t id, const char** string) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(string);
  *string = (id == 0) ? h->task : NULL;
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_get_queue_worker_time(const LuminaryHost* h, uint32_t id, double* time) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(time);
  *time = 0.0;
  if (id == 0 && h->task)
    *time = now_seconds() - ((double) h->task_start.tv_sec + 1e-9 * (double) h->task_start.tv_nsec);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_set_output_properties(LuminaryHost* h, LuminaryOutputProperties properties) {
  LUM_CHECK_NULL(h);
  pthread_mutex_lock(&h->lock);
  h->output_properties = properties;
  pthread_cond_broadcast(&h->wake); /* a parked render resumes when recurring outputs are switched on */
  pthread_mutex_unlock(&h->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_request_output(LuminaryHost* h, LuminaryOutputRequestProperties p, LuminaryOutputPromiseHandle* handle) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(handle);
  if (p.sample_count == 0 || p.sample_count > (1u << 20))
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "sample count %u is outside 1..2^20", p.sample_count);
  pthread_mutex_lock(&h->lock);
  h->requests = (OutputRequest*) realloc(h->requests, sizeof(OutputRequest) * (h->num_requests + 1));
  OutputRequest* q = &h->requests[h->num_requests];
  q->sample_count  = p.sample_count;
  q->width         = p.width;
  q->height        = p.height;
  q->done          = false;
  q->output        = LUMINARY_OUTPUT_HANDLE_INVALID;
  *handle          = h->num_requests++;
  if (h->parked_generation != 0) {
    /* a parked worker re-reads the request list and keeps accumulating; it is busy from this moment on */
    h->parked_generation = 0;
    h->busy              = true;
  }
  pthread_cond_broadcast(&h->wake);
  pthread_mutex_unlock(&h->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_try_await_output(LuminaryHost* h, LuminaryOutputPromiseHandle handle, LuminaryOutputHandle* output) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(output);
  *output = LUMINARY_OUTPUT_HANDLE_INVALID;
  pthread_mutex_lock(&h->lock);
  LuminaryResult r = LUMINARY_SUCCESS;
  if (handle >= h->num_requests)
    r = LUMINARY_ERROR_INVALID_API_ARGUMENT;
  else if (h->requests[handle].done)
    *output = h->requests[handle].output;
  else if (h->worker_error != LUMINARY_SUCCESS)
    r = h->worker_error | LUMINARY_ERROR_PROPAGATED; /* the render that should have produced it failed */
  pthread_mutex_unlock(&h->lock);
  if (r == LUMINARY_ERROR_INVALID_API_ARGUMENT)
    LUM_RETURN_ERROR(r, "output promise %u does not exist", handle);
  return r;
}

LuminaryResult luminary_host_acquire_output(LuminaryHost* h, LuminaryOutputHandle* output) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(output);
  pthread_mutex_lock(&h->lock);
  *output = h->latest_output;
  if (*output != LUMINARY_OUTPUT_HANDLE_INVALID)
    h->outputs[*output].refs++;
  pthread_mutex_unlock(&h->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_get_image(LuminaryHost* h, LuminaryOutputHandle output, LuminaryImage* image) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(image);
  pthread_mutex_lock(&h->lock);
  const bool ok = output < h->num_outputs && h->outputs[output].valid;
  if (ok) {
    const OutputSlot* o           = &h->outputs[output];
    image->buffer                 = o->buffer;
    image->width                  = o->width;
    image->height                 = o->height;
    image->ld                     = o->width;
    image->meta_data.time         = o->time;
    image->meta_data.sample_count = o->sample_count;
  }
  pthread_mutex_unlock(&h->lock);
  if (!ok)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "output handle %u is not valid", output);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_release_output(LuminaryHost* h, LuminaryOutputHandle output) {
  LUM_CHECK_NULL(h);
  pthread_mutex_lock(&h->lock);
  const bool ok = output < h->num_outputs && h->outputs[output].valid && h->outputs[output].refs > 0;
  if (ok) {
    OutputSlot* o = &h->outputs[output];
    if (--o->refs == 0) {
      free(o->buffer);
      memset(o, 0, sizeof(*o));
      if (h->latest_output == output)
        h->latest_output = LUMINARY_OUTPUT_HANDLE_INVALID;
    }
  }
  pthread_mutex_unlock(&h->lock);
  if (!ok)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "output handle %u is not valid", output);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_save_png(LuminaryHost* h, LuminaryOutputHandle handle, LuminaryPath* path) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(path);
  LUM_CHECK_NULL(path->string);
  LuminaryImage image;
  LUM_TRY(luminary_host_get_image(h, handle, &image));
  return lum_png_write_argb8(path->string, image.buffer, image.width, image.height, image.ld);
}

#define LOCKED_GET(field, dst)       \
  do {                               \
    LUM_CHECK_NULL(h);               \
    LUM_CHECK_NULL(dst);             \
    pthread_mutex_lock(&h->lock);    \
    *(dst) = h->field;               \
    pthread_mutex_unlock(&h->lock);  \
    return LUMINARY_SUCCESS;         \
  } while (0)
#define LOCKED_SET(field, src)       \
  do {                               \
    LUM_CHECK_NULL(h);               \
    LUM_CHECK_NULL(src);             \
    pthread_mutex_lock(&h->lock);    \
    h->field = *(src);               \
    pthread_mutex_unlock(&h->lock);  \
    return LUMINARY_SUCCESS;         \
  } while (0)

LuminaryResult luminary_host_get_settings(LuminaryHost* h, LuminaryRendererSettings* settings) { LOCKED_GET(settings, settings); }
LuminaryResult luminary_host_set_settings(LuminaryHost* h, const LuminaryRendererSettings* settings) {
  LUM_CHECK_NULL(settings);
  if (settings->width == 0 || settings->height == 0 || settings->width > 16384 || settings->height > 16384)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "resolution %ux%u is outside 1..16384", settings->width, settings->height);
  LOCKED_SET(settings, settings);
}
LuminaryResult luminary_host_get_camera(LuminaryHost* h, LuminaryCamera* camera) { LOCKED_GET(camera, camera); }
LuminaryResult luminary_host_set_camera(LuminaryHost* h, const LuminaryCamera* camera) { LOCKED_SET(camera, camera); }
LuminaryResult luminary_host_get_sky(LuminaryHost* h, LuminarySky* sky) { LOCKED_GET(sky, sky); }
LuminaryResult luminary_host_set_sky(LuminaryHost* h, const LuminarySky* sky) {
  LUM_CHECK_NULL(sky);
  if ((uint32_t) sky->mode >= LUMINARY_SKY_MODE_COUNT)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Invalid sky mode.");
  LOCKED_SET(sky, sky);
}

/* Entities outside the path (ocean, clouds, fog, particles): real layouts and the reference's defaults, readable and writable as long
 * as they stay inactive. Activating one answers LUMINARY_ERROR_NOT_IMPLEMENTED - never a silently different image. */
#define INACTIVE_ENTITY(type, field, label)                                                                                  \
  LuminaryResult luminary_host_get_##field(LuminaryHost* h, type* arg) {                                                     \
    LUM_CHECK_NULL(h);                                                                                                       \
    LUM_CHECK_NULL(arg);                                                                                                     \
    pthread_mutex_lock(&h->lock);                                                                                            \
    *arg = h->field;                                                                                                         \
    pthread_mutex_unlock(&h->lock);                                                                                          \
    return LUMINARY_SUCCESS;                                                                                                 \
  }                                                                                                                          \
  LuminaryResult luminary_host_set_##field(LuminaryHost* h, const type* arg) {                                               \
    LUM_CHECK_NULL(h);                                                                                                       \
    LUM_CHECK_NULL(arg);                                                                                                     \
    if (arg->active)                                                                                                         \
      LUM_RETURN_ERROR(LUMINARY_ERROR_NOT_IMPLEMENTED, label " is outside the path served by luminary_b200 and cannot be activated"); \
    pthread_mutex_lock(&h->lock);                                                                                            \
    h->field = *arg;                                                                                                         \
    pthread_mutex_unlock(&h->lock);                                                                                          \
    return LUMINARY_SUCCESS;                                                                                                 \
  }
INACTIVE_ENTITY(LuminaryOcean, ocean, "the ocean entity")
INACTIVE_ENTITY(LuminaryCloud, cloud, "the cloud entity")
INACTIVE_ENTITY(LuminaryFog, fog, "the fog entity")
INACTIVE_ENTITY(LuminaryParticles, particles, "the particles entity")

/* device_get_gbuffer_meta of the main device (host.c:997-1016): what the primary ray of pixel (x, y) hits. Served while the render
 * worker is parked inside a live render (all requested outputs produced); otherwise the query reports itself invalid, like the
 * reference before its first G-buffer pass. */
LuminaryResult luminary_host_get_pixel_info(LuminaryHost* h, uint16_t x, uint16_t y, LuminaryPixelQueryResult* result) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(result);
  memset(result, 0, sizeof(*result));
  result->instance_id = 0xFFFFFFFFu;
  result->material_id = 0xFFFFu;
  result->depth       = -1.0f; /* DEPTH_INVALID */
  pthread_mutex_lock(&h->lock);
  LuminaryResult r = LUMINARY_SUCCESS;
  if (h->parked_generation != 0 && !h->busy) {
    HostDevice* list[LUM_MAX_DEVICES];
    const uint32_t G = enabled_devices(h, list);
    if (G > 0 && x < h->settings.width && y < h->settings.height) {
      uint32_t instance = 0, tri = 0;
      float depth = 0.0f, ray[3] = {0.0f, 0.0f, 0.0f};
      const Lumb200Result dr = lumb200_device_query_pixel(list[0]->dev, x, y, 0, &instance, &tri, &depth, ray);
      if (dr != LUMB200_SUCCESS)
        r = from_device(dr);
      else {
        result->depth = depth;
        if (instance != 0xFFFFFFFFu && instance < h->num_instances) {
          const LumHostMesh* m = &h->meshes[h->instances[instance].mesh_id];
          result->instance_id  = instance;
          if (tri < m->triangle_count)
            result->material_id = m->material_id_buffer[tri];
          /* rel_hit_pos = ray x depth, stored as bfloat16 by the reference (optix_kernel_raytrace.cu:55-67) */
          const float rel[3] = {ray[0] * depth, ray[1] * depth, ray[2] * depth};
          float out[3];
          for (int k = 0; k < 3; k++) {
            uint32_t bits;
            memcpy(&bits, &rel[k], 4);
            bits &= 0xFFFF0000u;
            memcpy(&out[k], &bits, 4);
          }
          result->rel_hit_pos.x = out[0], result->rel_hit_pos.y = out[1], result->rel_hit_pos.z = out[2];
        }
        result->pixel_query_is_valid = (result->depth != -1.0f) || (result->instance_id != 0xFFFFFFFFu) || (result->material_id != 0xFFFFu);
      }
    }
  }
  pthread_mutex_unlock(&h->lock);
  return r;
}

/* host.c:1077-1084: marks the HDRI dirty; it is baked from the camera position of the next render (a changed sky re-bakes by itself) */
LuminaryResult luminary_host_request_sky_hdri_build(LuminaryHost* h) {
  LUM_CHECK_NULL(h);
  pthread_mutex_lock(&h->lock);
  h->hdri_request = true;
  pthread_mutex_unlock(&h->lock);
  return LUMINARY_SUCCESS;
}

/* host.h:39-40: interactive hot-plug of a device; here the enabled set takes effect at the next luminary_host_start_new_render */
LuminaryResult luminary_host_start_device(LuminaryHost* h, uint32_t index) { return luminary_host_set_device_enable(h, index, true); }
LuminaryResult luminary_host_shutdown_device(LuminaryHost* h, uint32_t index) { return luminary_host_set_device_enable(h, index, false); }

LuminaryResult luminary_host_get_material(LuminaryHost* h, uint16_t id, LuminaryMaterial* material) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(material);
  pthread_mutex_lock(&h->lock);
  const bool ok = id < h->num_materials;
  if (ok)
    *material = h->materials[id];
  pthread_mutex_unlock(&h->lock);
  if (!ok)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "material %u does not exist", id);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_set_material(LuminaryHost* h, uint16_t id, const LuminaryMaterial* material) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(material);
  pthread_mutex_lock(&h->lock);
  const bool ok = id < h->num_materials;
  if (ok) {
    h->materials[id]    = *material;
    h->materials[id].id = id;
  }
  pthread_mutex_unlock(&h->lock);
  if (!ok)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "material %u does not exist", id);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_get_instance(LuminaryHost* h, uint32_t id, LuminaryInstance* instance) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(instance);
  pthread_mutex_lock(&h->lock);
  const bool ok = id < h->num_instances;
  if (ok)
    *instance = h->instances[id];
  pthread_mutex_unlock(&h->lock);
  if (!ok)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "instance %u does not exist", id);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_set_instance(LuminaryHost* h, const LuminaryInstance* instance) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(instance);
  pthread_mutex_lock(&h->lock);
  const bool ok = instance->id < h->num_instances;
  if (ok)
    h->instances[instance->id] = *instance;
  pthread_mutex_unlock(&h->lock);
  if (!ok)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "instance %u does not exist", instance->id);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_new_instance(LuminaryHost* h, LuminaryInstance* instance) {
  LUM_CHECK_NULL(h);
  LUM_CHECK_NULL(instance);
  memset(instance, 0, sizeof(*instance));
  instance->scale.x = instance->scale.y = instance->scale.z = 1.0f;
  pthread_mutex_lock(&h->lock);
  instance->id                     = h->num_instances;
  h->instances                     = (LuminaryInstance*) realloc(h->instances, sizeof(LuminaryInstance) * (h->num_instances + 1));
  h->instances[h->num_instances++] = *instance;
  pthread_mutex_unlock(&h->lock);
  return LUMINARY_SUCCESS;
}

LuminaryResult luminary_host_get_num_meshes(LuminaryHost* h, uint32_t* n) { LOCKED_GET(num_meshes, n); }
LuminaryResult luminary_host_get_num_materials(LuminaryHost* h, uint32_t* n) { LOCKED_GET(num_materials, n); }
LuminaryResult luminary_host_get_num_instances(LuminaryHost* h, uint32_t* n) { LOCKED_GET(num_instances, n); }
