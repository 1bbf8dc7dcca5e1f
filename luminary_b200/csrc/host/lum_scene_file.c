/* lum_scene_file.c - entity defaults and the *.lum version 4 scene file reader of the host layer.
 *
 * File grammar (reference host/lum.c:47-128, host/lum_v4.c): first line "Luminary", second "VERSION n"; then
 * lines "SECTION KEY_____ value(s)" with 8-character keys. Only the entities on the path are kept (GENERAL,
 * MATERIAL legacy switches, CAMERA, SKY); CLOUD / FOG / OCEAN / PARTICLE / TOY lines are recognised and skipped,
 * unknown keys produce the reference's warning. Version 5 files are rejected: the reference itself parses them to a
 * binary and discards it (host/lum_v5.c:42-43). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lum_host_internal.h"

void lum_settings_default(LuminaryRendererSettings* s) { /* settings.c:6-28; see luminary.h for the two deviations */
  memset(s, 0, sizeof(*s));
  s->width                               = 2560;
  s->height                              = 1440;
  s->max_ray_depth                       = 4;
  s->bridge_max_num_vertices             = 15;
  s->undersampling                       = 2; /* settings.c:13; only lowers the resolution of the first RECURRING outputs (device.c:1300-1307) */
  s->supersampling                       = 1; /* 2x2 internal resolution, as in the reference */
  s->enable_adaptive_sampling            = true;
  s->adaptive_sampling_max_sampling_rate = 256;
  s->adaptive_sampling_avg_sampling_rate = 2;
  s->adaptive_sampling_update_interval   = 64;
  s->adaptive_sampling_exposure_aware    = true;
  s->adaptive_sampling_output_mode       = LUMINARY_ADAPTIVE_SAMPLING_OUTPUT_MODE_BEAUTY;
  s->shading_mode                        = LUMINARY_SHADING_MODE_DEFAULT;
  s->region_x                            = 0.0f;
  s->region_y                            = 0.0f;
  s->region_width                        = 1.0f;
  s->region_height                       = 1.0f;
}

void lum_camera_default(LuminaryCamera* c) { /* camera.c:7-66 */
  memset(c, 0, sizeof(*c));
  c->aperture_shape             = LUMINARY_APERTURE_ROUND;
  c->aperture_blade_count       = 7;
  c->exposure                   = 0.0f;
  c->bloom_blend                = 0.01f;
  c->dithering                  = true;
  c->tonemap                    = LUMINARY_TONEMAP_AGX;
  c->agx_custom_slope           = 1.0f;
  c->agx_custom_power           = 1.0f;
  c->agx_custom_saturation      = 1.0f;
  c->filter                     = LUMINARY_FILTER_NONE;
  c->wasd_speed                 = 1.0f;
  c->mouse_speed                = 1.0f;
  c->smoothing_factor           = 0.1f;
  c->purkinje                   = true;
  c->purkinje_kappa1            = 0.2f;
  c->purkinje_kappa2            = 0.29f;
  c->russian_roulette_threshold = 0.1f;
  c->camera_scale               = 1.0f;
  c->object_distance            = 1.0f;
  c->thin_lens.fov              = 1.0f;
  c->thin_lens.aperture_size    = 0.0f;
  /* physical camera block (Canon 50 mm preset of the reference); not used by this path */
  const float last_vertex         = 88.18f * (50.53f / 100.0f);
  c->physical.focal_length          = 50.53f;
  c->physical.front_focal_point     = last_vertex + 22.69f;
  c->physical.back_focal_point      = last_vertex - 65.18f;
  c->physical.front_principal_point = last_vertex - 27.84f;
  c->physical.back_principal_point  = last_vertex - 14.65f;
  c->physical.aperture_point        = last_vertex - 28.02f;
  c->physical.aperture_diameter     = 21.411f;
  c->physical.exit_pupil_diameter   = 28.0f;
  c->physical.image_plane_distance  = 65.18f - last_vertex;
  c->physical.sensor_width          = 20.0f;
}

void lum_inactive_entities_default(LuminaryOcean* o, LuminaryCloud* c, LuminaryFog* f, LuminaryParticles* p) {
  memset(o, 0, sizeof(*o)); /* ocean.c:6-22 */
  o->amplitude = 0.2f, o->frequency = 0.12f, o->refractive_index = 1.333f, o->water_type = LUMINARY_JERLOV_WATER_TYPE_IB;
  o->caustics_ris_sample_count = 32, o->caustics_domain_scale = 0.5f;
  memset(c, 0, sizeof(*c)); /* cloud.c:6-53 */
  c->steps = 96, c->shadow_steps = 8, c->atmosphere_scattering = true, c->seed = 1;
  c->noise_shape_scale = c->noise_detail_scale = c->noise_weather_scale = 1.0f;
  c->octaves = 9, c->droplet_diameter = 25.0f, c->density = 1.0f;
  const LuminaryCloudLayer low = {true, 5.0f, 1.5f, 1.0f, 0.0f, 1.0f, 0.0f, 2.5f, 0.0f};
  const LuminaryCloudLayer mid = {true, 6.0f, 5.5f, 1.0f, 0.0f, 1.0f, 0.0f, 2.5f, 0.0f};
  const LuminaryCloudLayer top = {true, 8.0f, 7.95f, 1.0f, 0.0f, 1.0f, 0.0f, 1.0f, 0.0f};
  c->low = low, c->mid = mid, c->top = top;
  memset(f, 0, sizeof(*f)); /* fog.c:6-16 */
  f->density = 1.0f, f->droplet_diameter = 10.0f, f->height = 500.0f, f->dist = 500.0f;
  memset(p, 0, sizeof(*p)); /* particles.c:6-24 */
  p->scale = 10.0f, p->albedo.r = p->albedo.g = p->albedo.b = 1.0f, p->direction_altitude = 1.234f;
  p->phase_diameter = 50.0f, p->count = 8192, p->size = 1.0f, p->size_variation = 0.1f;
}

void lum_sky_default(LuminarySky* s) { /* sky.c:5-41 */
  memset(s, 0, sizeof(*s));
  s->geometry_offset.y      = 0.1f;
  s->altitude               = 0.5f;
  s->azimuth                = 3.141f;
  s->moon_altitude          = -0.5f;
  s->moon_azimuth           = 0.0f;
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
  s->ozone_absorption       = true;
  s->aerial_perspective     = false;
  s->hdri_dim               = 2048;
  s->hdri_samples           = 32;
  s->stars_seed             = 0;
  s->stars_count            = 10000;
  s->stars_intensity        = 1.0f;
  s->constant_color.r       = 1.0f;
  s->constant_color.g       = 1.0f;
  s->constant_color.b       = 1.0f;
  s->mode                   = LUMINARY_SKY_MODE_DEFAULT;
}

void lum_material_default(LuminaryMaterial* m) { /* material.c:5-29 */
  memset(m, 0, sizeof(*m));
  m->base_substrate           = LUMINARY_MATERIAL_BASE_SUBSTRATE_OPAQUE;
  m->albedo.r                 = 0.9f;
  m->albedo.g                 = 0.9f;
  m->albedo.b                 = 0.9f;
  m->albedo.a                 = 0.9f;
  m->emission_scale           = 1.0f;
  m->roughness                = 0.7f;
  m->roughness_clamp          = 0.25f;
  m->refraction_index         = 1.0f;
  m->normal_map_is_compressed = true;
  m->albedo_tex = m->luminance_tex = m->roughness_tex = m->metallic_tex = m->normal_tex = 0xFFFF;
}

void lum_file_content_init(LumFileContent* c) {
  memset(c, 0, sizeof(*c));
  lum_settings_default(&c->settings);
  lum_camera_default(&c->camera);
  lum_sky_default(&c->sky);
  lum_wavefront_args_default(&c->wavefront_args);
}

void lum_file_content_free(LumFileContent* c) {
  for (uint32_t k = 0; k < c->num_mesh_files; k++)
    free(c->mesh_files[k]);
  free(c->mesh_files);
  c->mesh_files     = NULL;
  c->num_mesh_files = 0;
}

static bool key_is(const char* line, const char* key) { return strncmp(line, key, 8) == 0; }

static void parse_general(LumFileContent* c, const char* key, const char* value) {
  if (key_is(key, "MESHFILE")) {
    char name[4096];
    if (sscanf(value, "%4095s", name) == 1) {
      char** grown = (char**) realloc(c->mesh_files, sizeof(char*) * (c->num_mesh_files + 1));
      char* copy   = grown ? strdup(name) : NULL;
      if (grown)
        c->mesh_files = grown;
      if (copy)
        c->mesh_files[c->num_mesh_files++] = copy;
      else
        lum_log("error", "out of host memory while reading the mesh file list");
    }
  }
  else if (key_is(key, "WIDTH___"))
    sscanf(value, "%u", &c->settings.width);
  else if (key_is(key, "HEIGHT__"))
    sscanf(value, "%u", &c->settings.height);
  else if (key_is(key, "BOUNCES_"))
    sscanf(value, "%u", &c->settings.max_ray_depth);
  else if (key_is(key, "NUMLIGHT")) {
  } /* legacy */
  else
    lum_log("warn", "%8.8s is not a valid GENERAL setting.", key);
}

static void parse_material(LumFileContent* c, bool* legacy_thin_walled, const char* key, const char* value) {
  uint32_t b = 0;
  if (key_is(key, "EMISSION"))
    sscanf(value, "%f", &c->wavefront_args.emission_scale);
  else if (key_is(key, "COLORTRA")) {
    sscanf(value, "%u", &b);
    c->wavefront_args.force_transparency_cutout = b != 0;
  }
  else if (key_is(key, "IORSHADO")) {
    sscanf(value, "%u", &b);
    *legacy_thin_walled = b != 0;
  }
  else if (key_is(key, "INTERTRO")) {
    sscanf(value, "%u", &b);
    c->wavefront_args.legacy_smoothness = b != 0;
  }
  else
    lum_log("warn", "%8.8s is not a valid MATERIAL setting.", key);
}

static void parse_camera(LumFileContent* c, bool* force_no_bloom, const char* key, const char* value) {
  LuminaryCamera* cam = &c->camera;
  uint32_t b          = 0;
  if (key_is(key, "POSITION"))
    sscanf(value, "%f %f %f", &cam->pos.x, &cam->pos.y, &cam->pos.z);
  else if (key_is(key, "ROTATION"))
    sscanf(value, "%f %f %f", &cam->rotation.x, &cam->rotation.y, &cam->rotation.z);
  else if (key_is(key, "FOV_____"))
    sscanf(value, "%f", &cam->thin_lens.fov);
  else if (key_is(key, "FOCALLEN"))
    sscanf(value, "%f", &cam->object_distance);
  else if (key_is(key, "APERTURE"))
    sscanf(value, "%f", &cam->thin_lens.aperture_size);
  else if (key_is(key, "APESHAPE")) {
    sscanf(value, "%u", &b);
    cam->aperture_shape = (LuminaryApertureShape) b;
  }
  else if (key_is(key, "APEBLACO"))
    sscanf(value, "%u", &cam->aperture_blade_count);
  else if (key_is(key, "EXPOSURE")) {
    sscanf(value, "%f", &cam->exposure);
    cam->exposure = logf(cam->exposure); /* legacy linear -> exponential scale, lum_v4.c:183 */
  }
  else if (key_is(key, "BLOOM___")) {
    sscanf(value, "%u", &b);
    *force_no_bloom = b == 0;
  }
  else if (key_is(key, "BLOOMBLE"))
    sscanf(value, "%f", &cam->bloom_blend);
  else if (key_is(key, "DITHER__")) {
    sscanf(value, "%u", &b);
    cam->dithering = b != 0;
  }
  else if (key_is(key, "TONEMAP_")) {
    sscanf(value, "%u", &b);
    cam->tonemap = (LuminaryToneMap) b;
  }
  else if (key_is(key, "AGXSLOPE"))
    sscanf(value, "%f", &cam->agx_custom_slope);
  else if (key_is(key, "AGXPOWER"))
    sscanf(value, "%f", &cam->agx_custom_power);
  else if (key_is(key, "AGXSATUR"))
    sscanf(value, "%f", &cam->agx_custom_saturation);
  else if (key_is(key, "FILTER__")) {
    sscanf(value, "%u", &b);
    cam->filter = (LuminaryFilter) b;
  }
  else if (key_is(key, "PURKINJE")) {
    sscanf(value, "%u", &b);
    cam->purkinje = b != 0;
  }
  else if (key_is(key, "RUSSIANR"))
    sscanf(value, "%f", &cam->russian_roulette_threshold);
  else if (key_is(key, "FILMGRAI"))
    sscanf(value, "%f", &cam->film_grain);
  else if (key_is(key, "AUTOEXP_") || key_is(key, "MINEXPOS") || key_is(key, "MAXEXPOS") || key_is(key, "LENSFLAR") || key_is(key, "LENSFTHR")
           || key_is(key, "FIREFLYC")) {
  } /* legacy keys the reference accepts silently */
  else
    lum_log("warn", "%8.8s is not a valid CAMERA setting.", key);
}

static void parse_sky(LumFileContent* c, const char* key, const char* value) {
  LuminarySky* s = &c->sky;
  uint32_t b     = 0;
  if (key_is(key, "MODE____")) {
    sscanf(value, "%u", &b);
    s->mode = (LuminarySkyMode) b;
  }
  else if (key_is(key, "COLORCON"))
    sscanf(value, "%f %f %f", &s->constant_color.r, &s->constant_color.g, &s->constant_color.b);
  else if (key_is(key, "OFFSET__"))
    sscanf(value, "%f %f %f", &s->geometry_offset.x, &s->geometry_offset.y, &s->geometry_offset.z);
  else if (key_is(key, "AZIMUTH_"))
    sscanf(value, "%f", &s->azimuth);
  else if (key_is(key, "ALTITUDE"))
    sscanf(value, "%f", &s->altitude);
  else if (key_is(key, "MOONALTI"))
    sscanf(value, "%f", &s->moon_altitude);
  else if (key_is(key, "MOONAZIM"))
    sscanf(value, "%f", &s->moon_azimuth);
  else if (key_is(key, "MOONTEXO"))
    sscanf(value, "%f", &s->moon_tex_offset);
  else if (key_is(key, "SUNSTREN"))
    sscanf(value, "%f", &s->sun_strength);
  else if (key_is(key, "DENSITY_"))
    sscanf(value, "%f", &s->base_density);
  else if (key_is(key, "STEPS___"))
    sscanf(value, "%u", &s->steps);
  else if (key_is(key, "STARSEED"))
    sscanf(value, "%u", &s->stars_seed);
  else if (key_is(key, "STARINTE"))
    sscanf(value, "%f", &s->stars_intensity);
  else if (key_is(key, "STARNUM_"))
    sscanf(value, "%u", &s->stars_count);
  else if (key_is(key, "OZONEABS")) {
    sscanf(value, "%u", &b);
    s->ozone_absorption = b != 0;
  }
  else if (key_is(key, "AERIALPE")) {
    sscanf(value, "%u", &b);
    s->aerial_perspective = b != 0;
  }
  else if (key_is(key, "HDRIDIM_")) {
    sscanf(value, "%u", &s->hdri_dim);
    s->hdri_dim = s->hdri_dim ? s->hdri_dim : 1;
  }
  else if (key_is(key, "HDRISAMP"))
    sscanf(value, "%u", &s->hdri_samples);
  else if (key_is(key, "RAYLEDEN"))
    sscanf(value, "%f", &s->rayleigh_density);
  else if (key_is(key, "MIEDENSI"))
    sscanf(value, "%f", &s->mie_density);
  else if (key_is(key, "OZONEDEN"))
    sscanf(value, "%f", &s->ozone_density);
  else if (key_is(key, "RAYLEFAL"))
    sscanf(value, "%f", &s->rayleigh_falloff);
  else if (key_is(key, "MIEFALLO"))
    sscanf(value, "%f", &s->mie_falloff);
  else if (key_is(key, "GROUNDVI"))
    sscanf(value, "%f", &s->ground_visibility);
  else if (key_is(key, "DIAMETER"))
    sscanf(value, "%f", &s->mie_diameter);
  else if (key_is(key, "OZONETHI"))
    sscanf(value, "%f", &s->ozone_layer_thickness);
  else if (key_is(key, "MSFACTOR"))
    sscanf(value, "%f", &s->multiscattering_factor);
  else if (key_is(key, "HDRIMIPB") || key_is(key, "HDRIORIG")) {
  } /* HDRI bake parameters: accepted, the HDRI mode is not on the path */
  else
    lum_log("warn", "%8.8s is not a valid SKY setting.", key);
}

LuminaryResult lum_file_read(const char* path, LumFileContent* c) {
  LUM_CHECK_NULL(path);
  LUM_CHECK_NULL(c);
  FILE* f = fopen(path, "rb");
  if (!f)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "File %s could not be opened.", path);
  char line[4096];
  if (!fgets(line, sizeof(line), f) || strncmp(line, "Luminary", 8) != 0) {
    fclose(f);
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "File is not a Luminary file.");
  }
  uint32_t version = 0;
  if (!fgets(line, sizeof(line), f) || !(line[0] == 'v' || line[0] == 'V') || sscanf(line, "%*s %u", &version) != 1) {
    fclose(f);
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Luminary file has no version information.");
  }
  if (version < 4) {
    fclose(f);
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Luminary file is version %u but minimum supported version is 4.", version);
  }
  if (version == 5) {
    fclose(f);
    LUM_RETURN_ERROR(LUMINARY_ERROR_NOT_IMPLEMENTED, "Luminary file version 5 is not supported by this path (the reference discards its content too).");
  }
  if (version > 5) {
    fclose(f);
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Luminary file is version %u is unknown. Current supported range [4, 4].", version);
  }

  bool force_no_bloom = false, legacy_thin_walled = false;
  c->camera.use_physical_camera = false; /* legacy scenes cannot use the physical camera, lum_v4.c:688 */
  while (fgets(line, sizeof(line), f)) {
    const size_t len = strlen(line);
    if (line[0] == '#' || line[0] == '\n' || line[0] == '\r')
      continue;
    /* "SECTION KEY_____ value": the key starts after the first blank, the value 9 characters later */
    const char* sp = strchr(line, ' ');
    if (!sp || (size_t) (sp - line) + 9 > len) {
      lum_log("warn", "Scene file contains unknown line!\n Content: %s", line);
      continue;
    }
    const char* key   = sp + 1;
    const char* value = (strlen(key) > 9) ? key + 9 : "";
    if (line[0] == 'G')
      parse_general(c, key, value);
    else if (line[0] == 'M')
      parse_material(c, &legacy_thin_walled, key, value);
    else if (line[0] == 'C' && line[1] == 'A')
      parse_camera(c, &force_no_bloom, key, value);
    else if (line[0] == 'S')
      parse_sky(c, key, value);
    else if ((line[0] == 'C' && line[1] == 'L') || line[0] == 'F' || line[0] == 'O' || line[0] == 'P' || line[0] == 'T') {
    } /* clouds, fog, ocean, particles, legacy toy: entities outside the path */
    else
      lum_log("warn", "Scene file contains unknown line!\n Content: %s", line);
  }
  fclose(f);
  if (force_no_bloom)
    c->camera.bloom_blend = 0.0f;
  (void) legacy_thin_walled;
  c->wavefront_args.force_bidirectional_emission = true; /* lum_v4.c:752 */
  return LUMINARY_SUCCESS;
}
