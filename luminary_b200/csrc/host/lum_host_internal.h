/* lum_host_internal.h - shared declarations of the C host layer behind include/luminary/luminary.h.
 *
 * The host layer keeps the reference's division of labour: the application thread only edits a
 * caller-side copy of the scene, a worker thread owns the devices and calls the C ABI of include/lumb200.h. */
#ifndef LUM_HOST_INTERNAL_H
#define LUM_HOST_INTERNAL_H

#include <luminary/luminary.h>
#include <stdarg.h>

#include "../../../include/lumb200.h"

/* error plumbing: reference internal_error.h (__RETURN_ERROR / __FAILURE_HANDLE) */
void lum_set_error(const char* fmt, ...);
void lum_log(const char* level, const char* fmt, ...);
#define LUM_RETURN_ERROR(code, ...) \
  do {                              \
    lum_set_error(__VA_ARGS__);     \
    return (code);                  \
  } while (0)
#define LUM_TRY(expr)                            \
  do {                                           \
    LuminaryResult _r = (expr);                  \
    if (_r != LUMINARY_SUCCESS)                  \
      return _r | LUMINARY_ERROR_PROPAGATED;     \
  } while (0)
#define LUM_CHECK_NULL(arg)                                                     \
  do {                                                                          \
    if (!(arg))                                                                 \
      LUM_RETURN_ERROR(LUMINARY_ERROR_ARGUMENT_NULL, "argument %s is NULL", #arg); \
  } while (0)

/* host-side triangle soup of one mesh: reference mesh.h:8-20 (TriangleGeomData) */
typedef struct LumHostMesh {
  uint32_t triangle_count;
  float* vertex_buffer;         /* 9 floats per triangle */
  float* normal_buffer;         /* 9 */
  float* uv_buffer;             /* 6 */
  uint16_t* material_id_buffer; /* 1 */
} LumHostMesh;

void lum_host_mesh_free(LumHostMesh* mesh);

/* host-side texture: reference `Texture` (texture.h:21-40) as produced by png_load (host/png.c:415-712): always four
 * components, u8 or u16, wrap / linear (texture_create, texture.c:77-88). data == NULL marks an invalid texture
 * (texture_invalidate): it keeps its id and loads return their default value. */
typedef struct LumHostTexture {
  uint32_t width, height, pitch;
  uint32_t type; /* LUMB200_TEXTURE_U8 / LUMB200_TEXTURE_U16 */
  uint32_t num_components;
  float gamma;
  void* data;
} LumHostTexture;

void lum_host_texture_free(LumHostTexture* tex);
/* PNG reader: 8 / 16 bit, colour types 0 / 2 / 4 / 6, not interlaced (reference png_load). */
LuminaryResult lum_png_read(const char* path, LumHostTexture* tex);

/* reference wavefront.h: WavefrontArguments */
typedef struct LumWavefrontArgs {
  bool legacy_smoothness;
  bool force_transparency_cutout;
  float emission_scale;
  bool force_bidirectional_emission;
} LumWavefrontArgs;

void lum_wavefront_args_default(LumWavefrontArgs* args);

/* Reads one *.obj (+ its *.mtl libraries). On success `has_mesh` tells whether a mesh was produced (the reference
 * produces none for files without an `o` statement), `materials` holds the default material of the file followed by
 * one entry per newmtl, with ids material_offset + k. `textures` receives the files named by map_Kd / map_Ke / map_Ns /
 * map_refl / map_Bump statements (one entry per distinct path, ids texture_offset + k; files that fail to load stay in
 * the list as invalid textures). */
LuminaryResult lum_wavefront_load(
  const char* obj_path, LumWavefrontArgs args, uint32_t material_offset, uint32_t texture_offset, LumHostMesh* mesh, bool* has_mesh,
  LuminaryMaterial** materials, uint32_t* num_materials, LumHostTexture** textures, uint32_t* num_textures);

/* defaults: reference settings.c:6-28, camera.c:7-66, sky.c, material.c:5-29, mesh instance defaults */
void lum_settings_default(LuminaryRendererSettings* settings);
void lum_camera_default(LuminaryCamera* camera);
void lum_sky_default(LuminarySky* sky);
void lum_material_default(LuminaryMaterial* material);
/* ocean.c:6-22, cloud.c:6-53, fog.c:6-16, particles.c:6-24 of the reference: all inactive */
void lum_inactive_entities_default(LuminaryOcean* ocean, LuminaryCloud* cloud, LuminaryFog* fog, LuminaryParticles* particles);

/* *.lum version 4 (reference host/lum.c:47-128, host/lum_v4.c) */
typedef struct LumFileContent {
  LuminaryRendererSettings settings;
  LuminaryCamera camera;
  LuminarySky sky;
  LumWavefrontArgs wavefront_args;
  char** mesh_files; /* as written in the file (relative to the *.lum) */
  uint32_t num_mesh_files;
} LumFileContent;

void lum_file_content_init(LumFileContent* content);
void lum_file_content_free(LumFileContent* content);
LuminaryResult lum_file_read(const char* path, LumFileContent* content);

/* PNG writer: 8-bit RGBA, stored (uncompressed) deflate blocks. argb8 = LuminaryARGB8 pixels, ld in pixels. */
LuminaryResult lum_png_write_argb8(const char* path, const uint8_t* argb8, uint32_t width, uint32_t height, size_t ld);

struct LuminaryPath {
  char* string;
};

#endif /* LUM_HOST_INTERNAL_H */
