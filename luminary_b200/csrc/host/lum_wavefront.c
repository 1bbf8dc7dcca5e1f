/* lum_wavefront.c - *.obj / *.mtl loader of the host layer, plain C.
 *
 * Same observable behaviour as the reference loader (src/luminary/host/wavefront.c):
 *   - one mesh per file, non-indexed output: 9 position floats, 9 normal floats, 6 uv floats, 1 material id per
 *     triangle (wavefront_convert_content :828-996);
 *   - statements v, vn, vt, f (triangles and quads, quads split (0,1,2) (0,2,3), read_face :425-564), o, mtllib, usemtl;
 *   - material 0 of every file is the default material (:25-48, :65-67), `usemtl` of an unknown name selects it (:724-737);
 *   - a file without any `o` statement yields no mesh, only a warning (:845-848);
 *   - negative indices are resolved against the FINAL element counts (QUIRK, :877-879);
 *   - triangles whose two edges are both shorter than FLT_EPSILON per component are dropped (:901-905);
 *   - missing / degenerate normals fall back to the face normal (:921-975);
 *   - *.mtl: newmtl, Kd, d, Ks, Ns, Ke (scaled by emission_scale), Ni (:285-421); conversion to LuminaryMaterial
 *     with roughness = 1 - Ns / 1000, metallic = Ks.r > 0.5, emission_active = Ke > 0 (:758-824).
 * Texture maps: map_Kd (albedo), map_Ke (luminance), map_Ns (roughness), map_refl (metallic), map_Bump (normal) with the
 * option skipping of _wavefront_parse_map (:159-283); one texture per distinct path, loaded from PNG (lum_png.c). */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lum_host_internal.h"

typedef struct {
  int32_t v[3], vt[3], vn[3];
  uint16_t material;
} ObjTri;

enum { WF_ALBEDO = 0, WF_LUMINANCE = 1, WF_ROUGHNESS = 2, WF_METALLIC = 3, WF_NORMAL = 4, WF_TEXTURE_TYPES = 5 };
#define WF_TEXTURE_NONE 0xFFFFu

typedef struct {
  size_t hash;
  float kd[3], ks[3], ke[3];
  float dissolve, ns, ni;
  uint16_t texture[WF_TEXTURE_TYPES]; /* index into ObjContent.textures or WF_TEXTURE_NONE */
} ObjMaterial;

typedef struct {
  float* v;
  size_t nv, cv; /* vec3 */
  float* vn;
  size_t nvn, cvn;
  float* vt;
  size_t nvt, cvt; /* vec2 */
  ObjTri* tris;
  size_t ntris, ctris;
  ObjMaterial* mats;
  size_t nmats, cmats;
  size_t* loaded_mtls;
  size_t nloaded;
  LumHostTexture* textures;
  size_t* texture_hashes;
  size_t ntextures, ctextures, chashes;
  uint32_t num_objects;
  LumWavefrontArgs args;
} ObjContent;

#define GROW(ptr, count, cap, elems)                                         \
  do {                                                                       \
    if ((count) + 1 > (cap)) {                                               \
      (cap)  = (cap) ? (cap) * 2 : 1024;                                     \
      (ptr)  = realloc((ptr), sizeof(*(ptr)) * (elems) * (cap));             \
      if (!(ptr))                                                            \
        LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "out of host memory"); \
    }                                                                        \
  } while (0)

static size_t hash_djb2(const char* s) {
  size_t h = 5381;
  for (; *s; s++)
    h = ((h << 5) + h) + (unsigned char) *s;
  return h;
}

static ObjMaterial obj_default_material(void) {
  ObjMaterial m;
  memset(&m, 0, sizeof(m));
  m.kd[0] = m.kd[1] = m.kd[2] = 0.9f;
  m.dissolve                  = 1.0f;
  m.ns                        = 300.0f;
  m.ni                        = 1.0f;
  for (int k = 0; k < WF_TEXTURE_TYPES; k++)
    m.texture[k] = WF_TEXTURE_NONE;
  return m;
}

static uint32_t read_floats(const char* s, uint32_t n, float* dst) {
  uint32_t got = 0;
  char* end;
  while (got < n) {
    const float f = strtof(s, &end);
    if (end == s)
      break;
    dst[got++] = f;
    s          = end;
  }
  return got;
}

static void trim_line(char* s) {
  size_t n = strlen(s);
  while (n && (s[n - 1] == '\n' || s[n - 1] == '\r' || s[n - 1] == ' ' || s[n - 1] == '\t'))
    s[--n] = '\0';
}

/* _wavefront_parse_map, reference wavefront.c:159-283. `mtl_path` locates the texture file (path_apply). */
static LuminaryResult parse_map(ObjContent* c, const char* mtl_path, const char* line, size_t cur) {
  const size_t line_len = strlen(line);
  if (line_len < 8)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Line is too short to be a valid map_ line.");
  uint32_t path_offset = 7;
  int type;
  if (!strncmp(line + 4, "Kd", 2))
    type = WF_ALBEDO;
  else if (!strncmp(line + 4, "Ke", 2))
    type = WF_LUMINANCE;
  else if (!strncmp(line + 4, "Ns", 2))
    type = WF_ROUGHNESS;
  else if (!strncmp(line + 4, "refl", 4))
    type = WF_METALLIC, path_offset = 9;
  else if (!strncmp(line + 4, "Bump", 4))
    type = WF_NORMAL, path_offset = 9;
  else
    return LUMINARY_SUCCESS; /* not a supported type */
  if (line_len <= path_offset)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Line is too short to be a valid map_ line.");

  const char* path = line + path_offset;
  if (path[0] == '-') { /* skip "-option args..." groups until a word that is not an option */
    const char* command = path + 1;
    bool is_command     = true;
    while (is_command) {
      uint32_t num_args = 0;
      if (command[0] == 'o' || command[0] == 's')
        num_args = 3;
      else if (command[0] == 't')
        num_args = (command[1] == ' ') ? 3 : 1;
      else if (command[0] == 'm')
        num_args = 2;
      else if (command[0] == 'c' || command[0] == 'b')
        num_args = 1;
      for (uint32_t arg = 0; arg <= num_args; arg++) {
        command = strchr(command, ' ');
        if (!command)
          LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Something went wrong parsing the following line in an *.mtl file: %s", line);
        is_command = (command[1] == '-');
        command++;
      }
      if (is_command)
        command++;
    }
    path = command;
  }

  const size_t hash = hash_djb2(path);
  uint32_t id       = WF_TEXTURE_NONE;
  for (size_t k = 0; k < c->ntextures; k++)
    if (c->texture_hashes[k] == hash)
      id = (uint32_t) k;
  if (id == WF_TEXTURE_NONE) {
    if (c->ntextures >= 0xFFFFu)
      LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Exceeded limit of 65535 textures.");
    char file[4096];
    const char* slash = strrchr(mtl_path, '/');
    if (slash && path[0] != '/')
      snprintf(file, sizeof(file), "%.*s/%s", (int) (slash - mtl_path), mtl_path, path);
    else
      snprintf(file, sizeof(file), "%s", path);
    for (char* p = file; *p; p++) /* mtl files written on Windows */
      if (*p == '\\')
        *p = '/';
    size_t hash_cap = c->chashes;
    GROW(c->textures, c->ntextures, c->ctextures, 1);
    GROW(c->texture_hashes, c->ntextures, hash_cap, 1);
    c->chashes = hash_cap;
    LumHostTexture tex;
    if (lum_png_read(file, &tex) != LUMINARY_SUCCESS) {
      lum_log("error", "Failed to load texture %s: it stays in the scene as an invalid texture.", file);
      memset(&tex, 0, sizeof(tex)); /* texture_invalidate: keeps its id, loads return their default */
      tex.gamma = 1.0f;
    }
    id                         = (uint32_t) c->ntextures;
    c->textures[id]            = tex;
    c->texture_hashes[id]      = hash;
    c->ntextures++;
  }
  c->mats[cur].texture[type] = (uint16_t) id;
  return LUMINARY_SUCCESS;
}

static LuminaryResult read_mtl(ObjContent* c, const char* obj_path, const char* mtl_name) {
  char path[4096];
  const char* slash = strrchr(obj_path, '/');
  if (slash)
    snprintf(path, sizeof(path), "%.*s/%s", (int) (slash - obj_path), obj_path, mtl_name);
  else
    snprintf(path, sizeof(path), "%s", mtl_name);
  lum_log("log", "Reading *.mtl file (%s)", path);
  FILE* f = fopen(path, "r");
  if (!f)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Failed to open file *.mtl file (%s)", path);
  char line[4096];
  size_t cur = c->nmats - 1;
  while (fgets(line, sizeof(line), f)) {
    trim_line(line);
    if (!strncmp(line, "newmtl", 6)) {
      GROW(c->mats, c->nmats, c->cmats, 1);
      ObjMaterial m   = obj_default_material();
      m.hash          = hash_djb2(strlen(line) > 7 ? line + 7 : ""); /* an empty name must not read past the terminator */
      c->mats[c->nmats] = m;
      cur             = c->nmats++;
    }
    else if (line[0] == 'K' && line[1] == 'd') {
      if (read_floats(line + 3, 3, c->mats[cur].kd) != 3)
        lum_log("warn", "Expected three values in diffuse reflectivity in *.mtl file. Line: %s.", line);
    }
    else if (line[0] == 'd' && line[1] == ' ') {
      if (!read_floats(line + 2, 1, &c->mats[cur].dissolve))
        lum_log("warn", "Expected dissolve in *.mtl file but didn't find a number. Line: %s.", line);
    }
    else if (line[0] == 'K' && line[1] == 's') {
      if (read_floats(line + 3, 3, c->mats[cur].ks) != 3)
        lum_log("warn", "Expected three values in specular reflectivity in *.mtl file. Line: %s.", line);
    }
    else if (line[0] == 'N' && line[1] == 's') {
      if (!read_floats(line + 3, 1, &c->mats[cur].ns))
        lum_log("warn", "Expected specular_exponent in *.mtl file but didn't find a number. Line: %s.", line);
    }
    else if (line[0] == 'K' && line[1] == 'e') {
      float e[3];
      if (read_floats(line + 3, 3, e) == 3) {
        for (int k = 0; k < 3; k++)
          c->mats[cur].ke[k] = e[k] * c->args.emission_scale;
      }
      else
        lum_log("warn", "Expected three values in emission in *.mtl file. Line: %s.", line);
    }
    else if (line[0] == 'N' && line[1] == 'i') {
      if (!read_floats(line + 3, 1, &c->mats[cur].ni))
        lum_log("warn", "Expected refraction index in *.mtl file but didn't find a number. Line: %s.", line);
    }
    else if (!strncmp(line, "map_", 4)) {
      const LuminaryResult r = parse_map(c, path, line, cur);
      if (r != LUMINARY_SUCCESS) {
        fclose(f);
        return r;
      }
    }
  }
  fclose(f);
  return LUMINARY_SUCCESS;
}

/* "f a/b/c ..." -> up to 4 corners of (v, vt, vn); absent entries are 0. Returns the corner count (0 on error). */
static int read_face(const char* s, int32_t out[4][3]) {
  int corners = 0;
  s++; /* skip 'f' */
  while (*s) {
    while (*s == ' ' || *s == '\t')
      s++;
    if (!*s)
      break;
    if (corners == 4)
      return 5; /* polygon with more than four corners */
    int32_t vals[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) {
      char* end;
      const long v = strtol(s, &end, 10);
      if (end != s)
        vals[k] = (int32_t) v;
      s = end;
      if (*s == '/')
        s++;
      else
        break;
    }
    while (*s && *s != ' ' && *s != '\t')
      s++;
    out[corners][0] = vals[0], out[corners][1] = vals[1], out[corners][2] = vals[2];
    corners++;
  }
  return corners;
}

static LuminaryResult push_tri(ObjContent* c, int32_t f[4][3], int a, int b, int d, uint16_t material) {
  GROW(c->tris, c->ntris, c->ctris, 1);
  ObjTri t;
  t.v[0] = f[a][0], t.v[1] = f[b][0], t.v[2] = f[d][0];
  t.vt[0] = f[a][1], t.vt[1] = f[b][1], t.vt[2] = f[d][1];
  t.vn[0] = f[a][2], t.vn[1] = f[b][2], t.vn[2] = f[d][2];
  t.material         = material;
  c->tris[c->ntris++] = t;
  return LUMINARY_SUCCESS;
}

static LuminaryResult read_obj(ObjContent* c, const char* obj_path) {
  lum_log("log", "Reading *.obj file (%s)", obj_path);
  FILE* f = fopen(obj_path, "rb");
  if (!f)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "File %s could not be opened!", obj_path);
  /* an unseekable path (FIFO, device) makes ftell return -1: refuse it instead of allocating (size_t) -1 + 2 bytes */
  if (fseek(f, 0, SEEK_END) != 0) {
    fclose(f);
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "File %s is not seekable.", obj_path);
  }
  const long size = ftell(f);
  if (size < 0 || fseek(f, 0, SEEK_SET) != 0) {
    fclose(f);
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "File %s is not seekable.", obj_path);
  }
  char* data = (char*) malloc((size_t) size + 2);
  if (!data) {
    fclose(f);
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "out of host memory reading %s", obj_path);
  }
  const size_t got = fread(data, 1, (size_t) size, f);
  fclose(f);
  data[got]     = '\n';
  data[got + 1] = '\0';

  uint16_t current_material = 0;
  LuminaryResult result     = LUMINARY_SUCCESS;
  char* line                = data;
  char* eol;
  /* QUIRK: like the reference (wavefront.c:628) only newline-terminated lines are statements */
  while (result == LUMINARY_SUCCESS && (eol = strchr(line, '\n'))) {
    *eol = '\0';
    if (line[0] == 'v' && line[1] == ' ') {
      GROW(c->v, c->nv, c->cv, 3);
      float* d = c->v + 3 * c->nv++;
      d[0] = d[1] = d[2] = 0.0f;
      read_floats(line + 2, 3, d);
    }
    else if (line[0] == 'v' && line[1] == 'n') {
      GROW(c->vn, c->nvn, c->cvn, 3);
      float* d = c->vn + 3 * c->nvn++;
      d[0] = d[1] = d[2] = 0.0f;
      read_floats(line + 3, 3, d);
    }
    else if (line[0] == 'v' && line[1] == 't') {
      GROW(c->vt, c->nvt, c->cvt, 2);
      float* d = c->vt + 2 * c->nvt++;
      d[0] = d[1] = 0.0f;
      read_floats(line + 3, 2, d);
    }
    else if (line[0] == 'f' && (line[1] == ' ' || line[1] == '\t')) {
      int32_t fc[4][3];
      const int corners = read_face(line, fc);
      if (corners == 3) {
        result = push_tri(c, fc, 0, 1, 2, current_material);
      }
      else if (corners == 4) {
        result = push_tri(c, fc, 0, 1, 2, current_material);
        if (result == LUMINARY_SUCCESS)
          result = push_tri(c, fc, 0, 2, 3, current_material);
      }
      else {
        lum_log("error", "A face is of unsupported format. %s", line);
      }
    }
    else if (line[0] == 'o' && (line[1] == ' ' || line[1] == '\0')) {
      c->num_objects++;
    }
    else if (!strncmp(line, "mtllib", 6)) {
      char* name = line + 6;
      while (*name == ' ')
        name++;
      trim_line(name);
      const size_t h = hash_djb2(name);
      bool loaded    = false;
      for (size_t k = 0; k < c->nloaded; k++)
        loaded |= c->loaded_mtls[k] == h;
      if (!loaded) {
        size_t* grown = realloc(c->loaded_mtls, sizeof(size_t) * (c->nloaded + 1));
        if (!grown)
          continue; /* out of memory: the library is simply not loaded a second time */
        c->loaded_mtls = grown;
        c->loaded_mtls[c->nloaded++] = h;
        result                       = read_mtl(c, obj_path, name);
      }
    }
    else if (!strncmp(line, "usemtl", 6)) {
      char* name = line + 6;
      while (*name == ' ')
        name++;
      trim_line(name);
      const size_t h   = hash_djb2(name);
      current_material = 0;
      for (size_t k = 1; k < c->nmats; k++) {
        if (c->mats[k].hash == h) {
          current_material = (uint16_t) k;
          break;
        }
      }
    }
    line = eol + 1;
  }
  free(data);
  return result;
}

static void convert_material(const ObjContent* c, size_t k, uint32_t id, uint32_t texture_offset, LuminaryMaterial* out) {
  const ObjMaterial* w = &c->mats[k];
  lum_material_default(out);
  out->id                       = id;
  out->base_substrate           = LUMINARY_MATERIAL_BASE_SUBSTRATE_OPAQUE;
  out->albedo.r                 = w->kd[0];
  out->albedo.g                 = w->kd[1];
  out->albedo.b                 = w->kd[2];
  out->albedo.a                 = w->dissolve;
  out->emission.r               = w->ke[0];
  out->emission.g               = w->ke[1];
  out->emission.b               = w->ke[2];
  out->emission_scale           = c->args.emission_scale;
  out->refraction_index         = w->ni;
  out->roughness                = 1.0f - w->ns / 1000.0f;
  out->roughness_clamp          = 0.25f;
  out->roughness_as_smoothness  = c->args.legacy_smoothness;
  out->emission_active          = (w->texture[WF_LUMINANCE] != WF_TEXTURE_NONE) || (w->ke[0] > 0.0f) || (w->ke[1] > 0.0f) || (w->ke[2] > 0.0f);
  out->thin_walled              = false;
  out->normal_map_is_compressed = true;
  out->bidirectional_emission   = c->args.force_bidirectional_emission;
  out->metallic                 = w->ks[0] > 0.5f;
#define WF_TEX(type) ((w->texture[type] != WF_TEXTURE_NONE) ? (uint16_t) (texture_offset + w->texture[type]) : (uint16_t) WF_TEXTURE_NONE)
  out->albedo_tex    = WF_TEX(WF_ALBEDO);
  out->luminance_tex = WF_TEX(WF_LUMINANCE);
  out->roughness_tex = WF_TEX(WF_ROUGHNESS);
  out->metallic_tex  = WF_TEX(WF_METALLIC);
  out->normal_tex    = WF_TEX(WF_NORMAL);
#undef WF_TEX
}

static uint32_t resolve(int32_t idx, size_t count) { return (idx > 0) ? (uint32_t) (idx - 1) : (uint32_t) (idx + (int64_t) count); }

static bool bad(float f) { return isnan(f) || isinf(f); }

static LuminaryResult convert_mesh(const ObjContent* c, uint32_t material_offset, LumHostMesh* mesh) {
  const size_t n = c->ntris;
  mesh->vertex_buffer      = (float*) malloc(sizeof(float) * 9 * (n ? n : 1));
  mesh->normal_buffer      = (float*) malloc(sizeof(float) * 9 * (n ? n : 1));
  mesh->uv_buffer          = (float*) malloc(sizeof(float) * 6 * (n ? n : 1));
  mesh->material_id_buffer = (uint16_t*) malloc(sizeof(uint16_t) * (n ? n : 1));
  if (!mesh->vertex_buffer || !mesh->normal_buffer || !mesh->uv_buffer || !mesh->material_id_buffer)
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "out of host memory converting a mesh of %zu triangles", n);
  uint32_t out = 0;
  for (size_t i = 0; i < n; i++) {
    const ObjTri* t = &c->tris[i];
    uint32_t vi[3];
    bool ok = true;
    for (int k = 0; k < 3; k++) {
      vi[k] = resolve(t->v[k], c->nv);
      ok &= vi[k] < c->nv;
    }
    if (!ok)
      continue;
    const float* v1 = c->v + 3 * vi[0];
    const float* v2 = c->v + 3 * vi[1];
    const float* v3 = c->v + 3 * vi[2];
    const float e1[3] = {v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]};
    const float e2[3] = {v3[0] - v1[0], v3[1] - v1[1], v3[2] - v1[2]};
    if (fabsf(e1[0]) < FLT_EPSILON && fabsf(e1[1]) < FLT_EPSILON && fabsf(e1[2]) < FLT_EPSILON && fabsf(e2[0]) < FLT_EPSILON
        && fabsf(e2[1]) < FLT_EPSILON && fabsf(e2[2]) < FLT_EPSILON)
      continue;
    float* dv = mesh->vertex_buffer + 9 * (size_t) out;
    memcpy(dv + 0, v1, 12), memcpy(dv + 3, v2, 12), memcpy(dv + 6, v3, 12);

    float fn[3]       = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const float fn_rl = 1.0f / sqrtf(fn[0] * fn[0] + fn[1] * fn[1] + fn[2] * fn[2]);
    if (!bad(fn_rl))
      fn[0] *= fn_rl, fn[1] *= fn_rl, fn[2] *= fn_rl;

    float* duv = mesh->uv_buffer + 6 * (size_t) out;
    float* dn  = mesh->normal_buffer + 9 * (size_t) out;
    for (int k = 0; k < 3; k++) {
      const uint32_t ti = resolve(t->vt[k], c->nvt);
      duv[2 * k + 0]    = (ti < c->nvt) ? c->vt[2 * ti + 0] : 0.0f;
      duv[2 * k + 1]    = (ti < c->nvt) ? c->vt[2 * ti + 1] : 0.0f;
      const uint32_t ni = resolve(t->vn[k], c->nvn);
      float nn[3]       = {fn[0], fn[1], fn[2]};
      if (ni < c->nvn)
        memcpy(nn, c->vn + 3 * ni, 12);
      const float rl = 1.0f / sqrtf(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
      if (bad(rl)) {
        dn[3 * k + 0] = fn[0], dn[3 * k + 1] = fn[1], dn[3 * k + 2] = fn[2];
      }
      else {
        dn[3 * k + 0] = nn[0] * rl, dn[3 * k + 1] = nn[1] * rl, dn[3 * k + 2] = nn[2] * rl;
      }
    }
    mesh->material_id_buffer[out] = (uint16_t) (material_offset + t->material);
    out++;
  }
  mesh->triangle_count = out;
  return LUMINARY_SUCCESS;
}

void lum_wavefront_args_default(LumWavefrontArgs* args) { /* wavefront.c:998-1007 */
  args->legacy_smoothness            = false;
  args->force_transparency_cutout    = false;
  args->emission_scale               = 1.0f;
  args->force_bidirectional_emission = false;
}

void lum_host_mesh_free(LumHostMesh* mesh) {
  free(mesh->vertex_buffer);
  free(mesh->normal_buffer);
  free(mesh->uv_buffer);
  free(mesh->material_id_buffer);
  memset(mesh, 0, sizeof(*mesh));
}

LuminaryResult lum_wavefront_load(
  const char* obj_path, LumWavefrontArgs args, uint32_t material_offset, uint32_t texture_offset, LumHostMesh* mesh, bool* has_mesh,
  LuminaryMaterial** materials, uint32_t* num_materials, LumHostTexture** textures, uint32_t* num_textures) {
  LUM_CHECK_NULL(textures);
  LUM_CHECK_NULL(num_textures);
  *textures     = NULL;
  *num_textures = 0;
  LUM_CHECK_NULL(obj_path);
  LUM_CHECK_NULL(mesh);
  LUM_CHECK_NULL(has_mesh);
  LUM_CHECK_NULL(materials);
  LUM_CHECK_NULL(num_materials);
  memset(mesh, 0, sizeof(*mesh));
  *has_mesh      = false;
  *materials     = NULL;
  *num_materials = 0;

  ObjContent c;
  memset(&c, 0, sizeof(c));
  c.args  = args;
  c.mats  = (ObjMaterial*) malloc(sizeof(ObjMaterial) * 16);
  c.cmats = 16;
  if (!c.mats)
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "out of host memory");
  c.mats[0] = obj_default_material();
  c.nmats   = 1;

  LuminaryResult result = read_obj(&c, obj_path);
  if (result == LUMINARY_SUCCESS) {
    if (c.num_objects == 0) {
      lum_log("warn", "Wavefront file contained no objects.");
    }
    else {
      *materials = (LuminaryMaterial*) malloc(sizeof(LuminaryMaterial) * c.nmats);
      if (!*materials)
        result = LUMINARY_ERROR_OUT_OF_MEMORY;
      else {
        for (size_t k = 0; k < c.nmats; k++)
          convert_material(&c, k, material_offset + (uint32_t) k, texture_offset, *materials + k);
        *num_materials = (uint32_t) c.nmats;
        result         = convert_mesh(&c, material_offset, mesh);
        *has_mesh      = result == LUMINARY_SUCCESS;
      }
    }
  }
  free(c.v), free(c.vn), free(c.vt), free(c.tris), free(c.mats), free(c.loaded_mtls), free(c.texture_hashes);
  if (result == LUMINARY_SUCCESS && *has_mesh && texture_offset + c.ntextures > 0xFFFFu) {
    lum_set_error("Exceeded limit of 65535 textures.");
    result = LUMINARY_ERROR_API_EXCEPTION;
  }
  if (result == LUMINARY_SUCCESS && *has_mesh) {
    *textures     = c.textures; /* ownership moves to the caller */
    *num_textures = (uint32_t) c.ntextures;
  }
  else {
    for (size_t k = 0; k < c.ntextures; k++)
      lum_host_texture_free(&c.textures[k]);
    free(c.textures);
  }
  if (result != LUMINARY_SUCCESS) {
    *has_mesh = false;
    lum_host_mesh_free(mesh);
    free(*materials);
    *materials     = NULL;
    *num_materials = 0;
  }
  return result;
}
