/*
 * orc_trace.c - oracle: world-space flattening of instances, a CPU BVH2 and closest-hit queries.
 * TEST INFRASTRUCTURE ONLY (see lum_oracle.h).
 *
 * What it restates: the semantics of the reference's closest-hit launch
 * (device/optix/optix_kernel_raytrace.cu:82-95,147-183 with the any-hit program
 * device/cuda/optix_anyhit.cuh:15-31): closest intersection over all instances' triangles, no face
 * culling, tmin = 0, the "ignore handle" triangle rejected, miss => HIT_TYPE_SKY. The ray/triangle and
 * ray/box arithmetic of the reference is inside NVIDIA OptiX (closed source), so it cannot be restated;
 * two tests are provided instead:
 *   - orc_tri_mt:         the reference's own Moeller-Trumbore (cuda/math.cuh:1337-1358), used by the
 *                         reference wherever it re-intersects a triangle itself (light_triangle.cuh:10-31);
 *   - orc_tri_watertight: Woop/Benthin/Wald 2013 in a fixed operation order. The product's traversal
 *                         kernel evaluates exactly this sequence of IEEE operations, so ids are bit-exact.
 * Equal-t ties are resolved towards the smaller flattened primitive index in both implementations.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "lum_oracle.h"
#include "orc_internal.h"

/* ------------------------------------------------------------------ */
/* triangle tests                                                       */
/* ------------------------------------------------------------------ */
float orc_tri_mt(const float* v9, OrcVec3 origin, OrcVec3 ray, float* u_out, float* v_out) {
  const OrcVec3 vertex = v_get(v9[0], v9[1], v9[2]);
  const OrcVec3 edge1  = v_sub(v_get(v9[3], v9[4], v9[5]), vertex);
  const OrcVec3 edge2  = v_sub(v_get(v9[6], v9[7], v9[8]), vertex);

  const OrcVec3 h = v_cross(ray, edge2);
  const float a   = v_dot(edge1, h);
  const float f   = 1.0f / a;
  const OrcVec3 s = v_sub(origin, vertex);
  const float u   = f * v_dot(s, h);
  const OrcVec3 q = v_cross(s, edge1);
  const float v   = f * v_dot(ray, q);

  if (u_out)
    *u_out = u;
  if (v_out)
    *v_out = v;

  if (v < 0.0f || u < 0.0f || !(u + v <= 1.0f))
    return ORC_FLT_MAX;

  const float t = f * v_dot(edge2, q);
  /* __fslctf(t, FLT_MAX, t): t >= 0 ? t : FLT_MAX (NaN -> FLT_MAX) */
  return (t >= 0.0f) ? t : ORC_FLT_MAX;
}

typedef struct {
  int kx, ky, kz;
  float Sx, Sy, Sz;
  float o[3];
} RayPre;

static void ray_pre(RayPre* p, OrcVec3 origin, OrcVec3 ray) {
  const float d[3] = {ray.x, ray.y, ray.z};
  const float ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]);
  int kz;
  if (ax >= ay && ax >= az)
    kz = 0;
  else if (ay >= az)
    kz = 1;
  else
    kz = 2;
  int kx = (kz + 1) % 3;
  int ky = (kx + 1) % 3;
  if (d[kz] < 0.0f) {
    const int tmp = kx;
    kx            = ky;
    ky            = tmp;
  }
  p->kx   = kx;
  p->ky   = ky;
  p->kz   = kz;
  p->Sx   = d[kx] / d[kz];
  p->Sy   = d[ky] / d[kz];
  p->Sz   = 1.0f / d[kz];
  p->o[0] = origin.x;
  p->o[1] = origin.y;
  p->o[2] = origin.z;
}

static inline bool tri_wt(const RayPre* p, const float* v9, float* t_out, float* u_out, float* v_out) {
  const int kx = p->kx, ky = p->ky, kz = p->kz;
  const float A[3] = {v9[0] - p->o[0], v9[1] - p->o[1], v9[2] - p->o[2]};
  const float B[3] = {v9[3] - p->o[0], v9[4] - p->o[1], v9[5] - p->o[2]};
  const float C[3] = {v9[6] - p->o[0], v9[7] - p->o[1], v9[8] - p->o[2]};

  const float Ax = A[kx] - p->Sx * A[kz];
  const float Ay = A[ky] - p->Sy * A[kz];
  const float Bx = B[kx] - p->Sx * B[kz];
  const float By = B[ky] - p->Sy * B[kz];
  const float Cx = C[kx] - p->Sx * C[kz];
  const float Cy = C[ky] - p->Sy * C[kz];

  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;

  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    const double CxBy = (double) Cx * (double) By;
    const double CyBx = (double) Cy * (double) Bx;
    U                 = (float) (CxBy - CyBx);
    const double AxCy = (double) Ax * (double) Cy;
    const double AyCx = (double) Ay * (double) Cx;
    V                 = (float) (AxCy - AyCx);
    const double BxAy = (double) Bx * (double) Ay;
    const double ByAx = (double) By * (double) Ax;
    W                 = (float) (BxAy - ByAx);
  }

  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f))
    return false;

  const float det = U + V + W;
  if (det == 0.0f)
    return false;

  const float Az = p->Sz * A[kz];
  const float Bz = p->Sz * B[kz];
  const float Cz = p->Sz * C[kz];
  const float T  = U * Az + V * Bz + W * Cz;

  const float rcp = 1.0f / det;
  *t_out          = T * rcp;
  *u_out          = V * rcp;
  *v_out          = W * rcp;
  return true;
}

bool orc_tri_watertight(const float* v9, OrcVec3 origin, OrcVec3 ray, float* t, float* u, float* v) {
  RayPre p;
  ray_pre(&p, origin, ray);
  return tri_wt(&p, v9, t, u, v);
}

/* ------------------------------------------------------------------ */
/* scene + BVH2 (binned SAH)                                            */
/* ------------------------------------------------------------------ */
typedef struct {
  float lo[3], hi[3];
} Box;

static inline void box_init(Box* b) {
  for (int k = 0; k < 3; k++) {
    b->lo[k] = INFINITY;
    b->hi[k] = -INFINITY;
  }
}
static inline void box_grow_pt(Box* b, const float* p) {
  for (int k = 0; k < 3; k++) {
    b->lo[k] = fminf(b->lo[k], p[k]);
    b->hi[k] = fmaxf(b->hi[k], p[k]);
  }
}
static inline void box_grow(Box* b, const Box* o) {
  for (int k = 0; k < 3; k++) {
    b->lo[k] = fminf(b->lo[k], o->lo[k]);
    b->hi[k] = fmaxf(b->hi[k], o->hi[k]);
  }
}
static inline float box_area(const Box* b) {
  const float dx = b->hi[0] - b->lo[0], dy = b->hi[1] - b->lo[1], dz = b->hi[2] - b->lo[2];
  if (!(dx >= 0.0f))
    return 0.0f;
  return 2.0f * (dx * dy + dy * dz + dz * dx);
}

typedef struct {
  const float* tris; /* 9 floats per prim */
  Box* prim_box;
  float* centroid; /* 3 per prim */
  uint32_t* order;
  OrcBvhNode* nodes;
  uint32_t num_nodes;
} Builder;

#define BVH_BINS 16
#define BVH_LEAF 4

static void build_rec(Builder* B, uint32_t node_id, uint32_t first, uint32_t count) {
  OrcBvhNode* node = &B->nodes[node_id];
  Box bounds, cbounds;
  box_init(&bounds);
  box_init(&cbounds);
  for (uint32_t i = first; i < first + count; i++) {
    const uint32_t p = B->order[i];
    box_grow(&bounds, &B->prim_box[p]);
    box_grow_pt(&cbounds, &B->centroid[3 * p]);
  }
  memcpy(node->lo, bounds.lo, sizeof(float) * 3);
  memcpy(node->hi, bounds.hi, sizeof(float) * 3);

  if (count <= BVH_LEAF) {
    node->left  = first;
    node->count = count;
    return;
  }

  int best_axis   = -1;
  int best_split  = 0;
  float best_cost = INFINITY;

  for (int axis = 0; axis < 3; axis++) {
    const float cmin = cbounds.lo[axis], cmax = cbounds.hi[axis];
    if (!(cmax > cmin))
      continue;
    const float scale = BVH_BINS / (cmax - cmin);
    Box bins[BVH_BINS];
    uint32_t bin_count[BVH_BINS] = {0};
    for (int b = 0; b < BVH_BINS; b++)
      box_init(&bins[b]);
    for (uint32_t i = first; i < first + count; i++) {
      const uint32_t p = B->order[i];
      int b            = (int) ((B->centroid[3 * p + axis] - cmin) * scale);
      if (b >= BVH_BINS)
        b = BVH_BINS - 1;
      if (b < 0)
        b = 0;
      bin_count[b]++;
      box_grow(&bins[b], &B->prim_box[p]);
    }
    float right_area[BVH_BINS];
    uint32_t right_cnt[BVH_BINS];
    Box acc;
    box_init(&acc);
    uint32_t c = 0;
    for (int b = BVH_BINS - 1; b > 0; b--) {
      box_grow(&acc, &bins[b]);
      c += bin_count[b];
      right_area[b] = box_area(&acc);
      right_cnt[b]  = c;
    }
    box_init(&acc);
    c = 0;
    for (int b = 0; b < BVH_BINS - 1; b++) {
      box_grow(&acc, &bins[b]);
      c += bin_count[b];
      if (c == 0 || right_cnt[b + 1] == 0)
        continue;
      const float cost = box_area(&acc) * c + right_area[b + 1] * right_cnt[b + 1];
      if (cost < best_cost) {
        best_cost  = cost;
        best_axis  = axis;
        best_split = b;
      }
    }
  }

  uint32_t mid;
  if (best_axis < 0) {
    mid = first + count / 2; /* all centroids coincide */
  }
  else {
    const float cmin  = cbounds.lo[best_axis];
    const float scale = BVH_BINS / (cbounds.hi[best_axis] - cmin);
    uint32_t i = first, j = first + count;
    while (i < j) {
      const uint32_t p = B->order[i];
      int b            = (int) ((B->centroid[3 * p + best_axis] - cmin) * scale);
      if (b >= BVH_BINS)
        b = BVH_BINS - 1;
      if (b < 0)
        b = 0;
      if (b <= best_split)
        i++;
      else {
        j--;
        B->order[i] = B->order[j];
        B->order[j] = p;
      }
    }
    mid = i;
    if (mid == first || mid == first + count)
      mid = first + count / 2;
  }

  const uint32_t left = B->num_nodes;
  B->num_nodes += 2;
  node        = &B->nodes[node_id];
  node->left  = left;
  node->count = 0;
  build_rec(B, left, first, mid - first);
  build_rec(B, left + 1, mid, first + count - mid);
}

static void build_bvh(const float* tris, uint32_t n, OrcBvhNode** nodes_out, uint32_t* num_nodes_out, uint32_t** order_out) {
  Builder B;
  B.tris      = tris;
  B.prim_box  = (Box*) malloc(sizeof(Box) * (n ? n : 1));
  B.centroid  = (float*) malloc(sizeof(float) * 3 * (n ? n : 1));
  B.order     = (uint32_t*) malloc(sizeof(uint32_t) * (n ? n : 1));
  B.nodes     = (OrcBvhNode*) malloc(sizeof(OrcBvhNode) * (2 * (size_t) n + 2));
  B.num_nodes = 1;
  for (uint32_t p = 0; p < n; p++) {
    Box b;
    box_init(&b);
    box_grow_pt(&b, tris + 9 * (size_t) p);
    box_grow_pt(&b, tris + 9 * (size_t) p + 3);
    box_grow_pt(&b, tris + 9 * (size_t) p + 6);
    B.prim_box[p] = b;
    for (int k = 0; k < 3; k++)
      B.centroid[3 * p + k] = 0.5f * (b.lo[k] + b.hi[k]);
    B.order[p] = p;
  }
  if (n == 0) {
    memset(&B.nodes[0], 0, sizeof(OrcBvhNode));
    B.nodes[0].lo[0] = B.nodes[0].lo[1] = B.nodes[0].lo[2] = INFINITY;
    B.nodes[0].hi[0] = B.nodes[0].hi[1] = B.nodes[0].hi[2] = -INFINITY;
  }
  else {
    build_rec(&B, 0, 0, n);
  }
  free(B.prim_box);
  free(B.centroid);
  *nodes_out     = B.nodes;
  *num_nodes_out = B.num_nodes;
  *order_out     = B.order;
}

OrcScene* orc_scene_create(
  const OrcMesh* meshes, uint32_t num_meshes, const OrcInstance* instances, uint32_t num_instances, const OrcMaterialPacked* materials,
  uint32_t num_materials) {
  OrcScene* s = (OrcScene*) calloc(1, sizeof(OrcScene));
  s->num_meshes    = num_meshes;
  s->num_instances = num_instances;
  s->num_materials = num_materials;
  s->meshes        = (OrcMesh*) malloc(sizeof(OrcMesh) * (num_meshes ? num_meshes : 1));
  for (uint32_t m = 0; m < num_meshes; m++) {
    const uint32_t n = meshes[m].num_tris;
    OrcMesh* d       = &s->meshes[m];
    d->num_tris      = n;
    float* vb        = (float*) malloc(sizeof(float) * 9 * (n ? n : 1));
    float* nb        = (float*) malloc(sizeof(float) * 9 * (n ? n : 1));
    float* ub        = (float*) malloc(sizeof(float) * 6 * (n ? n : 1));
    uint16_t* mb     = (uint16_t*) malloc(sizeof(uint16_t) * (n ? n : 1));
    memcpy(vb, meshes[m].vertex, sizeof(float) * 9 * n);
    memcpy(nb, meshes[m].normal, sizeof(float) * 9 * n);
    memcpy(ub, meshes[m].uv, sizeof(float) * 6 * n);
    memcpy(mb, meshes[m].material, sizeof(uint16_t) * n);
    d->vertex   = vb;
    d->normal   = nb;
    d->uv       = ub;
    d->material = mb;
  }
  s->instances = (OrcInstance*) malloc(sizeof(OrcInstance) * (num_instances ? num_instances : 1));
  memcpy(s->instances, instances, sizeof(OrcInstance) * num_instances);
  s->materials = (OrcMaterialPacked*) malloc(sizeof(OrcMaterialPacked) * (num_materials ? num_materials : 1));
  memcpy(s->materials, materials, sizeof(OrcMaterialPacked) * num_materials);

  s->instance_prim_offset = (uint32_t*) malloc(sizeof(uint32_t) * (num_instances + 1));
  uint32_t total          = 0;
  for (uint32_t i = 0; i < num_instances; i++) {
    s->instance_prim_offset[i] = total;
    total += s->meshes[instances[i].mesh_id].num_tris;
  }
  s->instance_prim_offset[num_instances] = total;
  s->num_prims                           = total;

  s->world         = (float*) malloc(sizeof(float) * 9 * (size_t) (total ? total : 1));
  s->prim_instance = (uint32_t*) malloc(sizeof(uint32_t) * (total ? total : 1));
  s->prim_tri      = (uint32_t*) malloc(sizeof(uint32_t) * (total ? total : 1));

  /* world-space vertex = transform_apply(trans, v), cuda/math.cuh:459-491 (light_triangle.cuh:56-58 does the same for lights) */
  for (uint32_t i = 0; i < num_instances; i++) {
    const OrcMesh* m      = &s->meshes[instances[i].mesh_id];
    const OrcTransform* t = &s->instances[i].transform;
    const uint32_t base   = s->instance_prim_offset[i];
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t) m->num_tris; k++) {
      for (int v = 0; v < 3; v++) {
        const float* src        = m->vertex + 9 * k + 3 * v;
        const OrcVec3 w         = orc_transform_apply(t, v_get(src[0], src[1], src[2]));
        float* dst              = s->world + 9 * ((size_t) base + k) + 3 * v;
        dst[0]                  = w.x;
        dst[1]                  = w.y;
        dst[2]                  = w.z;
      }
      s->prim_instance[base + k] = i;
      s->prim_tri[base + k]      = (uint32_t) k;
    }
  }

  build_bvh(s->world, total, &s->nodes, &s->num_nodes, &s->prim_order);
  return s;
}

void orc_scene_destroy(OrcScene* s) {
  if (!s)
    return;
  for (uint32_t m = 0; m < s->num_meshes; m++) {
    free((void*) s->meshes[m].vertex);
    free((void*) s->meshes[m].normal);
    free((void*) s->meshes[m].uv);
    free((void*) s->meshes[m].material);
  }
  free(s->meshes);
  free(s->instances);
  free(s->materials);
  free(s->instance_prim_offset);
  free(s->world);
  free(s->prim_instance);
  free(s->prim_tri);
  free(s->nodes);
  free(s->prim_order);
  free(s->light_nodes);
  free(s->light_order);
  free(s->light_world);
  free(s->textures);
  orc_sky_free(s->sky);
  free(s);
}

uint32_t orc_scene_num_prims(const OrcScene* s) { return s->num_prims; }
const float* orc_scene_world_tris(const OrcScene* s) { return s->world; }
void orc_scene_prim_handle(const OrcScene* s, uint32_t prim, uint32_t* instance_id, uint32_t* tri_id) {
  *instance_id = s->prim_instance[prim];
  *tri_id      = s->prim_tri[prim];
}

/* ------------------------------------------------------------------ */
/* traversal                                                            */
/* ------------------------------------------------------------------ */
static inline bool slab(const OrcBvhNode* n, const float* o, const float* inv, float tmin, float tmax, float* tnear_out) {
  float tn = tmin, tf = tmax;
  for (int k = 0; k < 3; k++) {
    const float t0 = (n->lo[k] - o[k]) * inv[k];
    const float t1 = (n->hi[k] - o[k]) * inv[k];
    tn             = fmaxf(tn, fminf(t0, t1));
    tf             = fminf(tf, fmaxf(t0, t1));
  }
  *tnear_out = tn;
  return tn <= tf * 1.0000004f;
}

/* alpha_scene != NULL: hits on albedo-textured materials whose texel alpha is 0 are ignored, the geometry_trace
 * any-hit program of the reference (cuda/optix_anyhit.cuh:15-31 -> optix_alpha_test, optix_common.cuh:20-46) */
static OrcHit bvh_closest_impl(
  const OrcBvhNode* nodes, const uint32_t* order, const float* tris, OrcVec3 origin, OrcVec3 ray, float tmin, float tmax, uint32_t ignore_prim,
  uint64_t* nodes_visited, uint64_t* tris_tested, const OrcScene* alpha_scene) {
  OrcHit best;
  best.prim = ORC_HIT_SKY;
  best.t    = tmax;
  best.u = best.v = 0.0f;

  RayPre pre;
  ray_pre(&pre, origin, ray);
  const float o[3]   = {origin.x, origin.y, origin.z};
  const float inv[3] = {1.0f / ray.x, 1.0f / ray.y, 1.0f / ray.z};

  uint32_t stack[128];
  int sp      = 0;
  stack[sp++] = 0;
  uint64_t nv = 0, tt = 0;

  while (sp > 0) {
    const OrcBvhNode* n = &nodes[stack[--sp]];
    float tn;
    nv++;
    if (!slab(n, o, inv, tmin, best.t, &tn))
      continue;
    if (n->count) {
      for (uint32_t i = 0; i < n->count; i++) {
        const uint32_t p = order[n->left + i];
        if (p == ignore_prim)
          continue;
        float t, u, v;
        tt++;
        if (!tri_wt(&pre, tris + 9 * (size_t) p, &t, &u, &v))
          continue;
        if (!(t >= tmin))
          continue;
        if (alpha_scene && orc_alpha_cutout(alpha_scene, p, u, v))
          continue;
        if (t < best.t || (t == best.t && best.prim != ORC_HIT_SKY && p < best.prim)) {
          best.t    = t;
          best.u    = u;
          best.v    = v;
          best.prim = p;
        }
      }
    }
    else {
      float t0, t1;
      const bool h0 = slab(&nodes[n->left], o, inv, tmin, best.t, &t0);
      const bool h1 = slab(&nodes[n->left + 1], o, inv, tmin, best.t, &t1);
      if (h0 && h1) {
        if (t0 <= t1) {
          stack[sp++] = n->left + 1;
          stack[sp++] = n->left;
        }
        else {
          stack[sp++] = n->left;
          stack[sp++] = n->left + 1;
        }
      }
      else if (h0)
        stack[sp++] = n->left;
      else if (h1)
        stack[sp++] = n->left + 1;
    }
  }
  if (nodes_visited)
    *nodes_visited += nv;
  if (tris_tested)
    *tris_tested += tt;
  if (best.prim == ORC_HIT_SKY)
    best.t = ORC_FLT_MAX;
  return best;
}

OrcHit orc_bvh_closest(
  const OrcBvhNode* nodes, const uint32_t* order, const float* tris, OrcVec3 origin, OrcVec3 ray, float tmin, float tmax, uint32_t ignore_prim,
  uint64_t* nodes_visited, uint64_t* tris_tested) {
  return bvh_closest_impl(nodes, order, tris, origin, ray, tmin, tmax, ignore_prim, nodes_visited, tris_tested, NULL);
}

OrcHit orc_closest_hit(
  const OrcScene* s, OrcVec3 origin, OrcVec3 ray, float tmin, float tmax, uint32_t ignore_prim, uint64_t* nodes_visited, uint64_t* tris_tested) {
  return bvh_closest_impl(
    s->nodes, s->prim_order, s->world, origin, ray, tmin, tmax, ignore_prim, nodes_visited, tris_tested, s->num_textures ? s : NULL);
}

OrcHit orc_closest_hit_bruteforce(const OrcScene* s, OrcVec3 origin, OrcVec3 ray, float tmin, float tmax, uint32_t ignore_prim, int use_mt) {
  OrcHit best;
  best.prim = ORC_HIT_SKY;
  best.t    = tmax;
  best.u = best.v = 0.0f;
  RayPre pre;
  ray_pre(&pre, origin, ray);
  for (uint32_t p = 0; p < s->num_prims; p++) {
    if (p == ignore_prim)
      continue;
    float t, u, v;
    if (use_mt) {
      t = orc_tri_mt(s->world + 9 * (size_t) p, origin, ray, &u, &v);
      if (t == ORC_FLT_MAX)
        continue;
    }
    else if (!tri_wt(&pre, s->world + 9 * (size_t) p, &t, &u, &v))
      continue;
    if (!(t >= tmin))
      continue;
    if (s->num_textures && orc_alpha_cutout(s, p, u, v))
      continue;
    if (t < best.t) { /* ascending p => ties keep the smaller index */
      best.t    = t;
      best.u    = u;
      best.v    = v;
      best.prim = p;
    }
  }
  if (best.prim == ORC_HIT_SKY)
    best.t = ORC_FLT_MAX;
  return best;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

double orc_trace_rays(
  const OrcScene* s, const float* origins, const float* dirs, uint32_t n, uint32_t* out_prim, float* out_t, float* out_u, float* out_v,
  int num_threads) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#endif
  const double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < (int64_t) n; i++) {
    const OrcVec3 o = v_get(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
    const OrcVec3 d = v_get(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]);
    const OrcHit h  = orc_closest_hit(s, o, d, 0.0f, ORC_FLT_MAX, 0xFFFFFFFFu, NULL, NULL);
    out_prim[i]     = h.prim;
    out_t[i]        = h.t;
    if (out_u)
      out_u[i] = h.u;
    if (out_v)
      out_v[i] = h.v;
  }
  return now_s() - t0;
}

/* primary rays of one pass: tasks_create (cuda/kernels.cuh:45-193) + the closest-hit launch */
double orc_trace_primary(
  const OrcScene* s, const OrcCamera* cam, const OrcSettings* set, uint32_t sample_id, uint32_t* out_instance, uint32_t* out_tri, float* out_t,
  float* out_u, float* out_v, int num_threads, uint64_t* nodes_visited, uint64_t* tris_tested) {
#ifdef _OPENMP
  if (num_threads > 0)
    omp_set_num_threads(num_threads);
#endif
  const uint32_t n = set->width * set->height;
  uint64_t nv_total = 0, tt_total = 0;
  const double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nv_total, tt_total)
  for (int64_t i = 0; i < (int64_t) n; i++) {
    const uint32_t y = (uint32_t) (i / set->width);
    const uint32_t x = (uint32_t) (i - (int64_t) y * set->width);
    OrcVec3 o, d;
    orc_camera_sample(cam, set, orc_path_id_get(x, y, sample_id), &o, &d);
    uint64_t nv = 0, tt = 0;
    const OrcHit h = orc_closest_hit(s, o, d, 0.0f, ORC_FLT_MAX, 0xFFFFFFFFu, &nv, &tt);
    nv_total += nv;
    tt_total += tt;
    if (h.prim == ORC_HIT_SKY) {
      out_instance[i] = ORC_HIT_SKY;
      out_tri[i]      = 0;
    }
    else {
      out_instance[i] = s->prim_instance[h.prim];
      out_tri[i]      = s->prim_tri[h.prim];
    }
    out_t[i] = h.t;
    if (out_u)
      out_u[i] = h.u;
    if (out_v)
      out_v[i] = h.v;
  }
  if (nodes_visited)
    *nodes_visited = nv_total;
  if (tris_tested)
    *tris_tested = tt_total;
  return now_s() - t0;
}

/* BVH2 builder exposed to orc_shade.c (emitter-only BVH) */
void orc_build_bvh_public(const float* tris, uint32_t n, OrcBvhNode** nodes_out, uint32_t* num_nodes_out, uint32_t** order_out) {
  build_bvh(tris, n, nodes_out, num_nodes_out, order_out);
}
