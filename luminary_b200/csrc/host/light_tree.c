/*
 * light_tree.c - host-side build of the many-lights tree consumed by the NEE kernels.
 *
 * Host code stays in C (BASELINE.json north_star). This is the B200 framework's counterpart of the
 * reference's CPU builder (device/device_light.c): emissive triangles of all active instances become
 * "fragments" (:2020-2113), a binary tree is built top-down with a 32-bin power-weighted SAH sweep
 * (:270-486), every node gets the power-weighted mean / spatial variance of its two children (:488-584),
 * and the binary tree is collapsed into a root of up to 128 children (16 sections of 8) plus 8-wide nodes,
 * quantised to the 16/48/64-byte device records (:663-1153, layouts device_utils.h:283-327). The output is
 * exactly the payload of device_update_light_tree_data (device/device.h:171): root blob, node blob and the
 * TriangleHandle map that defines light ids.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/lumb200.h"

void lumb200_set_last_error(const char* fmt, ...);

#define LT_BINS 32
#define LT_ROOT_MAX_CHILDREN 128
#define LT_NODE_CHILDREN 8
#define LT_NULL 0xFFFFFFFFu
#define LT_MAX_VALUE 1e10f
#define LT_FRAGMENT_ERROR_COMP (FLT_EPSILON * 16.0f)

typedef struct {
  float x, y, z;
} f3;

static f3 f3_make(float x, float y, float z) {
  f3 r = {x, y, z};
  return r;
}
static f3 f3_add(f3 a, f3 b) { return f3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static f3 f3_sub(f3 a, f3 b) { return f3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static f3 f3_mul(f3 a, f3 b) { return f3_make(a.x * b.x, a.y * b.y, a.z * b.z); }
static f3 f3_scale(f3 a, float s) { return f3_make(a.x * s, a.y * s, a.z * s); }
static f3 f3_min(f3 a, f3 b) { return f3_make(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
static f3 f3_max(f3 a, f3 b) { return f3_make(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
static float f3_dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* vec128_hsum of a 3-vector (w = 0), host_intrinsics.h:97-107: the SSE shuffle adds (x + z) + (y + w) */
static float hsum3(float x, float y, float z) { return (x + z) + (y + 0.0f); }
static float hsum4(float x, float y, float z, float w) { return (x + z) + (y + w); }
/* vec128_dot = hsum(a * b), host_intrinsics.h:182-184 */
static float f3_dot_h(f3 a, f3 b) { return hsum3(a.x * b.x, a.y * b.y, a.z * b.z); }
static f3 f3_cross(f3 a, f3 b) { return f3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static float f3_axis(f3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

typedef struct {
  f3 lo, hi, middle;
  f3 v0, v1, v2;
  /* QUIRK: the reference rotates vertices with vec128_rotate_quaternion, which also scales the quaternion's w lane
   * (host_intrinsics.h:202-213), so the 4th lane of a rotated vertex holds q.w * 2 * dot(q, v) instead of 0. The lane survives
   * into LightTreeFragment.middle / v0 / v1 / v2 and the 4-lane vec128_dot of the spatial variance (device_light.c:551-558)
   * sums it. Carried here so that the std-dev bytes of rotated emitters match the reference. */
  float w0, w1, w2, middle_w;
  float power;
  uint32_t instance_id, tri_id;
} Fragment;

typedef struct {
  uint32_t first, count; /* fragment range */
  uint32_t child;        /* index of the left child, right = child + 1; LT_NULL for leaves */
  float left_power, right_power;
  /* traversal structure (device_light.c:586-636) */
  f3 left_mean, right_mean;
  float left_variance, right_variance;
} BinNode;

typedef struct {
  f3 mean;
  float variance, power;
  int is_leaf;
} Child;

/* rotation_euler_angles_to_quaternion, host_math.c:6-21 */
static void euler_to_quat(const float r[3], float q[4]) {
  const float cr = cosf(r[0] * 0.5f), sr = sinf(r[0] * 0.5f);
  const float cp = cosf(r[1] * 0.5f), sp = sinf(r[1] * 0.5f);
  const float cy = cosf(r[2] * 0.5f), sy = sinf(r[2] * 0.5f);
  q[3] = cr * cp * cy + sr * sp * sy;
  q[0] = sr * cp * cy - cr * sp * sy;
  q[1] = cr * sp * cy + sr * cp * sy;
  q[2] = cr * cp * sy - sr * sp * cy;
}

/* v' = q v q^-1 with q = (-x, -y, -z, w): the builder rotates by the conjugate, matching the quaternion16
 * convention of the device transforms (device_light.c:2027, device_structs.c:388-399) */
static f3 rotate_conj(const float q[4], f3 v, float* w_lane) {
  const f3 u       = f3_make(-q[0], -q[1], -q[2]);
  const float s    = q[3];
  const float d_uv = f3_dot(u, v), d_uu = f3_dot(u, u);
  *w_lane          = s * (2.0f * d_uv); /* see Fragment.w0 */
  f3 r = f3_scale(u, 2.0f * d_uv);
  r    = f3_add(r, f3_scale(v, s * s - d_uu));
  r    = f3_add(r, f3_scale(f3_cross(u, v), 2.0f * s));
  return r;
}

/* device_pack_float, device_packing.c:45-70: bfloat16 with directed rounding */
static uint16_t pack_bf16(float v, int mode /* 0 floor, 1 ceil */) {
  uint32_t b;
  memcpy(&b, &v, 4);
  if (mode == 1) {
    if (v >= 0.0f)
      b += (1u << 16) - 1;
  }
  else {
    if (v < 0.0f)
      b += (1u << 16) - 1;
  }
  return (uint16_t) (b >> 16);
}
static float unpack_bf16(uint16_t v) {
  const uint32_t b = ((uint32_t) v) << 16;
  float f;
  memcpy(&f, &b, 4);
  return f;
}

/* ---------------------------------------------------------------------------------------------- */
typedef struct {
  Fragment* frags;
  uint32_t num_frags;
  BinNode* nodes;
  uint32_t num_nodes, cap_nodes;
} Work;

static void fit_bounds(const Fragment* f, uint32_t n, f3* hi, f3* lo) {
  f3 h = f3_make(-LT_MAX_VALUE, -LT_MAX_VALUE, -LT_MAX_VALUE), l = f3_make(LT_MAX_VALUE, LT_MAX_VALUE, LT_MAX_VALUE);
  for (uint32_t i = 0; i < n; i++) {
    h = f3_max(h, f[i].hi);
    l = f3_min(l, f[i].lo);
  }
  *hi = h, *lo = l;
}

/* vec128_box_area, host_intrinsics.h:189-197: hsum of (x*y, x*z, y*z, 0) */
static float box_area(f3 d) { return hsum3(d.x * d.y, d.x * d.z, d.y * d.z); }

typedef struct {
  f3 hi, lo;
  uint32_t count;
  float power;
} Bin;

static int push_node(Work* w, BinNode n) {
  if (w->num_nodes == w->cap_nodes) {
    w->cap_nodes = w->cap_nodes * 2 + 16;
    BinNode* p   = (BinNode*) realloc(w->nodes, sizeof(BinNode) * w->cap_nodes);
    if (!p)
      return 0;
    w->nodes = p;
  }
  w->nodes[w->num_nodes++] = n;
  return 1;
}

static int build_binary(Work* w) {
  BinNode root;
  memset(&root, 0, sizeof(root));
  root.first = 0, root.count = w->num_frags, root.child = LT_NULL;
  if (!push_node(w, root))
    return 0;

  Bin bins[LT_BINS];
  uint32_t begin = 0, end = 1;
  while (begin != end) {
    for (uint32_t ni = begin; ni < end; ni++) {
      BinNode node = w->nodes[ni];
      if (node.count == 1)
        continue;
      Fragment* fr = w->frags + node.first;

      f3 ph, pl;
      fit_bounds(fr, node.count, &ph, &pl);
      const f3 pd          = f3_sub(ph, pl);
      const float max_axis = fmaxf(pd.x, fmaxf(pd.y, pd.z));

      double best_cost = DBL_MAX, best_plane = 0.0;
      int best_axis = -1;
      uint32_t best_split = 0;
      float best_lp = 0.0f, best_rp = 0.0f;

      for (int a = 0; a < 3; a++) {
        const double lo_a = f3_axis(pl, a), hi_a = f3_axis(ph, a);
        const double interval = (hi_a - lo_a) / LT_BINS;
        if (interval <= LT_FRAGMENT_ERROR_COMP * fabs(lo_a))
          continue;
        for (int b = 0; b < LT_BINS; b++) {
          bins[b].hi    = f3_make(-LT_MAX_VALUE, -LT_MAX_VALUE, -LT_MAX_VALUE);
          bins[b].lo    = f3_make(LT_MAX_VALUE, LT_MAX_VALUE, LT_MAX_VALUE);
          bins[b].count = 0;
          bins[b].power = 0.0f;
        }
        const double inv = 1.0 / interval;
        for (uint32_t i = 0; i < node.count; i++) {
          int pos = ((int) ceil((f3_axis(fr[i].middle, a) - lo_a) * inv)) - 1;
          if (pos < 0)
            pos = 0;
          if (pos >= LT_BINS)
            pos = LT_BINS - 1;
          bins[pos].count++;
          bins[pos].power += fr[i].power;
          bins[pos].hi = f3_max(bins[pos].hi, fr[i].hi);
          bins[pos].lo = f3_min(bins[pos].lo, fr[i].lo);
        }
        const double interval_cost = max_axis / interval;
        /* suffix boxes */
        f3 rh[LT_BINS + 1], rl[LT_BINS + 1];
        rh[LT_BINS] = f3_make(-LT_MAX_VALUE, -LT_MAX_VALUE, -LT_MAX_VALUE);
        rl[LT_BINS] = f3_make(LT_MAX_VALUE, LT_MAX_VALUE, LT_MAX_VALUE);
        for (int b = LT_BINS - 1; b >= 0; b--) {
          rh[b] = f3_max(rh[b + 1], bins[b].hi);
          rl[b] = f3_min(rl[b + 1], bins[b].lo);
        }
        float lp = 0.0f, rp = 0.0f;
        for (int b = 0; b < LT_BINS; b++)
          rp += bins[b].power;
        f3 lh = f3_make(-LT_MAX_VALUE, -LT_MAX_VALUE, -LT_MAX_VALUE), ll = f3_make(LT_MAX_VALUE, LT_MAX_VALUE, LT_MAX_VALUE);
        uint32_t left = 0;
        for (int k = 1; k < LT_BINS; k++) {
          lh = f3_max(lh, bins[k - 1].hi);
          ll = f3_min(ll, bins[k - 1].lo);
          lp += bins[k - 1].power;
          rp -= bins[k - 1].power;
          const float la = box_area(f3_sub(lh, ll));
          const float ra = box_area(f3_sub(rh[k], rl[k]));
          const double cost = interval_cost * (lp * la + rp * ra);
          left += bins[k - 1].count;
          if (left == 0 || left == node.count)
            continue;
          if (cost < best_cost) {
            best_cost  = cost;
            best_split = left;
            best_plane = lo_a + k * interval;
            best_axis  = a;
            best_lp    = lp;
            best_rp    = rp;
          }
        }
      }

      if (best_axis >= 0) {
        /* _light_tree_divide_middles_along_axis, device_light.c:245-268 */
        uint32_t left = 0, right = 0;
        while (left + right < node.count) {
          const Fragment f = fr[left];
          if ((double) f3_axis(f.middle, best_axis) > best_plane) {
            const uint32_t swap = node.count - 1 - right;
            fr[left]            = fr[swap];
            fr[swap]            = f;
            right++;
          }
          else
            left++;
        }
        /* the child sizes stay the bin counts of the sweep (optimal_split, device_light.c:361), even if the plane
         * partition above disagrees by a rounding - exactly as the reference does */
        (void) left;
      }
      if (best_axis < 0) {
        best_split = node.count / 2;
        best_lp = best_rp = 0.0f;
        for (uint32_t i = 0; i < node.count; i++) {
          if (i < best_split)
            best_lp += fr[i].power;
          else
            best_rp += fr[i].power;
        }
      }

      node.left_power  = best_lp;
      node.right_power = best_rp;
      node.child       = w->num_nodes;

      BinNode l, r;
      memset(&l, 0, sizeof(l));
      memset(&r, 0, sizeof(r));
      l.first = node.first, l.count = best_split, l.child = LT_NULL;
      r.first = node.first + best_split, r.count = node.count - best_split, r.child = LT_NULL;
      if (!push_node(w, l) || !push_node(w, r))
        return 0;
      w->nodes[ni] = node;
    }
    begin = end;
    end   = w->num_nodes;
  }
  return 1;
}

/* _lights_get_vmf_and_mean_and_variance, device_light.c:488-584 */
static void mean_and_variance(const Work* w, const BinNode* n, float parent_power, float* power, f3* mean, float* variance) {
  const Fragment* fr = w->frags + n->first;
  if (*power < parent_power * 1e-5f) {
    float p = 0.0f;
    for (uint32_t i = 0; i < n->count; i++)
      p += fr[i].power;
    *power = p;
  }
  const float inv = 1.0f / *power;
  f3 p            = f3_make(0, 0, 0);
  float pw        = 0.0f; /* 4th lane of the mean, see Fragment.w0 */
  for (uint32_t i = 0; i < n->count; i++)
  {
    /* vec128_fmadd(frag.middle, weight, p): the one explicit FMA of the reference's builder (device_light.c:531) */
    const float wgt = fr[i].power * inv;
    p               = f3_make(fmaf(fr[i].middle.x, wgt, p.x), fmaf(fr[i].middle.y, wgt, p.y), fmaf(fr[i].middle.z, wgt, p.z));
    pw              = fmaf(fr[i].middle_w, wgt, pw);
  }
  float var = 0.0f;
  for (uint32_t i = 0; i < n->count; i++) {
    const float wgt = (1.0f / 3.0f) * fr[i].power * inv;
    const f3 d0 = f3_sub(fr[i].v0, p), d1 = f3_sub(fr[i].v1, p), d2 = f3_sub(fr[i].v2, p);
    const float e0 = fr[i].w0 - pw, e1 = fr[i].w1 - pw, e2 = fr[i].w2 - pw;
    var += wgt * hsum4(d0.x * d0.x, d0.y * d0.y, d0.z * d0.z, e0 * e0);
    var += wgt * hsum4(d1.x * d1.x, d1.y * d1.y, d1.z * d1.z, e1 * e1);
    var += wgt * hsum4(d2.x * d2.x, d2.y * d2.y, d2.z * d2.z, e2 * e2);
  }
  *mean     = p;
  *variance = var;
}

typedef struct {
  uint32_t* jobs; /* binary node index per device node */
  uint32_t num_jobs, cap_jobs;
  uint32_t* new_order; /* new light id -> old fragment index */
  uint32_t lights_ptr;
} Collapse;

static int push_job(Collapse* c, uint32_t v) {
  if (c->num_jobs == c->cap_jobs) {
    c->cap_jobs = c->cap_jobs * 2 + 16;
    uint32_t* p = (uint32_t*) realloc(c->jobs, sizeof(uint32_t) * c->cap_jobs);
    if (!p)
      return 0;
    c->jobs = p;
  }
  c->jobs[c->num_jobs++] = v;
  return 1;
}

/* _light_tree_collapse_binary_node, device_light.c:663-848 */
static void collapse_node(const Work* w, Collapse* c, const BinNode* base, Child* children, uint32_t* child_bin, uint32_t max_children, uint32_t* light_ptr,
                          uint32_t* child_count_out, uint32_t* leaf_count_out) {
  uint32_t count = 0;
  int work       = 0;
  const BinNode* nodes = w->nodes;
  if (base->count > 1) {
    children[count].mean = base->left_mean, children[count].variance = base->left_variance, children[count].power = base->left_power;
    children[count].is_leaf = 0;
    child_bin[count++]      = base->child;
    children[count].mean = base->right_mean, children[count].variance = base->right_variance, children[count].power = base->right_power;
    children[count].is_leaf = 0;
    child_bin[count++]      = base->child + 1;
    work                    = count < max_children;
  }
  else { /* single light in the scene */
    memset(&children[0], 0, sizeof(Child));
    children[0].is_leaf = 1;
    children[0].power   = 1.0f;
    child_bin[count++]  = 0;
  }

  while (work) {
    work              = 0;
    float best_cost   = 0.0f;
    uint32_t selected = 0;
    for (uint32_t k = 0; k < max_children; k++) {
      const uint32_t b = child_bin[k];
      if (b == LT_NULL)
        continue;
      const BinNode* bn = &nodes[b];
      if (bn->count == 1)
        continue;
      const float cost = (bn->left_power + bn->right_power) * (bn->left_variance + bn->right_variance);
      if (cost > best_cost) {
        best_cost = cost;
        selected  = k;
        work      = 1;
      }
    }
    if (!work)
      break;
    const BinNode* bn = &nodes[child_bin[selected]];
    Child l, r;
    l.mean = bn->left_mean, l.variance = bn->left_variance, l.power = bn->left_power, l.is_leaf = 0;
    r.mean = bn->right_mean, r.variance = bn->right_variance, r.power = bn->right_power, r.is_leaf = 0;
    child_bin[selected] = bn->child;
    children[selected]  = l;
    uint32_t slot       = 0;
    for (; slot < max_children; slot++)
      if (child_bin[slot] == LT_NULL)
        break;
    child_bin[slot] = bn->child + 1;
    children[slot]  = r;
    count++;
    if (count == max_children)
      break;
  }

  /* children are appended into the first free slot, so they are already dense in [0, count) */
  uint32_t leaves = 0;
  for (uint32_t k = 0; k < count; k++) {
    const BinNode* bn = &nodes[child_bin[k]];
    if (bn->count == 1) {
      if (*light_ptr == LT_NULL)
        *light_ptr = c->lights_ptr;
      c->new_order[c->lights_ptr++] = bn->first;
      children[k].is_leaf           = 1;
      child_bin[k]                  = LT_NULL;
      leaves++;
    }
  }
  /* stable partition: leaves first, order of the leaves preserved */
  for (uint32_t k = 0; k < leaves; k++) {
    if (!children[k].is_leaf) {
      uint32_t s = k + 1;
      for (; s < count; s++)
        if (children[s].is_leaf)
          break;
      const Child tc    = children[k];
      children[k]       = children[s];
      children[s]       = tc;
      const uint32_t tb = child_bin[k];
      child_bin[k]      = child_bin[s];
      child_bin[s]      = tb;
    }
  }
  *child_count_out = count;
  *leaf_count_out  = leaves;
}

typedef struct {
  f3 min_mean;
  int8_t ex, ey, ez, ev;
  float cx, cy, cz, cv, max_power;
} Frame;

static int8_t exp_for(float hi, float lo) { return (hi != lo) ? (int8_t) ceilf(log2f((hi - lo) * 1.0f / 255.0f)) : 0; }

static Frame make_frame(const Child* ch, uint32_t n, uint16_t* px, uint16_t* py, uint16_t* pz) {
  f3 mn = f3_make(LT_MAX_VALUE, LT_MAX_VALUE, LT_MAX_VALUE), mx = f3_make(-LT_MAX_VALUE, -LT_MAX_VALUE, -LT_MAX_VALUE);
  float max_var = 0.0f, max_pow = 0.0f;
  for (uint32_t k = 0; k < n; k++) {
    mn      = f3_min(mn, ch[k].mean);
    mx      = f3_max(mx, ch[k].mean);
    max_var = fmaxf(max_var, ch[k].variance);
    max_pow = fmaxf(max_pow, ch[k].power);
  }
  const float max_std = sqrtf(max_var);
  Frame f;
  *px = pack_bf16(mn.x, 0), *py = pack_bf16(mn.y, 0), *pz = pack_bf16(mn.z, 0);
  f.min_mean = f3_make(unpack_bf16(*px), unpack_bf16(*py), unpack_bf16(*pz));
  f.ex       = exp_for(mx.x, f.min_mean.x);
  f.ey       = exp_for(mx.y, f.min_mean.y);
  f.ez       = exp_for(mx.z, f.min_mean.z);
  /* the reference takes log2 of 0 for a single point light set; pick exponent 0 there */
  f.ev        = (max_std > 0.0f) ? (int8_t) ceilf(log2f(max_std * 1.0f / 255.0f)) : 0;
  f.cx        = 1.0f / exp2f(f.ex);
  f.cy        = 1.0f / exp2f(f.ey);
  f.cz        = 1.0f / exp2f(f.ez);
  f.cv        = 1.0f / exp2f(f.ev);
  f.max_power = max_pow;
  return f;
}

static uint32_t clamp_u(float v, uint32_t hi) {
  if (!(v > 0.0f))
    return 0;
  return (v >= (float) hi) ? hi : (uint32_t) v;
}

/* device records, device_utils.h:283-327 */
#pragma pack(push, 1)
typedef struct {
  uint16_t x, y, z, num_root_lights, power_normalization;
  uint8_t num_sections, padding1;
  int8_t exp_x, exp_y, exp_z, exp_std_dev;
} RootHeader;
typedef struct {
  uint8_t rel_mean_x[8], rel_mean_y[8], rel_mean_z[8], rel_std_dev[8];
  uint16_t rel_power[8];
} RootSection;
typedef struct {
  uint16_t x, y, z, padding;
  int8_t exp_x, exp_y, exp_z, exp_std_dev;
  uint8_t num_lights, padding1;
  uint16_t padding2;
  uint32_t child_ptr, light_ptr;
  uint8_t rel_mean_x[8], rel_mean_y[8], rel_mean_z[8], rel_std_dev[8], rel_power[8];
} TreeNode;
#pragma pack(pop)

typedef char assert_root_header[(sizeof(RootHeader) == 16) ? 1 : -1];
typedef char assert_root_section[(sizeof(RootSection) == 48) ? 1 : -1];
typedef char assert_tree_node[(sizeof(TreeNode) == 64) ? 1 : -1];

Lumb200Result lumb200_host_build_light_tree(
  const Lumb200Mesh* meshes, uint32_t num_meshes, const Lumb200Instance* instances, uint32_t num_instances, const Lumb200Material* materials,
  uint32_t num_materials, Lumb200LightTreeBuffers* out) {
  return lumb200_host_build_light_tree_textured(meshes, num_meshes, instances, num_instances, materials, num_materials, NULL, out);
}

Lumb200Result lumb200_host_build_light_tree_textured(
  const Lumb200Mesh* meshes, uint32_t num_meshes, const Lumb200Instance* instances, uint32_t num_instances, const Lumb200Material* materials,
  uint32_t num_materials, const float* const* triangle_intensities, Lumb200LightTreeBuffers* out) {
  if (!out || (!meshes && num_meshes) || (!instances && num_instances) || (!materials && num_materials)) {
    lumb200_set_last_error("NULL argument");
    return LUMB200_ERROR_ARGUMENT_NULL;
  }
  memset(out, 0, sizeof(*out));

  /* ---- fragments (device_light.c:2020-2113, 2170-2212) ---- */
  Work w;
  memset(&w, 0, sizeof(w));
  uint32_t cap = 0;
  /* The reference caches the triangles of a mesh grouped by "material slot" - the materials of the mesh in order of first
   * appearance (device_light.c:1646-1690) - and emits an instance's fragments slot by slot (:2047-2110). The initial fragment
   * order decides ties of the in-place partitions, so it is reproduced: per mesh, triangle ids stably sorted by slot. */
  uint32_t** mesh_order = (uint32_t**) calloc(num_meshes ? num_meshes : 1, sizeof(uint32_t*));
  if (!mesh_order) {
    lumb200_set_last_error("out of memory");
    return LUMB200_ERROR_OUT_OF_MEMORY;
  }
#define LT_FREE_MESH_ORDER()                 \
  do {                                       \
    for (uint32_t _m = 0; _m < num_meshes; _m++) \
      free(mesh_order[_m]);                  \
    free(mesh_order);                        \
  } while (0)
  for (uint32_t i = 0; i < num_instances; i++) {
    const Lumb200Instance* in = &instances[i];
    if (!in->active)
      continue;
    if (in->mesh_id >= num_meshes) {
      free(w.frags);
      LT_FREE_MESH_ORDER();
      lumb200_set_last_error("instance %u references mesh %u which does not exist", i, in->mesh_id);
      return LUMB200_ERROR_INVALID_API_ARGUMENT;
    }
    const Lumb200Mesh* m = &meshes[in->mesh_id];
    float q[4];
    euler_to_quat(in->rotation, q);
    const f3 scale = f3_make(in->scale[0], in->scale[1], in->scale[2]);
    const f3 offs  = f3_make(in->translation[0], in->translation[1], in->translation[2]);
    if (!mesh_order[in->mesh_id]) {
      uint32_t* slot_of  = (uint32_t*) malloc(sizeof(uint32_t) * 65536);
      uint32_t* slot_cnt = (uint32_t*) calloc(65536 + 1, sizeof(uint32_t));
      uint32_t* order    = (uint32_t*) malloc(sizeof(uint32_t) * (m->triangle_count ? m->triangle_count : 1));
      if (!slot_of || !slot_cnt || !order) {
        free(slot_of), free(slot_cnt), free(order), free(w.frags);
        LT_FREE_MESH_ORDER();
        lumb200_set_last_error("out of memory");
        return LUMB200_ERROR_OUT_OF_MEMORY;
      }
      memset(slot_of, 0xFF, sizeof(uint32_t) * 65536);
      uint32_t num_slots = 0;
      for (uint32_t t = 0; t < m->triangle_count; t++) {
        const uint16_t mid = m->material_id_buffer[t];
        if (slot_of[mid] == 0xFFFFFFFFu)
          slot_of[mid] = num_slots++;
        slot_cnt[slot_of[mid] + 1]++;
      }
      for (uint32_t k = 0; k < num_slots; k++)
        slot_cnt[k + 1] += slot_cnt[k];
      for (uint32_t t = 0; t < m->triangle_count; t++)
        order[slot_cnt[slot_of[m->material_id_buffer[t]]]++] = t;
      free(slot_of), free(slot_cnt);
      mesh_order[in->mesh_id] = order;
    }
    const uint32_t* order = mesh_order[in->mesh_id];
    for (uint32_t ti = 0; ti < m->triangle_count; ti++) {
      const uint32_t t   = order[ti];
      const uint16_t mid = m->material_id_buffer[t];
      if (mid >= num_materials)
        continue;
      const Lumb200Material* mat = &materials[mid];
      if (!mat->emission_active)
        continue;
      /* _light_tree_update_cache_material, device_light.c:1821-1862: a luminance-textured material has the constant
       * intensity emission_scale, multiplied per triangle by the integrated texture intensity (:2082-2094), which is 1 until
       * the integration has run (:1686) */
      float intensity = fmaxf(mat->emission[0], fmaxf(mat->emission[1], mat->emission[2]));
      if (mat->luminance_tex != LUMB200_TEXTURE_NONE) {
        intensity = mat->emission_scale;
        if (intensity > 0.0f && triangle_intensities && triangle_intensities[in->mesh_id]) {
          const float tri_intensity = triangle_intensities[in->mesh_id][t];
          if (tri_intensity == 0.0f)
            continue;
          intensity *= tri_intensity;
        }
      }
      if (!(intensity > 0.0f))
        continue;
      const float* vb = m->vertex_buffer + 9 * (size_t) t;
      float w0, w1, w2;
      const f3 v0     = f3_add(f3_mul(rotate_conj(q, f3_make(vb[0], vb[1], vb[2]), &w0), scale), offs);
      const f3 v1     = f3_add(f3_mul(rotate_conj(q, f3_make(vb[3], vb[4], vb[5]), &w1), scale), offs);
      const f3 v2     = f3_add(f3_mul(rotate_conj(q, f3_make(vb[6], vb[7], vb[8]), &w2), scale), offs);
      w0 = w0 * 1.0f + 0.0f, w1 = w1 * 1.0f + 0.0f, w2 = w2 * 1.0f + 0.0f; /* scale.w = 1, offset.w = 0 */
      const f3 cr     = f3_cross(f3_sub(v1, v0), f3_sub(v2, v0));
      const float area = 0.5f * sqrtf(f3_dot_h(cr, cr)); /* vec128_norm2 */
      if (area == 0.0f)
        continue;
      if (w.num_frags == cap) {
        cap         = cap * 2 + 64;
        Fragment* p = (Fragment*) realloc(w.frags, sizeof(Fragment) * cap);
        if (!p) {
          free(w.frags);
          LT_FREE_MESH_ORDER();
          lumb200_set_last_error("out of memory");
          return LUMB200_ERROR_OUT_OF_MEMORY;
        }
        w.frags = p;
      }
      Fragment f;
      f.lo = f3_min(v0, f3_min(v1, v2));
      f.hi = f3_max(v0, f3_max(v1, v2));
      f.middle = f3_scale(f3_add(v0, f3_add(v1, v2)), 1.0f / 3.0f);
      f.v0 = v0, f.v1 = v1, f.v2 = v2;
      f.w0 = w0, f.w1 = w1, f.w2 = w2;
      f.middle_w    = (w0 + (w1 + w2)) * (1.0f / 3.0f);
      f.power       = intensity * area;
      f.instance_id = i;
      f.tri_id      = t;
      w.frags[w.num_frags++] = f;
    }
  }
  LT_FREE_MESH_ORDER();
#undef LT_FREE_MESH_ORDER
  if (w.num_frags == 0) {
    free(w.frags);
    return LUMB200_SUCCESS; /* no lights: empty tree (LIGHTS_ARE_PRESENT false) */
  }

  /* ---- binary tree + traversal structure ---- */
  if (!build_binary(&w)) {
    free(w.frags);
    free(w.nodes);
    lumb200_set_last_error("out of memory");
    return LUMB200_ERROR_OUT_OF_MEMORY;
  }
  for (uint32_t i = 0; i < w.num_nodes; i++) {
    BinNode* n = &w.nodes[i];
    if (n->child == LT_NULL)
      continue;
    const float parent_power = n->left_power + n->right_power;
    mean_and_variance(&w, &w.nodes[n->child], parent_power, &n->left_power, &n->left_mean, &n->left_variance);
    mean_and_variance(&w, &w.nodes[n->child + 1], parent_power, &n->right_power, &n->right_mean, &n->right_variance);
  }

  /* ---- collapse ---- */
  Collapse c;
  memset(&c, 0, sizeof(c));
  c.new_order = (uint32_t*) malloc(sizeof(uint32_t) * w.num_frags);
  memset(c.new_order, 0xFF, sizeof(uint32_t) * w.num_frags);

  uint8_t* root_blob = (uint8_t*) calloc(1, 16 + 48 * (LT_ROOT_MAX_CHILDREN / 8));
  TreeNode* tnodes   = NULL;
  uint32_t num_tnodes = 0, cap_tnodes = 0;
  size_t root_size = 0;
  int ok           = 1;

  {
    Child children[LT_ROOT_MAX_CHILDREN];
    uint32_t child_bin[LT_ROOT_MAX_CHILDREN];
    for (uint32_t k = 0; k < LT_ROOT_MAX_CHILDREN; k++)
      child_bin[k] = LT_NULL;
    uint32_t light_ptr = LT_NULL, count = 0, leaves = 0;
    collapse_node(&w, &c, &w.nodes[0], children, child_bin, LT_ROOT_MAX_CHILDREN, &light_ptr, &count, &leaves);

    RootHeader* h = (RootHeader*) root_blob;
    const Frame f = make_frame(children, count, &h->x, &h->y, &h->z);
    h->exp_x = f.ex, h->exp_y = f.ey, h->exp_z = f.ez, h->exp_std_dev = f.ev;
    h->num_sections        = (uint8_t) ((count + 7) / 8);
    h->num_root_lights     = (uint16_t) leaves;
    h->power_normalization = pack_bf16(f.max_power, 1);
    RootSection* sec       = (RootSection*) (root_blob + 16);
    for (uint32_t k = 0; k < count; k++) {
      RootSection* s     = &sec[k / 8];
      const uint32_t j   = k % 8;
      s->rel_mean_x[j]   = (uint8_t) clamp_u(floorf((children[k].mean.x - f.min_mean.x) * f.cx + 0.5f), 255);
      s->rel_mean_y[j]   = (uint8_t) clamp_u(floorf((children[k].mean.y - f.min_mean.y) * f.cy + 0.5f), 255);
      s->rel_mean_z[j]   = (uint8_t) clamp_u(floorf((children[k].mean.z - f.min_mean.z) * f.cz + 0.5f), 255);
      uint32_t sd        = clamp_u(sqrtf(children[k].variance) * f.cv + 0.5f, 255);
      uint32_t pw        = clamp_u(floorf(0xFFFF * children[k].power / f.max_power + 0.5f), 0xFFFF);
      s->rel_std_dev[j]  = (uint8_t) (sd < 1 ? 1 : sd);
      s->rel_power[j]    = (uint16_t) (pw < 1 ? 1 : pw);
    }
    root_size = 16 + 48 * (size_t) h->num_sections;
    for (uint32_t k = 0; k < count; k++)
      if (child_bin[k] != LT_NULL)
        ok = ok && push_job(&c, child_bin[k]);
  }

  for (uint32_t job = 0; ok && job < c.num_jobs; job++) {
    Child children[LT_NODE_CHILDREN];
    uint32_t child_bin[LT_NODE_CHILDREN];
    for (uint32_t k = 0; k < LT_NODE_CHILDREN; k++)
      child_bin[k] = LT_NULL;
    uint32_t light_ptr = LT_NULL, count = 0, leaves = 0;
    collapse_node(&w, &c, &w.nodes[c.jobs[job]], children, child_bin, LT_NODE_CHILDREN, &light_ptr, &count, &leaves);

    TreeNode n;
    memset(&n, 0, sizeof(n));
    n.child_ptr  = c.num_jobs;
    n.light_ptr  = light_ptr;
    n.num_lights = (uint8_t) leaves;
    const Frame f = make_frame(children, count, &n.x, &n.y, &n.z);
    n.exp_x = f.ex, n.exp_y = f.ey, n.exp_z = f.ez, n.exp_std_dev = f.ev;
    for (uint32_t k = 0; k < count; k++) {
      n.rel_mean_x[k]  = (uint8_t) clamp_u(floorf((children[k].mean.x - f.min_mean.x) * f.cx + 0.5f), 255);
      n.rel_mean_y[k]  = (uint8_t) clamp_u(floorf((children[k].mean.y - f.min_mean.y) * f.cy + 0.5f), 255);
      n.rel_mean_z[k]  = (uint8_t) clamp_u(floorf((children[k].mean.z - f.min_mean.z) * f.cz + 0.5f), 255);
      uint32_t sd      = clamp_u(sqrtf(children[k].variance) * f.cv + 0.5f, 255);
      uint32_t pw      = clamp_u(floorf(0xFF * children[k].power / f.max_power + 0.5f), 255);
      n.rel_std_dev[k] = (uint8_t) (sd < 1 ? 1 : sd);
      n.rel_power[k]   = (uint8_t) (pw < 1 ? 1 : pw);
    }
    if (num_tnodes == cap_tnodes) {
      cap_tnodes  = cap_tnodes * 2 + 64;
      TreeNode* p = (TreeNode*) realloc(tnodes, sizeof(TreeNode) * cap_tnodes);
      if (!p) {
        ok = 0;
        break;
      }
      tnodes = p;
    }
    tnodes[num_tnodes++] = n;
    for (uint32_t k = 0; k < count; k++)
      if (child_bin[k] != LT_NULL)
        ok = ok && push_job(&c, child_bin[k]);
  }

  if (ok && c.lights_ptr != w.num_frags)
    ok = 0;

  uint32_t* handles = ok ? (uint32_t*) malloc(sizeof(uint32_t) * 2 * w.num_frags) : NULL;
  if (handles) {
    for (uint32_t l = 0; l < w.num_frags; l++) {
      const Fragment* f  = &w.frags[c.new_order[l]];
      handles[2 * l + 0] = f->instance_id;
      handles[2 * l + 1] = f->tri_id;
    }
  }
  const uint32_t num_lights = w.num_frags;
  free(w.frags);
  free(w.nodes);
  free(c.jobs);
  free(c.new_order);
  if (!handles) {
    free(root_blob);
    free(tnodes);
    lumb200_set_last_error("light tree collapse failed (a light was lost or memory ran out)");
    return LUMB200_ERROR_API_EXCEPTION;
  }
  out->root_data      = root_blob;
  out->root_size      = root_size;
  out->nodes_data     = tnodes;
  out->nodes_size     = sizeof(TreeNode) * (size_t) num_tnodes;
  out->tri_handle_map = handles;
  out->num_lights     = num_lights;
  return LUMB200_SUCCESS;
}

void lumb200_host_free_light_tree(Lumb200LightTreeBuffers* t) {
  if (!t)
    return;
  free(t->root_data);
  free(t->nodes_data);
  free(t->tri_handle_map);
  memset(t, 0, sizeof(*t));
}
