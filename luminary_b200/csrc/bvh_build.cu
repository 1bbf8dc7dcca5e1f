// bvh_build.cu - on-device acceleration structure build.
//
// Replaces the reference's OptiX accel builds (device/optix_bvh.c:185-478: per-mesh GAS, instance IAS, light
// GAS). B200 has no RT cores, so instead of a two-level instanced structure every instance is flattened to
// world space (180 GB of HBM makes the duplication irrelevant; it removes the per-instance ray transform from
// the traversal loop) and ONE compressed 8-wide BVH is built on the device:
//   1. k_flatten        world vertices = transform_apply(instance, v)   (reference cuda/math.cuh:459-491)
//   2. k_bounds/k_morton 63-bit Morton codes of padded primitive boxes
//   3. cub radix sort   (library call; build path only, not the per-bounce hot path)
//   4. k_ploc_*         binary hierarchy by parallel locally-ordered clustering (Meister & Bittner 2018): clusters
//                       in Morton order repeatedly merge with their mutual nearest neighbour (surface area of the
//                       union, search radius 16). ~30 % fewer node visits per ray than the Karras 2012 radix tree
//                       (k_hierarchy + k_refit), which is kept as LUMB200_BVH_BUILDER=lbvh for comparison
//   4c. k_reinsert_*    parallel reinsertion (Meister & Bittner 2018) on that hierarchy: nodes move to where they shrink the
//                       summed inner-node area most, until a pass gains < 0.3 % (SAH of the collapsed tree -8 % atrium, -21 % terrain)
//   5. k_collapse       level-synchronous greedy surface-area collapse to 8-wide nodes, octant slot
//                       assignment, 8-bit quantisation with one cell of padding, triangles re-laid per node
// Compiled with -fmad=false: world vertices must be bit-identical to the CPU oracle's.
#include <cub/cub.cuh>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "lumb200_internal.cuh"

#ifndef LB_QUANT_MARGIN
#define LB_QUANT_MARGIN (1.0f / 64.0f)
#endif
#define BUILD_THREADS 256
#define LEAF_MAX_TRIS 3
// cost of one triangle test relative to one node step in the SAH of the collapse (sweep on B200, profiles/)
#ifndef LB_SAH_C_PRIM
#define LB_SAH_C_PRIM 0.5f
#endif
// tree-rotation passes over the PLOC hierarchy before the collapse (0 = off)
#ifndef LB_BVH_ROTATION_PASSES
#define LB_BVH_ROTATION_PASSES 0
#endif
// parallel-reinsertion passes over the selected PLOC hierarchy before the collapse (0 = off; LUMB200_BVH_REINSERT overrides)
#ifndef LB_BVH_REINSERT_PASSES
#define LB_BVH_REINSERT_PASSES 32
#endif
// a reinsertion pass that shrinks the summed inner-node area by less than this fraction is the last one (LUMB200_BVH_REINSERT_MIN_GAIN)
#ifndef LB_BVH_REINSERT_MIN_GAIN
#define LB_BVH_REINSERT_MIN_GAIN 0.003f
#endif

// ---------------------------------------------------------------------------------------------
// 1. flatten
// ---------------------------------------------------------------------------------------------
__global__ void k_flatten(LbSceneTables tab, float4* __restrict__ world, uint2* __restrict__ handle) {
  const uint32_t prim = blockIdx.x * blockDim.x + threadIdx.x;
  if (prim >= tab.num_prims)
    return;

  // largest instance i with offset[i] <= prim
  uint32_t lo = 0, hi = tab.num_instances;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (tab.instance_prim_offset[mid] <= prim)
      lo = mid;
    else
      hi = mid;
  }
  const uint32_t inst = lo;
  const uint32_t tri  = prim - tab.instance_prim_offset[inst];
  const uint32_t mesh = tab.instance_mesh[inst];

  const LbTransform tr = tab.instance_transform[inst];
  const float4* vb     = tab.mesh_vertices[mesh];

#pragma unroll
  for (int v = 0; v < 3; v++) {
    const float4 p = vb[3 * (size_t) tri + v];
    const V3 w     = transform_point(tr, v3(p.x, p.y, p.z));
    world[3 * (size_t) prim + v] = make_float4(w.x, w.y, w.z, (v == 0) ? __uint_as_float(prim) : 0.0f);
  }
  handle[prim] = make_uint2(inst, tri);
}

Lumb200Result lb_flatten_instances(const LbSceneTables& tables, float4* world_tris, uint2* prim_handle, cudaStream_t stream) {
  if (tables.num_prims == 0)
    return LUMB200_SUCCESS;
  const uint32_t blocks = (tables.num_prims + BUILD_THREADS - 1) / BUILD_THREADS;
  k_flatten<<<blocks, BUILD_THREADS, 0, stream>>>(tables, world_tris, prim_handle);
  LB_CHECK(cudaGetLastError());
  return LUMB200_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// 2. primitive boxes, scene bounds, Morton codes
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return (i >= 0) ? i : (i ^ 0x7FFFFFFF);
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float((i >= 0) ? i : (i ^ 0x7FFFFFFF)); }

struct BuildBounds {
  int lo[3];
  int hi[3];
};

__global__ void k_bounds_init(BuildBounds* b) {
  for (int k = 0; k < 3; k++) {
    b->lo[k] = float_to_ordered(FLT_MAX);
    b->hi[k] = float_to_ordered(-FLT_MAX);
  }
}

// Primitive box with a relative pad: the fp32 triangle test can accept a ray that passes a few ulp outside
// the exact triangle, the box must not cull it. Same pad in the oracle's BVH2 is unnecessary (its boxes are
// not quantised and its slab test is scaled by 1 + 4 ulp), here it is folded into the quantisation margin.
__global__ void k_prim_boxes(const float4* __restrict__ world, uint32_t n, float4* __restrict__ box_lo, float4* __restrict__ box_hi,
                             BuildBounds* bounds) {
  const uint32_t prim = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  if (prim < n) {
    const float4 a = world[3 * (size_t) prim + 0];
    const float4 b = world[3 * (size_t) prim + 1];
    const float4 c = world[3 * (size_t) prim + 2];
    lo[0] = fminf(a.x, fminf(b.x, c.x));
    lo[1] = fminf(a.y, fminf(b.y, c.y));
    lo[2] = fminf(a.z, fminf(b.z, c.z));
    hi[0] = fmaxf(a.x, fmaxf(b.x, c.x));
    hi[1] = fmaxf(a.y, fmaxf(b.y, c.y));
    hi[2] = fmaxf(a.z, fmaxf(b.z, c.z));
    for (int k = 0; k < 3; k++) {
      const float pad = 4e-7f * fmaxf(fabsf(lo[k]), fabsf(hi[k]));
      lo[k] -= pad;
      hi[k] += pad;
    }
    box_lo[prim] = make_float4(lo[0], lo[1], lo[2], 0.0f);
    box_hi[prim] = make_float4(hi[0], hi[1], hi[2], 0.0f);
  }
  // warp reduce then one atomic per warp
  for (int k = 0; k < 3; k++) {
    float l = lo[k], h = hi[k];
    for (int o = 16; o > 0; o >>= 1) {
      l = fminf(l, __shfl_xor_sync(0xFFFFFFFFu, l, o));
      h = fmaxf(h, __shfl_xor_sync(0xFFFFFFFFu, h, o));
    }
    if ((threadIdx.x & 31) == 0 && l <= h) {
      atomicMin(&bounds->lo[k], float_to_ordered(l));
      atomicMax(&bounds->hi[k], float_to_ordered(h));
    }
  }
}

__device__ __forceinline__ uint64_t expand21(uint64_t v) {
  v &= 0x1FFFFFull;
  v = (v | (v << 32)) & 0x1F00000000FFFFull;
  v = (v | (v << 16)) & 0x1F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}

__global__ void k_morton(const float4* __restrict__ box_lo, const float4* __restrict__ box_hi, uint32_t n, const BuildBounds* bounds,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t prim = blockIdx.x * blockDim.x + threadIdx.x;
  if (prim >= n)
    return;
  const float4 lo = box_lo[prim], hi = box_hi[prim];
  const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
  uint64_t q[3];
  for (int k = 0; k < 3; k++) {
    const float bl  = ordered_to_float(bounds->lo[k]);
    const float bh  = ordered_to_float(bounds->hi[k]);
    const float ext = bh - bl;
    float f         = (ext > 0.0f) ? (c[k] - bl) / ext : 0.0f;
    f               = fminf(fmaxf(f, 0.0f), 1.0f);
    q[k]            = (uint64_t) fminf(f * 2097152.0f, 2097151.0f);
  }
  keys[prim] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
  vals[prim] = prim;
}

// ---------------------------------------------------------------------------------------------
// 4. binary radix tree (Karras 2012) + bottom-up refit
// ---------------------------------------------------------------------------------------------
#define LEAF_FLAG 0x80000000u

struct Bvh2 {
  uint32_t* left;    // [n-1] child refs (LEAF_FLAG | sorted position, or internal index)
  uint32_t* right;   // [n-1]
  uint32_t* first;   // [n-1] range in sorted order
  uint32_t* last;    // [n-1]
  uint32_t* parent;  // [n-1] parent of internal node (root: 0xFFFFFFFF)
  uint32_t* leaf_parent;  // [n]
  float4* lo;        // [n-1]
  float4* hi;        // [n-1]
  uint32_t* flags;   // [n-1] arrival counters
  uint32_t* count;   // [n-1] primitives below the node
};

__device__ __forceinline__ int delta_fn(const uint64_t* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n)
    return -1;
  const uint64_t a = keys[i], b = keys[j];
  if (a == b)
    return 64 + __clz((uint32_t) i ^ (uint32_t) j);
  return __clzll((long long) (a ^ b));
}

__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, Bvh2 t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1)
    return;

  const int d    = (delta_fn(keys, n, i, i + 1) - delta_fn(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = delta_fn(keys, n, i, i - d);
  int lmax       = 2;
  while (delta_fn(keys, n, i, i + lmax * d) > dmin)
    lmax <<= 1;
  int l = 0;
  for (int s = lmax >> 1; s >= 1; s >>= 1)
    if (delta_fn(keys, n, i, i + (l + s) * d) > dmin)
      l += s;
  const int j     = i + l * d;
  const int dnode = delta_fn(keys, n, i, j);
  int s           = 0;
  int tt          = l;
  do {
    tt = (tt + 1) >> 1;
    if (delta_fn(keys, n, i, i + (s + tt) * d) > dnode)
      s += tt;
  } while (tt > 1);
  const int gamma = i + s * d + min(d, 0);

  const int lo_i = min(i, j), hi_i = max(i, j);
  uint32_t left, right;
  if (lo_i == gamma) {
    left                 = LEAF_FLAG | (uint32_t) gamma;
    t.leaf_parent[gamma] = i;
  }
  else {
    left            = gamma;
    t.parent[gamma] = i;
  }
  if (hi_i == gamma + 1) {
    right                    = LEAF_FLAG | (uint32_t) (gamma + 1);
    t.leaf_parent[gamma + 1] = i;
  }
  else {
    right               = gamma + 1;
    t.parent[gamma + 1] = i;
  }
  t.left[i]  = left;
  t.right[i] = right;
  t.first[i] = lo_i;
  t.last[i]  = hi_i;
  t.count[i] = (uint32_t) (hi_i - lo_i + 1);
  if (i == 0)
    t.parent[0] = 0xFFFFFFFFu;
}

__global__ void k_refit(int n, Bvh2 t, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ box_lo,
                        const float4* __restrict__ box_hi) {
  const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n)
    return;
  uint32_t node = t.leaf_parent[leaf];
  while (node != 0xFFFFFFFFu) {
    __threadfence();
    if (atomicAdd(&t.flags[node], 1u) == 0)
      return;  // first arrival: the sibling subtree is not finished yet
    const uint32_t l = t.left[node], r = t.right[node];
    float4 llo, lhi, rlo, rhi;
    if (l & LEAF_FLAG) {
      const uint32_t p = sorted_prim[l & ~LEAF_FLAG];
      llo              = box_lo[p];
      lhi              = box_hi[p];
    }
    else {
      llo = __ldcg(&t.lo[l]);
      lhi = __ldcg(&t.hi[l]);
    }
    if (r & LEAF_FLAG) {
      const uint32_t p = sorted_prim[r & ~LEAF_FLAG];
      rlo              = box_lo[p];
      rhi              = box_hi[p];
    }
    else {
      rlo = __ldcg(&t.lo[r]);
      rhi = __ldcg(&t.hi[r]);
    }
    t.lo[node] = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
    t.hi[node] = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
    node       = t.parent[node];
  }
}


// ---------------------------------------------------------------------------------------------
// 4b. PLOC: parallel locally-ordered clustering (Meister & Bittner, "Parallel Locally-Ordered Clustering for
// Bounding Volume Hierarchy Construction", TVCG 2018). Clusters live in Morton order; every iteration each cluster
// finds the neighbour within +-radius whose union box has the smallest surface area, mutual pairs merge into a new
// binary node, and the survivors are compacted (order preserved). Equal areas resolve to the smaller index, which
// guarantees at least one mutual pair per iteration.
// ---------------------------------------------------------------------------------------------
#define PLOC_NONE 0xFFFFFFFFu

__global__ void k_ploc_init(uint32_t n, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ box_lo,
                            const float4* __restrict__ box_hi, uint32_t* __restrict__ ids, float4* __restrict__ clo, float4* __restrict__ chi) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t p = sorted_prim[i];
  ids[i]           = LEAF_FLAG | i;
  clo[i]           = box_lo[p];
  chi[i]           = box_hi[p];
}

__global__ void k_ploc_nn(uint32_t m, int radius, const float4* __restrict__ clo, const float4* __restrict__ chi, uint32_t* __restrict__ nn) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
    return;
  const float4 lo = clo[i], hi = chi[i];
  float best      = FLT_MAX;
  uint32_t bj     = PLOC_NONE;
  const int j0 = max((int) i - radius, 0), j1 = min((int) i + radius, (int) m - 1);
  for (int j = j0; j <= j1; j++) {
    if (j == (int) i)
      continue;
    const float4 l = clo[j], h = chi[j];
    const float dx = fmaxf(hi.x, h.x) - fminf(lo.x, l.x);
    const float dy = fmaxf(hi.y, h.y) - fminf(lo.y, l.y);
    const float dz = fmaxf(hi.z, h.z) - fminf(lo.z, l.z);
    const float a  = dx * dy + dy * dz + dz * dx;
    if (a < best) {  // ascending j: ties keep the smaller index
      best = a;
      bj   = (uint32_t) j;
    }
  }
  nn[i] = bj;
}

__global__ void k_ploc_merge(uint32_t m, uint32_t* __restrict__ ids, float4* __restrict__ clo, float4* __restrict__ chi,
                             const uint32_t* __restrict__ nn, Bvh2 t, uint32_t* __restrict__ node_counter, uint32_t* __restrict__ keep) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
    return;
  const uint32_t j = nn[i];
  if (j == PLOC_NONE || nn[j] != i) {
    keep[i] = 1;
    return;
  }
  if (i > j) {
    keep[i] = 0;
    return;
  }
  const uint32_t idx = atomicAdd(node_counter, 1u);
  const uint32_t l = ids[i], r = ids[j];
  const float4 llo = clo[i], lhi = chi[i], rlo = clo[j], rhi = chi[j];
  const float4 ulo = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
  const float4 uhi = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
  t.left[idx]  = l;
  t.right[idx] = r;
  t.lo[idx]    = ulo;
  t.hi[idx]    = uhi;
  t.count[idx] = ((l & LEAF_FLAG) ? 1u : t.count[l]) + ((r & LEAF_FLAG) ? 1u : t.count[r]);
  t.parent[idx] = PLOC_NONE;
  if (l & LEAF_FLAG)
    t.leaf_parent[l & ~LEAF_FLAG] = idx;
  else
    t.parent[l] = idx;
  if (r & LEAF_FLAG)
    t.leaf_parent[r & ~LEAF_FLAG] = idx;
  else
    t.parent[r] = idx;
  ids[i]  = idx;
  clo[i]  = ulo;
  chi[i]  = uhi;
  keep[i] = 1;
}

__global__ void k_ploc_compact(uint32_t m, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ offset, const uint32_t* __restrict__ ids,
                               const float4* __restrict__ clo, const float4* __restrict__ chi, uint32_t* __restrict__ ids_out,
                               float4* __restrict__ clo_out, float4* __restrict__ chi_out, uint32_t* __restrict__ m_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m)
    return;
  if (keep[i]) {
    const uint32_t o = offset[i];
    ids_out[o]       = ids[i];
    clo_out[o]       = clo[i];
    chi_out[o]       = chi[i];
  }
  if (i == m - 1)
    *m_out = offset[i] + keep[i];
}

// Linearisation: position of a leaf / first position of a node in the depth-first order of the finished tree
// (sum of the left-sibling subtree sizes on the way to the root). The collapse addresses leaf ranges through it.
__device__ __forceinline__ uint32_t ploc_position(const Bvh2& t, uint32_t ref, uint32_t parent) {
  uint32_t pos = 0;
  for (uint32_t guard = 0; parent != PLOC_NONE && guard < (1u << 20); guard++) {
    if (t.right[parent] == ref) {
      const uint32_t l = t.left[parent];
      pos += (l & LEAF_FLAG) ? 1u : t.count[l];
    }
    ref    = parent;
    parent = t.parent[parent];
  }
  return pos;
}

// Tree rotations (Kensler, "Tree Rotations for Improving Bounding Volume Hierarchies", 2008) on the PLOC hierarchy, bottom-up: the
// second thread to arrive at a node owns its whole subtree (every thread below has terminated or is this one), so it may swap
// one child with a grandchild under the other child when that shrinks the surface area of the node in between. Only the node in
// between changes its box and count; leaf order is re-derived afterwards (k_ploc_leaf_positions). LUMB200_BVH_ROTATIONS passes.
__device__ __forceinline__ void rot_box(const Bvh2& t, uint32_t ref, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ box_lo,
                                        const float4* __restrict__ box_hi, float4& lo, float4& hi) {
  if (ref & LEAF_FLAG) {
    const uint32_t p = sorted_prim[ref & ~LEAF_FLAG];
    lo               = box_lo[p];
    hi               = box_hi[p];
  }
  else {
    lo = __ldcg(&t.lo[ref]);
    hi = __ldcg(&t.hi[ref]);
  }
}
__device__ __forceinline__ float rot_area(float4 lo, float4 hi) {
  const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ float rot_union_area(float4 alo, float4 ahi, float4 blo, float4 bhi) {
  return rot_area(make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f),
                  make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f));
}
__device__ __forceinline__ void rot_set_parent(const Bvh2& t, uint32_t ref, uint32_t parent) {
  if (ref & LEAF_FLAG)
    t.leaf_parent[ref & ~LEAF_FLAG] = parent;
  else
    t.parent[ref] = parent;
}

__global__ void k_bvh2_rotate(uint32_t n, Bvh2 t, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ box_lo,
                              const float4* __restrict__ box_hi, uint32_t* __restrict__ num_rotations) {
  const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n)
    return;
  uint32_t node = __ldcg(&t.leaf_parent[leaf]);
  while (node != 0xFFFFFFFFu) {
    __threadfence();
    if (atomicAdd(&t.flags[node], 1u) == 0)
      return;
    const uint32_t ch[2] = {__ldcg(&t.left[node]), __ldcg(&t.right[node])};
    float4 clo[2], chi[2];
    rot_box(t, ch[0], sorted_prim, box_lo, box_hi, clo[0], chi[0]);
    rot_box(t, ch[1], sorted_prim, box_lo, box_hi, clo[1], chi[1]);
    float best = 0.0f;
    int best_side = -1, best_grand = 0;
    uint32_t g[2][2] = {{0, 0}, {0, 0}};
    float4 glo[2][2], ghi[2][2];
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const uint32_t c = ch[side];
      if (c & LEAF_FLAG)
        continue;
      g[side][0] = __ldcg(&t.left[c]), g[side][1] = __ldcg(&t.right[c]);
      rot_box(t, g[side][0], sorted_prim, box_lo, box_hi, glo[side][0], ghi[side][0]);
      rot_box(t, g[side][1], sorted_prim, box_lo, box_hi, glo[side][1], ghi[side][1]);
      const float area_c = rot_area(clo[side], chi[side]);
      const int o        = side ^ 1;
#pragma unroll
      for (int k = 0; k < 2; k++) {  // the other child takes the place of grandchild k; grandchild k ^ 1 stays
        const float delta = rot_union_area(clo[o], chi[o], glo[side][k ^ 1], ghi[side][k ^ 1]) - area_c;
        if (delta < best - 1e-6f * area_c) {
          best = delta, best_side = side, best_grand = k;
        }
      }
    }
    if (best_side >= 0) {
      const int side = best_side, o = side ^ 1, k = best_grand;
      const uint32_t c = ch[side], other = ch[o], up = g[side][k], stay = g[side][k ^ 1];
      // c := {other, stay} (other replaces `up` in slot k), node := {c, up} (up replaces `other`)
      if (k == 0)
        t.left[c] = other;
      else
        t.right[c] = other;
      if (o == 0)
        t.left[node] = up;
      else
        t.right[node] = up;
      rot_set_parent(t, other, c);
      rot_set_parent(t, up, node);
      t.lo[c] = make_float4(fminf(clo[o].x, glo[side][k ^ 1].x), fminf(clo[o].y, glo[side][k ^ 1].y), fminf(clo[o].z, glo[side][k ^ 1].z), 0.0f);
      t.hi[c] = make_float4(fmaxf(chi[o].x, ghi[side][k ^ 1].x), fmaxf(chi[o].y, ghi[side][k ^ 1].y), fmaxf(chi[o].z, ghi[side][k ^ 1].z), 0.0f);
      t.count[c] = ((other & LEAF_FLAG) ? 1u : __ldcg(&t.count[other])) + ((stay & LEAF_FLAG) ? 1u : __ldcg(&t.count[stay]));
      atomicAdd(num_rotations, 1u);
    }
    node = __ldcg(&t.parent[node]);
  }
}

// ---------------------------------------------------------------------------------------------
// 4c. Parallel reinsertion (Meister & Bittner, "Parallel Reinsertion for Bounding Volume Hierarchy Optimization", EG 2018) on the
// PLOC hierarchy - the global restructuring that tree rotations cannot do. Every node `in` (leaf or inner, not the root or its
// children) looks for the position in the tree where removing it together with its parent P and re-attaching both as the new
// parent / sibling of a target node T shrinks the summed surface area of the inner nodes the most:
//   find  : walk the pivot up from P; at level k the subtree X_k that hangs off the path (X_0 = S = sibling of in, X_k = sibling of
//           A_(k-1), A_0 = P, A_k = parent(A_(k-1))) is searched by branch and bound. bound(X_k) = area(P) + sum over 0 < j < k of
//           area(A_j) - area(R_j) with R_j = box(X_0 u .. u X_j), the box A_j shrinks to once `in` is gone; descending into a node
//           x costs area(x u B) - area(x), ending at x costs area(x u B)
//   decide: a move rewires in, P, S, G = parent(P), T, Q = parent(T); the moves applied in one pass must own these six nodes
//           exclusively (64-bit atomicMax of (gain, id) into lock_topo; the owners are the "live" moves). That alone keeps every
//           pointer update consistent, but two moves can still tie a cycle (each carrying the other's target inside its moved
//           subtree), so a live move also yields when a node on its path in -> A_k -> T is itself the `in` of a live move with a
//           higher key, and when the path of a live move with a higher key runs through its own `in` (lock_path). With no moved
//           node on the path, P and T lie in the same un-moved piece of the tree, i.e. every moved subtree re-attaches to the piece
//           it hung in before and the nesting of the pieces stays the acyclic one of the input tree. Paths may otherwise cross:
//           the gains of crossing moves are estimates. The host checks after every pass that the root still counts n primitives.
//   apply : pointer updates of the surviving moves, then boxes and primitive counts of all inner nodes bottom-up (k_refit_count)
// LUMB200_BVH_REINSERT passes (0 = off).
// ---------------------------------------------------------------------------------------------
struct ReinsertPlan {
  uint32_t target;  // T (PLOC_NONE: no move)
  uint32_t top;     // X_k, the root of the searched subtree T was found in
  uint32_t level;   // k; bit 30: the move owns its six nodes, bit 31: it is applied in this pass
  float gain;
};
#define RIN_APPLY 0x80000000u
#define RIN_ALIVE 0x40000000u
#define RIN_LEVEL_MASK 0x3FFFFFFFu

__device__ __forceinline__ uint32_t rin_parent(const Bvh2& t, uint32_t ref) {
  return (ref & LEAF_FLAG) ? t.leaf_parent[ref & ~LEAF_FLAG] : t.parent[ref];
}
__device__ __forceinline__ uint32_t rin_sibling(const Bvh2& t, uint32_t ref, uint32_t parent) {
  const uint32_t l = t.left[parent];
  return (l == ref) ? t.right[parent] : l;
}
__device__ __forceinline__ uint32_t rin_id(uint32_t ref, uint32_t ni) {  // candidate / lock index of a node reference
  return (ref & LEAF_FLAG) ? ni + (ref & ~LEAF_FLAG) : ref;
}
__device__ __forceinline__ unsigned long long rin_key(float gain, uint32_t c) {
  return ((unsigned long long) __float_as_uint(gain) << 32) | (unsigned long long) c;
}
__device__ __forceinline__ void rin_box_const(const Bvh2& t, uint32_t ref, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ box_lo,
                                              const float4* __restrict__ box_hi, float4& lo, float4& hi) {
  if (ref & LEAF_FLAG) {
    const uint32_t p = sorted_prim[ref & ~LEAF_FLAG];
    lo               = box_lo[p];
    hi               = box_hi[p];
  }
  else {
    lo = t.lo[ref];
    hi = t.hi[ref];
  }
}

#define RIN_STACK 48

// candidate c < ni: inner node c; otherwise leaf c - ni
__global__ void k_reinsert_find(uint32_t n, uint32_t ni, Bvh2 t, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ box_lo,
                                const float4* __restrict__ box_hi, ReinsertPlan* __restrict__ plan, unsigned long long* __restrict__ lock_topo) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ni + n)
    return;
  const uint32_t in = (c < ni) ? c : (LEAF_FLAG | (c - ni));
  ReinsertPlan best;
  best.target = PLOC_NONE, best.top = PLOC_NONE, best.level = 0, best.gain = 0.0f;
  const uint32_t P = rin_parent(t, in);
  if (P == PLOC_NONE || t.parent[P] == PLOC_NONE) {
    plan[c] = best;
    return;
  }
  float4 blo, bhi;
  rin_box_const(t, in, sorted_prim, box_lo, box_hi, blo, bhi);
  const float area_p = rot_area(t.lo[P], t.hi[P]);
  const float eps    = 1e-5f * area_p;
  const uint32_t S   = rin_sibling(t, in, P);

  uint32_t st_node[RIN_STACK];
  float st_bound[RIN_STACK];

  uint32_t pivot = P, X = S, k = 0;
  float d = area_p;
  float4 rlo, rhi;  // R_k
  rin_box_const(t, S, sorted_prim, box_lo, box_hi, rlo, rhi);
  for (;;) {
    st_node[0]  = X;
    st_bound[0] = d;
    int sp      = 1;
    while (sp > 0) {
      sp--;
      const uint32_t x = st_node[sp];
      const float db   = st_bound[sp];
      if (db <= best.gain + eps)
        continue;
      float4 xlo, xhi;
      rin_box_const(t, x, sorted_prim, box_lo, box_hi, xlo, xhi);
      const float direct = rot_union_area(xlo, xhi, blo, bhi);
      const float g      = db - direct;
      if (g > best.gain + eps && x != S) {
        best.gain = g, best.target = x, best.top = X, best.level = k;
      }
      if (!(x & LEAF_FLAG)) {
        const float dc = g + rot_area(xlo, xhi);
        if (dc > best.gain + eps && sp + 2 <= RIN_STACK) {
          st_node[sp] = t.left[x], st_bound[sp] = dc, sp++;
          st_node[sp] = t.right[x], st_bound[sp] = dc, sp++;
        }
      }
    }
    // one level up
    const uint32_t up = t.parent[pivot];
    if (up == PLOC_NONE)
      break;
    if (k > 0) {
      float4 xlo, xhi;
      rin_box_const(t, X, sorted_prim, box_lo, box_hi, xlo, xhi);
      rlo = make_float4(fminf(rlo.x, xlo.x), fminf(rlo.y, xlo.y), fminf(rlo.z, xlo.z), 0.0f);
      rhi = make_float4(fmaxf(rhi.x, xhi.x), fmaxf(rhi.y, xhi.y), fmaxf(rhi.z, xhi.z), 0.0f);
      d += rot_area(t.lo[pivot], t.hi[pivot]) - rot_area(rlo, rhi);
    }
    X     = rin_sibling(t, pivot, up);
    pivot = up;
    k++;
  }
  plan[c] = best;
  if (best.target == PLOC_NONE)
    return;
  const unsigned long long key = rin_key(best.gain, c);
  const uint32_t G             = t.parent[P];
  const uint32_t Q             = rin_parent(t, best.target);
  atomicMax(&lock_topo[c], key);
  atomicMax(&lock_topo[rin_id(S, ni)], key);
  atomicMax(&lock_topo[P], key);
  atomicMax(&lock_topo[G], key);
  atomicMax(&lock_topo[rin_id(best.target, ni)], key);
  atomicMax(&lock_topo[Q], key);
}

// read-only on the topology, two steps: (1) the moves that own their six nodes stay alive and mark their path (without `in`) in
// lock_path; (2) of those, the ones with no stronger live move starting on their path or passing through their `in` are applied
__global__ void k_reinsert_decide_topo(uint32_t n, uint32_t ni, Bvh2 t, ReinsertPlan* __restrict__ plan, const unsigned long long* __restrict__ lock_topo,
                                       unsigned long long* __restrict__ lock_path) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ni + n)
    return;
  const ReinsertPlan pl = plan[c];
  if (pl.target == PLOC_NONE)
    return;
  const uint32_t in            = (c < ni) ? c : (LEAF_FLAG | (c - ni));
  const unsigned long long key = rin_key(pl.gain, c);
  const uint32_t P = rin_parent(t, in);
  const uint32_t S = rin_sibling(t, in, P);
  const uint32_t G = t.parent[P];
  const uint32_t Q = rin_parent(t, pl.target);
  const bool alive = lock_topo[c] == key && lock_topo[rin_id(S, ni)] == key && lock_topo[P] == key && lock_topo[G] == key &&
                     lock_topo[rin_id(pl.target, ni)] == key && lock_topo[Q] == key;
  if (!alive)
    return;
  plan[c].level = pl.level | RIN_ALIVE;
  uint32_t a    = P;
  for (uint32_t j = 0; j <= pl.level && a != PLOC_NONE; j++) {
    atomicMax(&lock_path[a], key);
    a = t.parent[a];
  }
  uint32_t x = pl.target;
  for (;;) {
    atomicMax(&lock_path[rin_id(x, ni)], key);
    if (x == pl.top)
      break;
    x = rin_parent(t, x);
  }
}

__device__ __forceinline__ bool rin_stronger_live_move(const ReinsertPlan* __restrict__ plan, uint32_t y, unsigned long long key) {
  // RIN_APPLY of y may be set concurrently by its own thread; only RIN_ALIVE (final since the previous kernel) is read
  return (__ldcg(&plan[y].level) & RIN_ALIVE) && rin_key(plan[y].gain, y) > key;
}

__global__ void k_reinsert_decide_path(uint32_t n, uint32_t ni, Bvh2 t, ReinsertPlan* __restrict__ plan, const unsigned long long* __restrict__ lock_path) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ni + n)
    return;
  const uint32_t target = plan[c].target;
  const uint32_t lvl    = __ldcg(&plan[c].level);
  if (target == PLOC_NONE || !(lvl & RIN_ALIVE))
    return;
  const uint32_t top = plan[c].top, level = lvl & RIN_LEVEL_MASK;
  const uint32_t in  = (c < ni) ? c : (LEAF_FLAG | (c - ni));
  const unsigned long long key = rin_key(plan[c].gain, c);
  bool ok    = lock_path[c] < key;  // no stronger live move passes through `in`
  uint32_t a = rin_parent(t, in);
  for (uint32_t j = 0; ok && j <= level && a != PLOC_NONE; j++) {
    ok = !rin_stronger_live_move(plan, a, key);
    a  = t.parent[a];
  }
  uint32_t x = target;
  while (ok) {
    ok = !rin_stronger_live_move(plan, rin_id(x, ni), key);
    if (x == top)
      break;
    x = rin_parent(t, x);
  }
  if (ok)
    plan[c].level = lvl | RIN_APPLY;
}

__device__ __forceinline__ void rin_replace_child(const Bvh2& t, uint32_t parent, uint32_t from, uint32_t to) {
  if (t.left[parent] == from)
    t.left[parent] = to;
  else
    t.right[parent] = to;
}

// every pointer written here belongs to one of the six nodes the move owns
__global__ void k_reinsert_apply(uint32_t n, uint32_t ni, Bvh2 t, const ReinsertPlan* __restrict__ plan, uint32_t* __restrict__ num_moves) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ni + n)
    return;
  const ReinsertPlan pl = plan[c];
  if (pl.target == PLOC_NONE || !(pl.level & RIN_APPLY))
    return;
  const uint32_t in = (c < ni) ? c : (LEAF_FLAG | (c - ni));
  const uint32_t P  = rin_parent(t, in);
  const uint32_t S  = rin_sibling(t, in, P);
  const uint32_t T  = pl.target;
  const uint32_t G  = t.parent[P];
  const uint32_t Q  = rin_parent(t, T);
  // S takes the place of P
  rin_replace_child(t, G, P, S);
  rot_set_parent(t, S, G);
  // P goes between Q and T, keeping `in` as its other child (Q == G and Q == S are covered by the order of the updates)
  rin_replace_child(t, Q, T, P);
  t.parent[P] = Q;
  rin_replace_child(t, P, S, T);
  rot_set_parent(t, T, P);
  atomicAdd(num_moves, 1u);
}

// boxes and primitive counts of every inner node, bottom-up (t.flags zeroed by the caller)
__global__ void k_refit_count(uint32_t n, Bvh2 t, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ box_lo,
                              const float4* __restrict__ box_hi) {
  const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n)
    return;
  uint32_t node = __ldcg(&t.leaf_parent[leaf]);
  while (node != 0xFFFFFFFFu) {
    __threadfence();
    if (atomicAdd(&t.flags[node], 1u) == 0)
      return;
    const uint32_t l = __ldcg(&t.left[node]), r = __ldcg(&t.right[node]);
    float4 llo, lhi, rlo, rhi;
    rot_box(t, l, sorted_prim, box_lo, box_hi, llo, lhi);
    rot_box(t, r, sorted_prim, box_lo, box_hi, rlo, rhi);
    t.lo[node]    = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
    t.hi[node]    = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
    t.count[node] = ((l & LEAF_FLAG) ? 1u : __ldcg(&t.count[l])) + ((r & LEAF_FLAG) ? 1u : __ldcg(&t.count[r]));
    node          = __ldcg(&t.parent[node]);
  }
}

// summed surface area of the inner nodes (the quantity reinsertion minimises), for LUMB200_BVH_VERBOSE
__global__ void k_bvh2_area_sum(uint32_t ni, Bvh2 t, double* __restrict__ sum) {
  const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
  double a            = 0.0;
  if (node < ni)
    a = (double) rot_area(t.lo[node], t.hi[node]);
  for (int o = 16; o > 0; o >>= 1)
    a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
  if ((threadIdx.x & 31) == 0)
    atomicAdd(sum, a);
}

__global__ void k_ploc_leaf_positions(uint32_t n, Bvh2 t, const uint32_t* __restrict__ sorted_prim, uint32_t* __restrict__ newpos,
                                      uint32_t* __restrict__ sorted_prim_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t pos   = ploc_position(t, LEAF_FLAG | i, t.leaf_parent[i]);
  newpos[i]            = pos;
  sorted_prim_out[pos] = sorted_prim[i];
}

__global__ void k_ploc_node_first(uint32_t ni, Bvh2 t) {
  const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= ni)
    return;
  t.first[node] = ploc_position(t, node, t.parent[node]);
}

__global__ void k_ploc_fix_refs(uint32_t ni, Bvh2 t, const uint32_t* __restrict__ newpos) {
  const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= ni)
    return;
  const uint32_t l = t.left[node], r = t.right[node];
  if (l & LEAF_FLAG)
    t.left[node] = LEAF_FLAG | newpos[l & ~LEAF_FLAG];
  if (r & LEAF_FLAG)
    t.right[node] = LEAF_FLAG | newpos[r & ~LEAF_FLAG];
  t.last[node] = t.first[node] + t.count[node] - 1u;
}

// ---------------------------------------------------------------------------------------------
// 5. collapse to the compressed 8-wide layout
// ---------------------------------------------------------------------------------------------
struct WorkItem {
  uint32_t bvh2;  // child ref (may carry LEAF_FLAG only for the degenerate single-primitive scene)
  uint32_t bvh8;
};

struct ChildBox {
  float lo[3], hi[3];
};

__device__ __forceinline__ uint32_t ref_count(const Bvh2& t, uint32_t ref) {
  return (ref & LEAF_FLAG) ? 1u : t.count[ref];
}
__device__ __forceinline__ uint32_t ref_first(const Bvh2& t, uint32_t ref) { return (ref & LEAF_FLAG) ? (ref & ~LEAF_FLAG) : t.first[ref]; }

__device__ __forceinline__ void ref_box(const Bvh2& t, uint32_t ref, const uint32_t* sorted_prim, const float4* box_lo, const float4* box_hi,
                                        ChildBox& b) {
  float4 lo, hi;
  if (ref & LEAF_FLAG) {
    const uint32_t p = sorted_prim[ref & ~LEAF_FLAG];
    lo               = box_lo[p];
    hi               = box_hi[p];
  }
  else {
    lo = t.lo[ref];
    hi = t.hi[ref];
  }
  b.lo[0] = lo.x, b.lo[1] = lo.y, b.lo[2] = lo.z;
  b.hi[0] = hi.x, b.hi[1] = hi.y, b.hi[2] = hi.z;
}

__device__ __forceinline__ float box_half_area(const ChildBox& b) {
  const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  return dx * dy + dy * dz + dz * dx;
}

// SAH-optimal collapse (Ylitie, Karras, Laine 2017, section 3.1): for every binary node n and every i in 1..7,
// cost[n][i-1] = cheapest representation of the subtree of n as at most i children slots of one 8-wide node:
//   C(n, 1) = min( leaf: A(n) * count * c_prim  (count <= LEAF_MAX_TRIS),  internal: D(n, 8) + A(n) * c_node )
//   C(n, i) = min( D(n, i), C(n, i - 1) ),   D(n, j) = min over 0 < k < j of C(left, k) + C(right, j - k)
// evaluated bottom-up (the second thread to arrive at a node owns it, as in k_refit); choice[] records the argmin so
// that k_collapse can replay the decisions top-down. A(.) is the half surface area; it is not normalised (a common
// factor does not change any argmin). A single primitive costs A(prim) * c_prim for every i.
struct CollapseDp {
  float* cost;      // [ni * 7]
  uint8_t* choice;  // [ni * 8]: [0] 0 = leaf / 1 = internal; [i-1], i = 2..7: 0 = same as C(n, i-1), else k = slots of the left child; [7] = k of D(n, 8)
  float c_node, c_prim;
};

__device__ __forceinline__ float box_half_area4(const float4 lo, const float4 hi) {
  const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}

__global__ void k_collapse_dp(int n, Bvh2 t, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ box_lo,
                              const float4* __restrict__ box_hi, CollapseDp dp) {
  const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
  if (leaf >= n)
    return;
  uint32_t node = t.leaf_parent[leaf];
  while (node != 0xFFFFFFFFu) {
    __threadfence();
    if (atomicAdd(&t.flags[node], 1u) == 0)
      return;
    float cl[7], cr[7];
    const uint32_t l = t.left[node], r = t.right[node];
    if (l & LEAF_FLAG) {
      const uint32_t p = sorted_prim[l & ~LEAF_FLAG];
      const float c    = box_half_area4(box_lo[p], box_hi[p]) * dp.c_prim;
      for (int i = 0; i < 7; i++)
        cl[i] = c;
    }
    else {
      for (int i = 0; i < 7; i++)
        cl[i] = __ldcg(&dp.cost[7 * (size_t) l + i]);
    }
    if (r & LEAF_FLAG) {
      const uint32_t p = sorted_prim[r & ~LEAF_FLAG];
      const float c    = box_half_area4(box_lo[p], box_hi[p]) * dp.c_prim;
      for (int i = 0; i < 7; i++)
        cr[i] = c;
    }
    else {
      for (int i = 0; i < 7; i++)
        cr[i] = __ldcg(&dp.cost[7 * (size_t) r + i]);
    }
    const float area     = box_half_area4(t.lo[node], t.hi[node]);
    const uint32_t count = t.count[node];
    float dist[9];
    uint8_t dk[9];
    for (int j = 2; j <= 8; j++) {
      float best = FLT_MAX;
      int bk     = 1;
      for (int k = 1; k < j; k++) {
        const float c = cl[k - 1] + cr[j - k - 1];
        if (c < best) {
          best = c;
          bk   = k;
        }
      }
      dist[j] = best;
      dk[j]   = (uint8_t) bk;
    }
    uint8_t* ch            = dp.choice + 8 * (size_t) node;
    float* co              = dp.cost + 7 * (size_t) node;
    const float c_internal = dist[8] + area * dp.c_node;
    const float c_leaf     = (count <= LEAF_MAX_TRIS) ? area * (float) count * dp.c_prim : FLT_MAX;
    float prev             = fminf(c_leaf, c_internal);
    ch[0]                  = (c_leaf <= c_internal) ? 0 : 1;
    ch[7]                  = dk[8];
    co[0]                  = prev;
    for (int i = 2; i <= 7; i++) {
      if (dist[i] < prev) {
        prev      = dist[i];
        ch[i - 1] = dk[i];
      }
      else
        ch[i - 1] = 0;
      co[i - 1] = prev;
    }
    node = t.parent[node];
  }
}

__global__ void k_collapse(const WorkItem* __restrict__ in, uint32_t n_in, WorkItem* __restrict__ out, uint32_t* __restrict__ counters,
                           Bvh2 t, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ box_lo,
                           const float4* __restrict__ box_hi, const float4* __restrict__ world, Bvh8Node* __restrict__ nodes,
                           float4* __restrict__ tris_out, const uint8_t* __restrict__ dp_choice) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_in)
    return;
  const WorkItem item = in[w];

  uint32_t c[8];
  ChildBox cb[8];
  bool is_leaf[8];
  int nc = 0;

  if (item.bvh2 & LEAF_FLAG) {
    c[0] = item.bvh2;  // single primitive scene
    nc   = 1;
  }
  else if (ref_count(t, item.bvh2) <= LEAF_MAX_TRIS && (!dp_choice || item.bvh8 == 0)) {
    c[0] = item.bvh2;  // whole scene fits one leaf slot
    nc   = 1;
  }
  else if (dp_choice) {
    // replay the decisions of k_collapse_dp: D(item, 8), then every (subtree, budget) pair until it resolves to one slot
    uint32_t st_ref[8];
    uint32_t st_budget[8];
    int sp = 0;
    {
      const uint32_t k = dp_choice[8 * (size_t) item.bvh2 + 7];
      st_ref[sp] = t.right[item.bvh2], st_budget[sp++] = 8u - k;
      st_ref[sp] = t.left[item.bvh2], st_budget[sp++] = k;
    }
    while (sp > 0) {
      const uint32_t m = st_ref[--sp];
      uint32_t j       = st_budget[sp];
      if (m & LEAF_FLAG) {
        c[nc]         = m;
        is_leaf[nc++] = true;
        continue;
      }
      const uint8_t* ch = dp_choice + 8 * (size_t) m;
      while (j > 1 && ch[j - 1] == 0)
        j--;
      if (j == 1) {
        c[nc]         = m;
        is_leaf[nc++] = ch[0] == 0;
      }
      else {
        const uint32_t k = ch[j - 1];
        st_ref[sp] = t.right[m], st_budget[sp++] = j - k;
        st_ref[sp] = t.left[m], st_budget[sp++] = k;
      }
    }
    for (int k = 0; k < nc; k++)
      ref_box(t, c[k], sorted_prim, box_lo, box_hi, cb[k]);
  }
  else {
    c[0] = t.left[item.bvh2];
    c[1] = t.right[item.bvh2];
    nc   = 2;
  }
  if (!dp_choice || nc == 1) {
    for (int k = 0; k < nc; k++)
      ref_box(t, c[k], sorted_prim, box_lo, box_hi, cb[k]);

    // greedy: open the child with the largest surface area until 8 children
    while (nc < 8 && nc > 1) {
      int best        = -1;
      float best_area = -1.0f;
      for (int k = 0; k < nc; k++) {
        if ((c[k] & LEAF_FLAG) || ref_count(t, c[k]) <= LEAF_MAX_TRIS)
          continue;
        const float a = box_half_area(cb[k]);
        if (a > best_area) {
          best_area = a;
          best      = k;
        }
      }
      if (best < 0)
        break;
      const uint32_t ref = c[best];
      c[best]            = t.left[ref];
      c[nc]              = t.right[ref];
      ref_box(t, c[best], sorted_prim, box_lo, box_hi, cb[best]);
      ref_box(t, c[nc], sorted_prim, box_lo, box_hi, cb[nc]);
      nc++;
    }
    for (int k = 0; k < nc; k++)
      is_leaf[k] = (c[k] & LEAF_FLAG) || ref_count(t, c[k]) <= LEAF_MAX_TRIS;
  }

  // node box
  float nlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, nhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int k = 0; k < nc; k++)
    for (int a = 0; a < 3; a++) {
      nlo[a] = fminf(nlo[a], cb[k].lo[a]);
      nhi[a] = fmaxf(nhi[a], cb[k].hi[a]);
    }

  // slot assignment: slot bit a set <=> child sits on the high side of axis a (greedy on centroid offsets)
  int slot_of[8];
  {
    const float ncx = 0.5f * (nlo[0] + nhi[0]), ncy = 0.5f * (nlo[1] + nhi[1]), ncz = 0.5f * (nlo[2] + nhi[2]);
    float off[8][3];
    for (int k = 0; k < nc; k++) {
      off[k][0] = 0.5f * (cb[k].lo[0] + cb[k].hi[0]) - ncx;
      off[k][1] = 0.5f * (cb[k].lo[1] + cb[k].hi[1]) - ncy;
      off[k][2] = 0.5f * (cb[k].lo[2] + cb[k].hi[2]) - ncz;
    }
    uint32_t child_done = 0, slot_done = 0;
    for (int it = 0; it < nc; it++) {
      float best = -FLT_MAX;
      int bk = 0, bs = 0;
      for (int k = 0; k < nc; k++) {
        if (child_done & (1u << k))
          continue;
        for (int s = 0; s < 8; s++) {
          if (slot_done & (1u << s))
            continue;
          const float score = ((s & 1) ? off[k][0] : -off[k][0]) + ((s & 2) ? off[k][1] : -off[k][1]) + ((s & 4) ? off[k][2] : -off[k][2]);
          if (score > best) {
            best = score;
            bk   = k;
            bs   = s;
          }
        }
      }
      slot_of[bk] = bs;
      child_done |= 1u << bk;
      slot_done |= 1u << bs;
    }
  }

  // quantisation frame: 253 cells cover the extent, one spare cell on both sides for the outward rounding + margin of the children
  float p[3], scale[3];
  uint32_t ebits[3];
  for (int a = 0; a < 3; a++) {
    const float ext = nhi[a] - nlo[a];
    float cell      = fmaxf(ext / 253.0f, 1e-30f);
    uint32_t bits   = __float_as_uint(cell);
    uint32_t e      = (bits >> 23) + ((bits & 0x7FFFFFu) ? 1u : 0u);
    e               = min(max(e, 1u), 254u);
    for (;;) {
      scale[a] = __uint_as_float(e << 23);
      p[a]     = nlo[a] - scale[a];
      // verify the high side fits (fp rounding of p can cost a fraction of a cell)
      if ((nhi[a] - p[a]) / scale[a] <= 254.0f || e >= 254u)
        break;
      e++;
    }
    ebits[a] = e;
  }

  Bvh8Node node;
  node.px = p[0], node.py = p[1], node.pz = p[2];
  node.ex = (uint8_t) ebits[0], node.ey = (uint8_t) ebits[1], node.ez = (uint8_t) ebits[2];
  node.imask = 0;
  for (int s = 0; s < 8; s++) {
    node.meta[s] = 0;
    node.qlox[s] = node.qloy[s] = node.qloz[s] = 255;
    node.qhix[s] = node.qhiy[s] = node.qhiz[s] = 0;
  }

  // classify children, count
  uint32_t inner_slots = 0;
  uint32_t total_tris  = 0;
  for (int k = 0; k < nc; k++) {
    const bool leaf = is_leaf[k];
    if (!leaf)
      inner_slots |= 1u << slot_of[k];
    else
      total_tris += ref_count(t, c[k]);
  }
  const uint32_t num_inner  = __popc(inner_slots);
  const uint32_t child_base = num_inner ? atomicAdd(&counters[0], num_inner) : 0u;
  const uint32_t tri_base   = total_tris ? atomicAdd(&counters[1], total_tris) : 0u;
  const uint32_t out_base   = num_inner ? atomicAdd(&counters[2], num_inner) : 0u;

  node.imask      = (uint8_t) inner_slots;
  node.child_base = child_base;
  node.tri_base   = tri_base;

  // leaf triangles are laid out in slot order so offsets are deterministic given the slot assignment
  uint32_t tri_cursor = 0;
  for (int s = 0; s < 8; s++) {
    int k = -1;
    for (int kk = 0; kk < nc; kk++)
      if (slot_of[kk] == s)
        k = kk;
    if (k < 0)
      continue;

    // quantise outwards with a margin of at least LB_QUANT_MARGIN cells: the traversal's folded plane arithmetic (traverse.cuh:
    // lb_node_hits) rounds t by at most 2^-9 of a cell, so 1 / 64 of a cell keeps the slab test conservative; the full extra cell
    // used before cost about 1.5 cells of inflation per side instead of 0.5
    uint8_t qlo[3], qhi[3];
    for (int a = 0; a < 3; a++) {
      float fl = floorf((cb[k].lo[a] - p[a]) / scale[a] - LB_QUANT_MARGIN);
      float fh = ceilf((cb[k].hi[a] - p[a]) / scale[a] + LB_QUANT_MARGIN);
      fl       = fminf(fmaxf(fl, 0.0f), 255.0f);
      fh       = fminf(fmaxf(fh, 0.0f), 255.0f);
      qlo[a]   = (uint8_t) fl;
      qhi[a]   = (uint8_t) fh;
    }
    node.qlox[s] = qlo[0], node.qloy[s] = qlo[1], node.qloz[s] = qlo[2];
    node.qhix[s] = qhi[0], node.qhiy[s] = qhi[1], node.qhiz[s] = qhi[2];

    if (inner_slots & (1u << s)) {
      node.meta[s]        = (uint8_t) (0x20u | (24u + s));
      const uint32_t rank = __popc(inner_slots & ((1u << s) - 1u));
      WorkItem o;
      o.bvh2             = c[k];
      o.bvh8             = child_base + rank;
      out[out_base + rank] = o;
    }
    else {
      const uint32_t cnt   = ref_count(t, c[k]);
      const uint32_t first = ref_first(t, c[k]);
      const uint32_t unary = (cnt == 1) ? 1u : (cnt == 2) ? 3u : 7u;
      node.meta[s]         = (uint8_t) ((unary << 5) | tri_cursor);
      for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t prim = sorted_prim[first + j];
        const size_t dst    = 3 * (size_t) (tri_base + tri_cursor + j);
        tris_out[dst + 0]   = world[3 * (size_t) prim + 0];
        tris_out[dst + 1]   = world[3 * (size_t) prim + 1];
        tris_out[dst + 2]   = world[3 * (size_t) prim + 2];
      }
      tri_cursor += cnt;
    }
  }

  nodes[item.bvh8] = node;
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
void lb_bvh8_free(LbBvhBuffers* b) {
  if (b->nodes)
    cudaFree(b->nodes);
  if (b->tris)
    cudaFree(b->tris);
  *b = LbBvhBuffers();
}

#define LB_FREE_ALL()            \
  do {                           \
    for (void* ptr : scratch)    \
      if (ptr)                   \
        cudaFree(ptr);           \
  } while (0)

Lumb200Result lb_bvh8_build(const float4* world_tris, uint32_t n, LbBvhBuffers* out, cudaStream_t stream, float* build_ms) {
  lb_bvh8_free(out);

  cudaEvent_t ev0, ev1;
  LB_CHECK(cudaEventCreate(&ev0));
  LB_CHECK(cudaEventCreate(&ev1));
  LB_CHECK(cudaEventRecord(ev0, stream));

  if (n == 0) {
    // a single empty node: every ray misses
    Bvh8Node empty;
    memset(&empty, 0, sizeof(empty));
    for (int s = 0; s < 8; s++) {
      empty.qlox[s] = empty.qloy[s] = empty.qloz[s] = 255;
    }
    empty.ex = empty.ey = empty.ez = 127;
    LB_CHECK(cudaMalloc(&out->nodes, sizeof(Bvh8Node)));
    LB_CHECK(cudaMalloc(&out->tris, sizeof(float4) * 3));
    LB_CHECK(cudaMemcpyAsync(out->nodes, &empty, sizeof(empty), cudaMemcpyHostToDevice, stream));
    LB_CHECK(cudaMemsetAsync(out->tris, 0, sizeof(float4) * 3, stream));
    LB_CHECK(cudaStreamSynchronize(stream));
    out->num_nodes = 1;
    out->num_tris  = 0;
    out->depth     = 1;
    out->bytes     = sizeof(Bvh8Node) + sizeof(float4) * 3;
    if (build_ms)
      *build_ms = 0.0f;
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    return LUMB200_SUCCESS;
  }

  std::vector<void*> scratch;
  auto alloc = [&](size_t bytes) -> void* {
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, bytes ? bytes : 16) != cudaSuccess)
      return nullptr;
    scratch.push_back(ptr);
    return ptr;
  };

  const uint32_t blocks = (n + BUILD_THREADS - 1) / BUILD_THREADS;
  const uint32_t ni     = (n > 1) ? n - 1 : 1;

  float4* box_lo      = (float4*) alloc(sizeof(float4) * n);
  float4* box_hi      = (float4*) alloc(sizeof(float4) * n);
  BuildBounds* bounds = (BuildBounds*) alloc(sizeof(BuildBounds));
  uint64_t* keys      = (uint64_t*) alloc(sizeof(uint64_t) * n);
  uint64_t* keys_s    = (uint64_t*) alloc(sizeof(uint64_t) * n);
  uint32_t* vals      = (uint32_t*) alloc(sizeof(uint32_t) * n);
  uint32_t* vals_s    = (uint32_t*) alloc(sizeof(uint32_t) * n);
  Bvh2 t;
  t.left        = (uint32_t*) alloc(sizeof(uint32_t) * ni);
  t.right       = (uint32_t*) alloc(sizeof(uint32_t) * ni);
  t.first       = (uint32_t*) alloc(sizeof(uint32_t) * ni);
  t.last        = (uint32_t*) alloc(sizeof(uint32_t) * ni);
  t.parent      = (uint32_t*) alloc(sizeof(uint32_t) * ni);
  t.leaf_parent = (uint32_t*) alloc(sizeof(uint32_t) * n);
  t.lo          = (float4*) alloc(sizeof(float4) * ni);
  t.hi          = (float4*) alloc(sizeof(float4) * ni);
  t.flags       = (uint32_t*) alloc(sizeof(uint32_t) * ni);
  t.count       = (uint32_t*) alloc(sizeof(uint32_t) * ni);
  WorkItem* q0       = (WorkItem*) alloc(sizeof(WorkItem) * n);
  WorkItem* q1       = (WorkItem*) alloc(sizeof(WorkItem) * n);
  uint32_t* counters = (uint32_t*) alloc(sizeof(uint32_t) * 4);
  Bvh8Node* nodes_tmp = (Bvh8Node*) alloc(sizeof(Bvh8Node) * (size_t) (n + 1));
  float4* tris_out    = nullptr;
  if (cudaMalloc(&tris_out, sizeof(float4) * 3 * (size_t) n) != cudaSuccess)
    tris_out = nullptr;

  bool ok = tris_out != nullptr;
  for (void* ptr : scratch)
    ok = ok && (ptr != nullptr);
  if (!ok) {
    LB_FREE_ALL();
    if (tris_out)
      cudaFree(tris_out);
    lumb200_set_last_error("out of device memory during BVH build (%u primitives)", n);
    return LUMB200_ERROR_OUT_OF_MEMORY;
  }

  k_bounds_init<<<1, 1, 0, stream>>>(bounds);
  k_prim_boxes<<<blocks, BUILD_THREADS, 0, stream>>>(world_tris, n, box_lo, box_hi, bounds);
  k_morton<<<blocks, BUILD_THREADS, 0, stream>>>(box_lo, box_hi, n, bounds, keys, vals);

  size_t temp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys, keys_s, vals, vals_s, (int) n, 0, 63, stream);
  void* temp = alloc(temp_bytes);
  if (!temp) {
    LB_FREE_ALL();
    cudaFree(tris_out);
    lumb200_set_last_error("out of device memory during BVH build (sort scratch)");
    return LUMB200_ERROR_OUT_OF_MEMORY;
  }
  cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_s, vals, vals_s, (int) n, 0, 63, stream);

  // SAH-optimal collapse (default) needs the DP tables; LUMB200_COLLAPSE=greedy keeps the largest-area-first heuristic
  CollapseDp dp;
  dp.cost   = nullptr;
  dp.choice = nullptr;
  dp.c_node = 1.0f;
  dp.c_prim = LB_SAH_C_PRIM;
  if (const char* e = getenv("LUMB200_SAH_CPRIM"))
    dp.c_prim = (float) atof(e);
  bool use_dp = n > 1;
  if (const char* ce = getenv("LUMB200_COLLAPSE"))
    use_dp = use_dp && strcmp(ce, "greedy") != 0;
  bool dp_done = false;
  uint32_t* leaf_order = vals_s;  // leaf position -> primitive, in the order the collapse addresses leaf ranges
  WorkItem root;
  root.bvh8 = 0;
  root.bvh2 = 0;
  auto run_dp = [&]() {
    cudaMemsetAsync(t.flags, 0, sizeof(uint32_t) * ni, stream);
    k_collapse_dp<<<blocks, BUILD_THREADS, 0, stream>>>((int) n, t, leaf_order, box_lo, box_hi, dp);
  };
  // C(root, 1) of the DP over the half area of the root box: the SAH cost of the collapsed 8-wide tree (c_node = 1, c_prim as above)
  auto root_sah = [&]() -> float {
    float c = 0.0f;
    float4 lo, hi;
    cudaMemcpyAsync(&c, dp.cost + 7 * (size_t) root.bvh2, sizeof(float), cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(&lo, t.lo + root.bvh2, sizeof(float4), cudaMemcpyDeviceToHost, stream);
    cudaMemcpyAsync(&hi, t.hi + root.bvh2, sizeof(float4), cudaMemcpyDeviceToHost, stream);
    if (cudaStreamSynchronize(stream) != cudaSuccess)
      return FLT_MAX;
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    const float a  = dx * dy + dy * dz + dz * dx;
    return (a > 0.0f) ? c / a : c;
  };

  const char* builder_env = getenv("LUMB200_BVH_BUILDER");
  const bool use_lbvh     = builder_env && strcmp(builder_env, "lbvh") == 0;

  if (n == 1) {
    root.bvh2 = LEAF_FLAG | 0u;
  }
  else if (use_lbvh) {
    cudaMemsetAsync(t.flags, 0, sizeof(uint32_t) * ni, stream);
    k_hierarchy<<<(ni + BUILD_THREADS - 1) / BUILD_THREADS, BUILD_THREADS, 0, stream>>>(keys_s, (int) n, t);
    k_refit<<<blocks, BUILD_THREADS, 0, stream>>>((int) n, t, vals_s, box_lo, box_hi);
    root.bvh2 = 0;
  }
  else {
    // PLOC search radius: a fixed value (LUMB200_PLOC_RADIUS=<n>) or, by default, the candidate whose collapsed 8-wide tree has
    // the lowest SAH cost. No single radius is right (profiles/r1_sweeps_v5.md: 32 saves 1.1 node visits per ray on terrain-10M
    // and costs 0.75 on atrium-1M); the collapse DP below yields the SAH cost of the final tree for free, so every candidate
    // is clustered + costed once and the winner is clustered again (PLOC is deterministic) when it was not the last one.
    std::vector<int> radii = {8, 16, 32, 64};
    if (const char* e = getenv("LUMB200_PLOC_RADIUS")) {
      if (strcmp(e, "auto") != 0)
        radii.assign(1, max(1, atoi(e)));
    }
    if (n < 4096)
      radii.assign(1, 16);  // tiny trees (the emitter BVH of most scenes): nothing to choose
    uint32_t* ids[2]  = {(uint32_t*) alloc(sizeof(uint32_t) * n), (uint32_t*) alloc(sizeof(uint32_t) * n)};
    float4* clo[2]    = {(float4*) alloc(sizeof(float4) * n), (float4*) alloc(sizeof(float4) * n)};
    float4* chi[2]    = {(float4*) alloc(sizeof(float4) * n), (float4*) alloc(sizeof(float4) * n)};
    uint32_t* nn      = (uint32_t*) alloc(sizeof(uint32_t) * n);
    uint32_t* keep    = (uint32_t*) alloc(sizeof(uint32_t) * n);
    uint32_t* offset  = (uint32_t*) alloc(sizeof(uint32_t) * n);
    uint32_t* newpos  = (uint32_t*) alloc(sizeof(uint32_t) * n);
    uint32_t* ploc_ct = (uint32_t*) alloc(sizeof(uint32_t) * 2);  // [0] node counter, [1] clusters after compaction
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, keep, offset, (int) n, stream);
    void* scan_temp = alloc(scan_bytes);
    dp.cost         = (float*) alloc(sizeof(float) * 7 * (size_t) ni);
    dp.choice       = (uint8_t*) alloc(8 * (size_t) ni);
    const int reinsert_passes    = (n >= 64) ? (getenv("LUMB200_BVH_REINSERT") ? atoi(getenv("LUMB200_BVH_REINSERT")) : LB_BVH_REINSERT_PASSES) : 0;
    const float reinsert_min_gain = getenv("LUMB200_BVH_REINSERT_MIN_GAIN") ? (float) atof(getenv("LUMB200_BVH_REINSERT_MIN_GAIN")) : LB_BVH_REINSERT_MIN_GAIN;
    ReinsertPlan* rin_plan       = nullptr;
    unsigned long long* rin_lock = nullptr;  // [0, cands): topology locks, [cands, 2 cands): path marks
    double* rin_area             = nullptr;
    if (reinsert_passes > 0) {
      rin_plan = (ReinsertPlan*) alloc(sizeof(ReinsertPlan) * ((size_t) ni + n));
      rin_lock = (unsigned long long*) alloc(sizeof(unsigned long long) * 2 * ((size_t) ni + n));
      rin_area = (double*) alloc(sizeof(double));
    }
    bool ploc_ok    = scan_temp != nullptr;
    for (void* ptr : scratch)
      ploc_ok = ploc_ok && (ptr != nullptr);
    if (!ploc_ok) {
      LB_FREE_ALL();
      cudaFree(tris_out);
      lumb200_set_last_error("out of device memory during BVH build (PLOC scratch)");
      return LUMB200_ERROR_OUT_OF_MEMORY;
    }
    // clusters the sorted primitives with the given radius into t / vals (leaf order); returns false on failure
    auto cluster = [&](int radius, bool optimise) -> bool {
      cudaMemsetAsync(ploc_ct, 0, sizeof(uint32_t) * 2, stream);
      k_ploc_init<<<blocks, BUILD_THREADS, 0, stream>>>(n, vals_s, box_lo, box_hi, ids[0], clo[0], chi[0]);
      uint32_t m = n;
      int cur    = 0;
      int guard  = 0;
      while (m > 1) {
        const uint32_t mb = (m + BUILD_THREADS - 1) / BUILD_THREADS;
        k_ploc_nn<<<mb, BUILD_THREADS, 0, stream>>>(m, radius, clo[cur], chi[cur], nn);
        k_ploc_merge<<<mb, BUILD_THREADS, 0, stream>>>(m, ids[cur], clo[cur], chi[cur], nn, t, ploc_ct, keep);
        cub::DeviceScan::ExclusiveSum(scan_temp, scan_bytes, keep, offset, (int) m, stream);
        k_ploc_compact<<<mb, BUILD_THREADS, 0, stream>>>(m, keep, offset, ids[cur], clo[cur], chi[cur], ids[cur ^ 1], clo[cur ^ 1], chi[cur ^ 1],
                                                         ploc_ct + 1);
        uint32_t m_next = 0;
        cudaMemcpyAsync(&m_next, ploc_ct + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
        cudaError_t perr = cudaStreamSynchronize(stream);
        if (perr != cudaSuccess || m_next == 0 || m_next >= m || ++guard > 4096) {
          lumb200_set_last_error("PLOC clustering failed (%u -> %u clusters): %s", m, m_next, cudaGetErrorString(perr));
          return false;
        }
        m = m_next;
        cur ^= 1;
      }
      // the last merge created the root
      uint32_t root_id = 0;
      cudaMemcpyAsync(&root_id, ids[cur], sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
      cudaStreamSynchronize(stream);
      root.bvh2 = root_id;
      {
        static const int passes = getenv("LUMB200_BVH_ROTATIONS") ? atoi(getenv("LUMB200_BVH_ROTATIONS")) : LB_BVH_ROTATION_PASSES;
        for (int pass = 0; pass < passes; pass++) {
          cudaMemsetAsync(t.flags, 0, sizeof(uint32_t) * ni, stream);
          cudaMemsetAsync(counters + 3, 0, sizeof(uint32_t), stream);
          k_bvh2_rotate<<<blocks, BUILD_THREADS, 0, stream>>>(n, t, vals_s, box_lo, box_hi, counters + 3);
          if (getenv("LUMB200_BVH_VERBOSE")) {
            uint32_t nrot = 0;
            cudaMemcpyAsync(&nrot, counters + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
            cudaStreamSynchronize(stream);
            fprintf(stderr, "[lumb200] rotation pass %d: %u rotations\n", pass, nrot);
          }
        }
        cudaMemsetAsync(t.flags, 0, sizeof(uint32_t) * ni, stream);
      }
      if (optimise && reinsert_passes > 0) {
        const bool verbose   = getenv("LUMB200_BVH_VERBOSE") != nullptr;
        const uint32_t cands = ni + n;
        const uint32_t cb    = (cands + BUILD_THREADS - 1) / BUILD_THREADS;
        auto area_sum = [&]() -> double {
          double a = 0.0;
          cudaMemsetAsync(rin_area, 0, sizeof(double), stream);
          k_bvh2_area_sum<<<(ni + BUILD_THREADS - 1) / BUILD_THREADS, BUILD_THREADS, 0, stream>>>(ni, t, rin_area);
          cudaMemcpyAsync(&a, rin_area, sizeof(double), cudaMemcpyDeviceToHost, stream);
          cudaStreamSynchronize(stream);
          return a;
        };
        // passes run until one of them improves the inner-node area sum by less than LB_BVH_REINSERT_MIN_GAIN (the estimates of
        // crossing moves make the last per cent oscillate) or the pass budget is spent
        double area_prev = area_sum();
        if (verbose)
          fprintf(stderr, "[lumb200] reinsertion: inner-node area sum %.6g before\n", area_prev);
        for (int pass = 0; pass < reinsert_passes; pass++) {
          const auto t0 = std::chrono::steady_clock::now();
          cudaMemsetAsync(rin_lock, 0, sizeof(unsigned long long) * 2 * (size_t) cands, stream);
          cudaMemsetAsync(counters + 3, 0, sizeof(uint32_t), stream);
          k_reinsert_find<<<cb, BUILD_THREADS, 0, stream>>>(n, ni, t, vals_s, box_lo, box_hi, rin_plan, rin_lock);
          k_reinsert_decide_topo<<<cb, BUILD_THREADS, 0, stream>>>(n, ni, t, rin_plan, rin_lock, rin_lock + cands);
          k_reinsert_decide_path<<<cb, BUILD_THREADS, 0, stream>>>(n, ni, t, rin_plan, rin_lock + cands);
          k_reinsert_apply<<<cb, BUILD_THREADS, 0, stream>>>(n, ni, t, rin_plan, counters + 3);
          cudaMemsetAsync(t.flags, 0, sizeof(uint32_t) * ni, stream);
          k_refit_count<<<blocks, BUILD_THREADS, 0, stream>>>(n, t, vals_s, box_lo, box_hi);
          uint32_t moves = 0, root_count = 0;
          cudaMemcpyAsync(&moves, counters + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
          cudaMemcpyAsync(&root_count, t.count + root_id, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
          if (cudaStreamSynchronize(stream) != cudaSuccess) {
            lumb200_set_last_error("BVH reinsertion pass %d failed: %s", pass, cudaGetErrorString(cudaGetLastError()));
            return false;
          }
          if (root_count != n) {  // a subtree came loose: must not happen (see the comment of section 4c), never traverse such a tree
            lumb200_set_last_error("BVH reinsertion pass %d left %u of %u primitives under the root", pass, root_count, n);
            return false;
          }
          const double area_now = area_sum();
          if (verbose)
            fprintf(stderr, "[lumb200] reinsertion pass %d: %u moves, inner-node area sum %.6g, %.2f ms\n", pass, moves, area_now,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
          const bool converged = moves == 0 || area_prev - area_now < (double) reinsert_min_gain * area_prev;
          area_prev            = area_now;
          if (converged)
            break;
        }
        cudaMemsetAsync(t.flags, 0, sizeof(uint32_t) * ni, stream);
      }
      k_ploc_leaf_positions<<<blocks, BUILD_THREADS, 0, stream>>>(n, t, vals_s, newpos, vals);
      k_ploc_node_first<<<(ni + BUILD_THREADS - 1) / BUILD_THREADS, BUILD_THREADS, 0, stream>>>(ni, t);
      k_ploc_fix_refs<<<(ni + BUILD_THREADS - 1) / BUILD_THREADS, BUILD_THREADS, 0, stream>>>(ni, t, newpos);
      return true;
    };
    leaf_order     = vals;
    int best_r     = radii[0];
    float best_sah = FLT_MAX;
    int built_r    = -1;
    for (int r : radii) {
      if (!cluster(r, radii.size() == 1)) {
        LB_FREE_ALL();
        cudaFree(tris_out);
        return LUMB200_ERROR_API_EXCEPTION;
      }
      built_r = r;
      if (!use_dp)
        break;
      run_dp();
      const float sah = root_sah();
      if (getenv("LUMB200_BVH_VERBOSE"))
        fprintf(stderr, "[lumb200] PLOC radius %d: SAH cost of the collapsed tree %.4f (%u primitives)\n", r, sah, n);
      if (sah < best_sah) {
        best_sah = sah;
        best_r   = r;
      }
    }
    if (use_dp && (built_r != best_r || (reinsert_passes > 0 && radii.size() > 1))) {
      if (!cluster(best_r, true)) {
        LB_FREE_ALL();
        cudaFree(tris_out);
        return LUMB200_ERROR_API_EXCEPTION;
      }
      run_dp();
      if (reinsert_passes > 0) {
        const float sah = root_sah();
        if (getenv("LUMB200_BVH_VERBOSE"))
          fprintf(stderr, "[lumb200] PLOC radius %d + reinsertion: SAH cost of the collapsed tree %.4f (was %.4f)\n", best_r, sah, best_sah);
        best_sah = sah;
      }
    }
    dp_done          = use_dp;
    out->ploc_radius = use_dp ? best_r : radii[0];
    out->sah_cost    = use_dp ? best_sah : 0.0f;
  }

  // SAH-optimal collapse decisions (default); LUMB200_COLLAPSE=greedy keeps the largest-area-first heuristic
  const uint8_t* dp_choice = nullptr;
  if (use_dp) {
    if (!dp_done) {
      if (!dp.cost)
        dp.cost = (float*) alloc(sizeof(float) * 7 * (size_t) ni);
      if (!dp.choice)
        dp.choice = (uint8_t*) alloc(8 * (size_t) ni);
      if (!dp.cost || !dp.choice) {
        LB_FREE_ALL();
        cudaFree(tris_out);
        lumb200_set_last_error("out of device memory during BVH build (collapse tables)");
        return LUMB200_ERROR_OUT_OF_MEMORY;
      }
      run_dp();
      out->sah_cost = root_sah();
    }
    dp_choice = dp.choice;
  }

  // counters: [0] next bvh8 node index, [1] next triangle slot, [2] items written to the next queue
  uint32_t h_counters[4] = {1, 0, 0, 0};
  cudaMemcpyAsync(counters, h_counters, sizeof(h_counters), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(q0, &root, sizeof(root), cudaMemcpyHostToDevice, stream);

  uint32_t n_items = 1;
  WorkItem* qin    = q0;
  WorkItem* qout   = q1;
  int level        = 0;
  while (n_items > 0) {
    k_collapse<<<(n_items + 63) / 64, 64, 0, stream>>>(qin, n_items, qout, counters, t, leaf_order, box_lo, box_hi, world_tris, nodes_tmp, tris_out,
                                                       dp_choice);
    cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, stream);
    cudaError_t err = cudaStreamSynchronize(stream);
    if (err != cudaSuccess) {
      LB_FREE_ALL();
      cudaFree(tris_out);
      lumb200_set_last_error("BVH collapse failed at level %d: %s", level, cudaGetErrorString(err));
      return LUMB200_ERROR_CUDA;
    }
    n_items       = h_counters[2];
    h_counters[2] = 0;
    cudaMemcpyAsync(counters + 2, &h_counters[2], sizeof(uint32_t), cudaMemcpyHostToDevice, stream);
    WorkItem* tmp = qin;
    qin           = qout;
    qout          = tmp;
    level++;
    if (level > 256) {
      LB_FREE_ALL();
      cudaFree(tris_out);
      lumb200_set_last_error("BVH collapse did not terminate");
      return LUMB200_ERROR_API_EXCEPTION;
    }
  }

  const uint32_t num_nodes = h_counters[0];
  if (h_counters[1] != n) {
    LB_FREE_ALL();
    cudaFree(tris_out);
    lumb200_set_last_error("BVH collapse emitted %u triangles, expected %u", h_counters[1], n);
    return LUMB200_ERROR_API_EXCEPTION;
  }

  uint4* nodes_final = nullptr;
  if (cudaMalloc(&nodes_final, sizeof(Bvh8Node) * (size_t) num_nodes) != cudaSuccess) {
    LB_FREE_ALL();
    cudaFree(tris_out);
    lumb200_set_last_error("out of device memory for BVH nodes");
    return LUMB200_ERROR_OUT_OF_MEMORY;
  }
  cudaMemcpyAsync(nodes_final, nodes_tmp, sizeof(Bvh8Node) * (size_t) num_nodes, cudaMemcpyDeviceToDevice, stream);
  cudaEventRecord(ev1, stream);
  cudaError_t err = cudaStreamSynchronize(stream);
  LB_FREE_ALL();
  if (err != cudaSuccess) {
    cudaFree(tris_out);
    cudaFree(nodes_final);
    lumb200_set_last_error("BVH build failed: %s", cudaGetErrorString(err));
    return LUMB200_ERROR_CUDA;
  }
  float ms = 0.0f;
  cudaEventElapsedTime(&ms, ev0, ev1);
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  if (build_ms)
    *build_ms = ms;

  out->nodes     = nodes_final;
  out->tris      = tris_out;
  out->num_nodes = num_nodes;
  out->num_tris  = n;
  out->bytes     = sizeof(Bvh8Node) * (size_t) num_nodes + sizeof(float4) * 3 * (size_t) n;
  out->depth     = (uint32_t) level;  // the collapse is level-synchronous: one iteration per level of the 8-wide tree
  return LUMB200_SUCCESS;
}
