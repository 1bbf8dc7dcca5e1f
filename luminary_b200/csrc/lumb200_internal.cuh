// lumb200_internal.cuh - device-side data layouts and helpers shared by the CUDA translation units.
//
// Layout summary (all arrays live in HBM, scene replicated per GPU):
//   tri_soup   float4[3 * P]   world-space triangles in BVH leaf order; v0.w = flattened prim index (bits)
//   nodes      Bvh8Node[N8]    80-byte compressed 8-wide nodes (5 x 16-byte loads)
//   prim_handle uint2[P]       flattened prim index -> (instance_id, tri_id)   (TriangleHandle, device_utils.h:235-238)
//   mesh vertex float4[3*T]    DeviceTriangleVertex {pos, packed normal}       (device_structs.h:270-273)
//   mesh textri uint4[T]       DeviceTriangleTexture {3 packed uv, material}   (device_structs.h:275-281)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lumb200.h"

#define LB_CHECK(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      lumb200_set_last_error("CUDA error %s (%s) in %s at %s:%d", cudaGetErrorName(_e), cudaGetErrorString(_e), #expr, __FILE__, __LINE__); \
      return LUMB200_ERROR_CUDA;                                                                    \
    }                                                                                               \
  } while (0)

extern "C" void lumb200_set_last_error(const char* fmt, ...);

#define LB_HIT_SKY 0xFFFFFFFEu
#define LB_PRIM_NONE 0xFFFFFFFFu
#define LB_LIGHT_ID_INVALID 0xFFFFFFFFu

// ---------------------------------------------------------------------------------------------
// BVH8 node, after Ylitie, Karras, Laine, "Efficient Incoherent Ray Traversal on GPUs Through
// Compressed Wide BVHs", HPG 2017. 80 bytes = 5 x uint4.
// ---------------------------------------------------------------------------------------------
struct __align__(16) Bvh8Node {
  float px, py, pz;       // quantisation origin (node box low corner)
  uint8_t ex, ey, ez;     // per-axis exponents (biased IEEE exponent of the grid step)
  uint8_t imask;          // bit i set: slot i holds an inner node
  uint32_t child_base;    // index of the first inner child
  uint32_t tri_base;      // index of the first triangle referenced by leaf slots
  uint8_t meta[8];        // inner: 0b001xxxxx with xxxxx = 24 + slot; leaf: unary tri count << 5 | tri offset; empty: 0
  uint8_t qlox[8], qloy[8], qloz[8];
  uint8_t qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Bvh8Node) == 80, "Bvh8Node must be 80 bytes");

struct Bvh8 {
  const uint4* nodes;    // Bvh8Node as uint4[5]
  const float4* tris;    // 3 float4 per triangle, leaf order
  uint32_t num_nodes;
  uint32_t num_tris;
};

// ---------------------------------------------------------------------------------------------
// Per-mesh / per-instance tables
// ---------------------------------------------------------------------------------------------
struct LbTransform {  // DeviceTransform, device_structs.h:292-297 (32 bytes)
  float tx, ty, tz;
  float sx, sy, sz;
  uint16_t qx, qy, qz, qw;
};
static_assert(sizeof(LbTransform) == 32, "LbTransform must be 32 bytes");

struct LbSceneTables {
  const float4* const* mesh_vertices;  // [mesh] -> float4[3 * tris]
  const uint4* const* mesh_textris;    // [mesh] -> uint4[tris]
  const uint32_t* instance_mesh;       // [instance]
  const LbTransform* instance_transform;
  const uint32_t* instance_prim_offset;  // [instance + 1]
  const uint2* prim_handle;              // [prim] (instance_id, tri_id)
  const uint4* materials;                // DeviceMaterialCompressed as 2 x uint4
  uint32_t num_instances;
  uint32_t num_prims;
  uint32_t num_materials;
};

// ---------------------------------------------------------------------------------------------
// Exact (non-contracted) fp32 helpers. The geometry translation units are compiled with
// -fmad=false, so a*b+c is two IEEE roundings exactly like the CPU oracle built with
// -ffp-contract=off; these helpers only fix the association order.
// ---------------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
};

__host__ __device__ __forceinline__ V3 v3(float x, float y, float z) {
  V3 r;
  r.x = x;
  r.y = y;
  r.z = z;
  return r;
}
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross3(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// quaternion_apply, reference cuda/math.cuh:411-427
__host__ __device__ __forceinline__ V3 quat_apply(float qx, float qy, float qz, float qw, V3 v) {
  const V3 u         = v3(qx, qy, qz);
  const float s      = qw;
  const float dot_uv = dot3(u, v);
  const float dot_uu = dot3(u, u);
  const V3 c         = cross3(u, v);
  V3 r               = u * (2.0f * dot_uv);
  r                  = r + v * (s * s - dot_uu);
  r                  = r + c * (2.0f * s);
  return r;
}

// transform_apply = rotate (quaternion16) -> scale -> translate, reference cuda/math.cuh:429-491
__host__ __device__ __forceinline__ V3 transform_rotate(const LbTransform& t, V3 v) {
  const float qx = (t.qx * (1.0f / 0x7FFF)) - 1.0f;
  const float qy = (t.qy * (1.0f / 0x7FFF)) - 1.0f;
  const float qz = (t.qz * (1.0f / 0x7FFF)) - 1.0f;
  const float qw = (t.qw * (1.0f / 0x7FFF)) - 1.0f;
  return quat_apply(qx, qy, qz, qw, v);
}
__host__ __device__ __forceinline__ V3 transform_rotate_inv(const LbTransform& t, V3 v) {
  const float qx = 1.0f - (t.qx * (1.0f / 0x7FFF));
  const float qy = 1.0f - (t.qy * (1.0f / 0x7FFF));
  const float qz = 1.0f - (t.qz * (1.0f / 0x7FFF));
  const float qw = (t.qw * (1.0f / 0x7FFF)) - 1.0f;
  return quat_apply(qx, qy, qz, qw, v);
}
__host__ __device__ __forceinline__ V3 transform_relative(const LbTransform& t, V3 v) {
  return transform_rotate(t, v) * v3(t.sx, t.sy, t.sz);
}
__host__ __device__ __forceinline__ V3 transform_point(const LbTransform& t, V3 v) {
  return transform_relative(t, v) + v3(t.tx, t.ty, t.tz);
}
__host__ __device__ __forceinline__ V3 transform_point_inv(const LbTransform& t, V3 v) {
  const V3 inv = v3(1.0f / t.sx, 1.0f / t.sy, 1.0f / t.sz);
  return transform_rotate_inv(t, (v - v3(t.tx, t.ty, t.tz)) * inv);
}

// ---------------------------------------------------------------------------------------------
// Build entry points (bvh_build.cu)
// ---------------------------------------------------------------------------------------------
struct LbBvhBuffers {
  uint4* nodes      = nullptr;
  float4* tris      = nullptr;
  uint32_t num_nodes = 0;
  uint32_t num_tris  = 0;
  size_t bytes       = 0;
  uint32_t depth     = 0;     // levels of the 8-wide tree (root = 1)
  float sah_cost     = 0.0f;  // SAH cost of the collapsed tree, C(root) / A(root) with c_node = 1 (0 when the greedy collapse ran)
  int ploc_radius    = 0;     // search radius the selected binary hierarchy was clustered with
};

// world_tris: 3 float4 per prim in flattened order (w of v0 = prim index bits). Builds into `out` (allocates).
Lumb200Result lb_bvh8_build(const float4* world_tris, uint32_t num_prims, LbBvhBuffers* out, cudaStream_t stream, float* build_ms);
void lb_bvh8_free(LbBvhBuffers* b);
Lumb200Result lb_flatten_instances(const LbSceneTables& tables, float4* world_tris, uint2* prim_handle, cudaStream_t stream);
