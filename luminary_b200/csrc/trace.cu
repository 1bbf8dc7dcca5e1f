// trace.cu - ray generation, closest-hit and shadow traversal kernels, and the material-keyed queue sort.
//
// Replaces, on the reference's per-bounce path (device/device_renderer.c:53-134):
//   tasks_create                 cuda/kernels.cuh:45-193            -> k_raygen
//   __raygen__optix (raytrace)   optix/optix_kernel_raytrace.cu:147  -> k_trace_closest
//   volume_process_events counts cuda/volume.cuh:204-228  }
//   tasks_sort                   cuda/kernels.cuh:394-484 }          -> k_sort_count / k_sort_scan / k_sort_scatter
//   __raygen__optix (shadow)     optix/optix_kernel_shadow.cu:15-100 -> k_trace_shadow (transmittance part)
// Compiled with -fmad=false (see traverse.cuh).
#include <float.h>
#include <stdlib.h>

#include "rng.cuh"
#include "texture.cuh"
#include "trace_loop.cuh"
#include "wavefront.cuh"

#define TRACE_THREADS LB_TRACE_THREADS

// ---------------------------------------------------------------------------------------------
// ray generation: thin-lens camera (cuda/camera.cuh:11-38, camera_thin_lens.cuh:8-86) in the oracle's
// operation order, so primary rays are bit-identical to oracle/orc_core.c: orc_camera_sample.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ V3 normalize3(V3 a) { return a * (1.0f / sqrtf(dot3(a, a))); }

// per-bounce queue counters (the persistent fetch cursor, the hit count, the shadow / enumeration queues)
__device__ __forceinline__ void reset_fetch_cursors(LbCounters* C) { C->fetch = C->fetch_enum = C->fetch_shadow = 0; }

__device__ __forceinline__ void reset_bounce_counters(LbCounters* C) {
  reset_fetch_cursors(C);
  C->n_hits      = 0;
#pragma unroll
  for (int slot = 0; slot < LB_NEE_SLOTS; slot++)
    C->n_shadow[slot] = 0;
  C->n_enum      = 0;
}

__device__ __forceinline__ void camera_sample(const LbCameraDev& cam, const LbFrame& F, const uint32_t* __restrict__ bluenoise, uint32_t px,
                                              uint32_t py, uint32_t sample_id, V3& origin, V3& dir) {
  const uint2 jq = lbrng::random_2d_bits(bluenoise, lbrng::T_CAMERA_JITTER, 0, 0, sample_id, 0);
  const float jx = lbrng::u32_to_float(jq.x);
  const float jy = lbrng::u32_to_float(jq.y);

  const float step = 2.0f * (cam.fov / F.width);
  const float vfov = step * F.height * 0.5f;

  V3 sensor;
  sensor.x = cam.fov - step * (px + jx);
  sensor.y = -vfov + step * (py + jy);
  sensor.z = 1.0f;

  const V3 s2f             = normalize3(v3(0.0f, 0.0f, 0.0f) - sensor);
  const float focal_length = fmaxf(cam.object_distance * (1.0f / 0.001f), 0.01f);
  const V3 focal_point     = s2f * (-focal_length / s2f.z);

  V3 aperture = v3(0.0f, 0.0f, 0.0f);
  if (cam.aperture_size != 0.0f) {
    lbrng::Sampler smp;
    smp.bluenoise = bluenoise, smp.px = px, smp.py = py, smp.sample_id = sample_id, smp.depth = 0;
    const float2 r      = smp.get2(lbrng::T_LENS);
    const float ap_size = cam.aperture_size * (1.0f / 0.001f);
    const float PI      = 3.141592653589f;
    if (cam.aperture_shape == 1) {
      const int blade        = (int) (smp.get1(lbrng::T_LENS_BLADE) * cam.aperture_blade_count);
      const float alpha      = sqrtf(r.x);
      const float beta       = r.y;
      const float u          = 1.0f - alpha;
      const float v          = alpha * beta;
      const float angle_step = (2.0f * PI) / cam.aperture_blade_count;
      const float a1         = angle_step * blade;
      const float a2         = angle_step * (blade + 1);
      aperture.x             = (sinf(a1) * u + sinf(a2) * v) * ap_size;
      aperture.y             = (cosf(a1) * u + cosf(a2) * v) * ap_size;
    }
    else {
      const float alpha = r.x * 2.0f * PI;
      const float beta  = sqrtf(r.y) * ap_size;
      aperture.x        = cosf(alpha) * beta;
      aperture.y        = sinf(alpha) * beta;
    }
  }

  V3 o = aperture;
  V3 d = normalize3(focal_point - aperture);

  o = quat_apply(cam.qx, cam.qy, cam.qz, cam.qw, o);
  o = o * (cam.camera_scale * 0.001f);
  o = o + v3(cam.px, cam.py, cam.pz);
  d = quat_apply(cam.qx, cam.qy, cam.qz, cam.qw, d);

  origin = o;
  dir    = d;
}

// tile_order: slot i of the wavefront holds the pixel number i of an 8 x 4 TILED enumeration of the frame instead of pixel i, so that the 32
// rays a warp of the traversal kernels fetches together cover a compact tile instead of a 32 x 1 strip (fewer distinct nodes per warp
// step). The tiles cover the part of the frame that is a multiple of 8 x 4; the right and bottom rims follow in row order, so the map is
// a bijection for every resolution. Paths are independent and every consumer goes through P.pixel[], so the image does not change.
__device__ __forceinline__ uint32_t lb_tiled_pixel(uint32_t i, uint32_t width, uint32_t height) {
  const uint32_t w8 = width & ~7u, h4 = height & ~3u;
  if (i < w8 * h4) {
    const uint32_t tile = i >> 5, in = i & 31u, tiles_per_row = w8 >> 3;
    const uint32_t ty = tile / tiles_per_row, tx = tile - ty * tiles_per_row;
    return (tx * 8u + (in & 7u)) + (ty * 4u + (in >> 3)) * width;
  }
  uint32_t r = i - w8 * h4;  // rims: right strip of the tiled rows, then the bottom rows
  const uint32_t rim_w = width - w8;
  if (r < rim_w * h4) {
    const uint32_t y = r / rim_w;
    return (w8 + (r - y * rim_w)) + y * width;
  }
  r -= rim_w * h4;
  return h4 * width + r;
}

__global__ void __launch_bounds__(256) k_raygen(LbPaths P, LbFrame F, LbCameraDev cam, const uint32_t* __restrict__ bluenoise,
                                                uint32_t sample_id, uint32_t* __restrict__ queue, LbCounters* C, bool tile_order) {
  const uint32_t n = F.width * F.height;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t pixel = tile_order ? lb_tiled_pixel(i, F.width, F.height) : i;
    const uint32_t y     = pixel / F.width;
    const uint32_t x     = pixel - y * F.width;
    V3 o, d;
    camera_sample(cam, F, bluenoise, x, y, sample_id, o, d);
    P.org[i]    = make_float4(o.x, o.y, o.z, 0.0f);
    P.dir[i]    = make_float4(d.x, d.y, d.z, FLT_MAX);
    P.prim[i]   = LB_PRIM_NONE;
    // record_pack(1,1,1): 0x3F800000 >> 11 = 0x7F000 per channel
    P.record[i] = make_uint2(0x7F000u | (0x7F000u << 21), (0x7F000u >> 11) | (0x7F000u << 10));
    P.pixel[i]  = pixel;
    P.state[i]  = LB_STATE_DELTA_PATH | LB_STATE_CAMERA_DIRECTION | LB_STATE_ALLOW_EMISSION | LB_STATE_ALLOW_AMBIENT;
    P.medium[i] = 0u;  // medium_stack_ior_modify({}, 1.0f, push): ior_compress(1.0f) == 0
    P.result[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int slot = 0; slot < LB_NEE_SLOTS; slot++)
      P.nee[LB_NEE_SLOTS * (size_t) i + slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    queue[i]    = i;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    C->n_active = n;
    C->n_next   = 0;
    reset_bounce_counters(C);
  }
}

// one primary ray for the G-buffer query of a pixel (luminary_host_get_pixel_info -> device_get_gbuffer_meta): slot 0 of the wavefront
__global__ void k_raygen_pixel(LbPaths P, LbFrame F, LbCameraDev cam, const uint32_t* __restrict__ bluenoise, uint32_t x, uint32_t y,
                               uint32_t sample_id, uint32_t* __restrict__ queue, LbCounters* C) {
  V3 o, d;
  camera_sample(cam, F, bluenoise, x, y, sample_id, o, d);
  P.org[0]  = make_float4(o.x, o.y, o.z, 0.0f);
  P.dir[0]  = make_float4(d.x, d.y, d.z, FLT_MAX);
  P.prim[0] = LB_PRIM_NONE;
  queue[0]  = 0;
  C->n_active = 1;
  C->n_next   = 0;
  reset_bounce_counters(C);
}

// tasks_create_adaptive_sampling (cuda/kernels.cuh:195-356): task -> block (binary search in the prefix sums, adaptive_sampling_find_block)
// -> pixel of the block and sample of the pixel; sample id = samples the pixel already has + local sample. Slot t of the wavefront holds
// task task_begin + t; tasks outside the image or beyond 2^20 samples leave their slot empty (pixel = ~0).
__global__ void __launch_bounds__(256) k_raygen_adaptive(LbPaths P, LbFrame F, LbCameraDev cam, const uint32_t* __restrict__ bluenoise, LbAdaptive A,
                                                         uint32_t stage, const uint32_t* __restrict__ task_prefix, uint32_t num_blocks,
                                                         uint32_t task_begin, uint32_t n_tasks, uint32_t* __restrict__ queue, LbCounters* C) {
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t t0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; t0 < n_tasks; t0 += gridDim.x * blockDim.x) {
    const uint32_t t = t0 + lane;
    bool valid       = false;
    if (t < n_tasks) {
      const uint32_t task_id = task_begin + t;
      uint32_t lo = 0, hi = num_blocks - 1u;
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (task_id < __ldg(task_prefix + mid))
          hi = mid;
        else
          lo = mid + 1u;
      }
      const uint32_t block = lo;
      const uint32_t base  = block ? __ldg(task_prefix + block - 1u) : 0u;
      const uint32_t word  = __ldg(A.words + block);
      const uint32_t tpp   = as_stage_count(word, stage - 1u);
      const uint32_t local = task_id - base;
      const uint32_t lp = local / tpp, ls = local - lp * tpp;
      const uint32_t by = block / A.bw, bx = block - by * A.bw;
      const uint32_t x = (bx << 2) + (lp & 3u), y = (by << 2) + (lp >> 2);
      const uint32_t sample_id = as_block_samples(word, A) + ls;
      P.pixel[t] = 0xFFFFFFFFu;
      if (x < F.width && y < F.height && sample_id < (1u << 20)) {
        valid = true;
        V3 o, d;
        camera_sample(cam, F, bluenoise, x, y, sample_id, o, d);
        P.org[t]       = make_float4(o.x, o.y, o.z, 0.0f);
        P.dir[t]       = make_float4(d.x, d.y, d.z, FLT_MAX);
        P.prim[t]      = LB_PRIM_NONE;
        P.record[t]    = make_uint2(0x7F000u | (0x7F000u << 21), (0x7F000u >> 11) | (0x7F000u << 10));
        P.pixel[t]     = x + y * F.width;
        P.state[t]     = LB_STATE_DELTA_PATH | LB_STATE_CAMERA_DIRECTION | LB_STATE_ALLOW_EMISSION | LB_STATE_ALLOW_AMBIENT;
        P.medium[t]    = 0u;
        P.sample_id[t] = sample_id;
        P.result[t]    = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int slot = 0; slot < LB_NEE_SLOTS; slot++)
          P.nee[LB_NEE_SLOTS * (size_t) t + slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
    }
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, valid);
    if (mask) {
      uint32_t pos = 0;
      if (lane == (uint32_t) (__ffs(mask) - 1))
        pos = atomicAdd(&C->n_active, (uint32_t) __popc(mask));
      pos = __shfl_sync(0xFFFFFFFFu, pos, __ffs(mask) - 1);
      if (valid)
        queue[pos + __popc(mask & ((1u << lane) - 1u))] = t;
    }
  }
}

__global__ void k_reset_counters(LbCounters* C) {
  C->n_active = 0;
  C->n_next   = 0;
  reset_bounce_counters(C);
}

void lb_launch_raygen_adaptive(const LbPaths& P, const LbFrame& F, const LbCameraDev& cam, const uint32_t* bluenoise, const LbAdaptive& A,
                               uint32_t stage, const uint32_t* task_prefix, uint32_t num_blocks, uint32_t task_begin, uint32_t n_tasks,
                               uint32_t* queue, LbCounters* C, int grid, cudaStream_t s) {
  k_reset_counters<<<1, 1, 0, s>>>(C);
  k_raygen_adaptive<<<grid, 256, 0, s>>>(P, F, cam, bluenoise, A, stage, task_prefix, num_blocks, task_begin, n_tasks, queue, C);
}

// ---------------------------------------------------------------------------------------------
// closest hit: persistent warps, one ray per lane, replacement rays fetched per lane (trace_loop.cuh)
// Semantics of the reference (optix_kernel_raytrace.cu:82-95, optix_anyhit.cuh:15-31): tmin 0, tmax FLT_MAX, the
// ignore handle is rejected, and so are hits on texels of an albedo texture whose alpha is 0 (optix_alpha_test,
// optix_common.cuh:20-46). kTex = false is the variant for scenes without albedo textures: no per-hit material lookup.
// ---------------------------------------------------------------------------------------------
template <bool kTex>
struct LbClosestPolicy {
  LbPaths P;
  LbTexScene T;
  const uint32_t* __restrict__ queue;
  float2* __restrict__ uv_out;
  uint32_t i, ignore_prim;
  LbHit best;

  __device__ __forceinline__ uint32_t peek(uint32_t k) const { return queue[k]; }
  __device__ __forceinline__ void prefetch(uint32_t slot) const {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.org + slot));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.dir + slot));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.prim + slot));
  }
  __device__ __forceinline__ void begin(uint32_t k, LbRay& r) {
    i              = queue[k];
    const float4 o = P.org[i];
    const float4 d = P.dir[i];
    r.ox = o.x, r.oy = o.y, r.oz = o.z;
    r.dx = d.x, r.dy = d.y, r.dz = d.z;
    r.tmin      = 0.0f;
    r.tmax      = FLT_MAX;
    ignore_prim = P.prim[i];
    best.prim   = LB_HIT_SKY;
    best.t      = FLT_MAX;
    best.u      = 0.0f;
    best.v      = 0.0f;
  }
  __device__ __forceinline__ bool hit(uint32_t prim, float t, float u, float v, float& tmax) {
    if (prim == ignore_prim)
      return false;
    if (t < best.t || (t == best.t && best.prim != LB_HIT_SKY && prim < best.prim)) {
      if (kTex && lb_alpha_cutout(T, prim, u, v))
        return false;
      best.prim = prim;
      best.t    = t;
      best.u    = u;
      best.v    = v;
      tmax      = t;
    }
    return false;
  }
  __device__ __forceinline__ bool end() {
    P.prim[i]  = best.prim;
    P.dir[i].w = (best.prim == LB_HIT_SKY) ? 3.402823466e+38f : best.t;
    if (uv_out)
      uv_out[i] = make_float2(best.u, best.v);
    return false;
  }
};

#ifndef LB_CLOSEST_MIN_BLOCKS
#define LB_CLOSEST_MIN_BLOCKS 8
#endif
template <bool kCount, bool kTex>
__global__ void __launch_bounds__(TRACE_THREADS, LB_CLOSEST_MIN_BLOCKS) k_trace_closest(Bvh8 bvh, LbPaths P, const uint32_t* __restrict__ queue, LbCounters* C,
                                                                 float2* __restrict__ uv_out, LbTraceTuning tune, LbTexScene T) {
  const uint32_t n = C->n_active;
  LbTraversalCount cnt;
  cnt.nodes = 0, cnt.tris = 0;
  LbClosestPolicy<kTex> pol;
  pol.P      = P;
  pol.T      = T;
  pol.queue  = queue;
  pol.uv_out = uv_out;
  lb_trace_warp<LbClosestPolicy<kTex>, kCount>(bvh, n, &C->fetch, pol, cnt, tune, &C->stack_overflow);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(&C->closest_rays, (unsigned long long) n);
  if (kCount) {
    atomicAdd(&C->closest_nodes, (unsigned long long) cnt.nodes);
    atomicAdd(&C->closest_tris, (unsigned long long) cnt.tris);
  }
}

// ---------------------------------------------------------------------------------------------
// shadow rays: transmittance along the NEE segments (geometry light, BSDF-sampled light, ambient) that k_shade
// appended to the shadow-ray queue - one ray per lane instead of up to three per path.
// Semantics of the reference's shadow any-hit programs (cuda/optix_anyhit.cuh:49-139): skip the target light
// and the surface the ray starts on, stop at the first fully opaque hit (visibility 0), otherwise multiply the
// per-material transparency. shadow_tab[m] = (r, g, b multiplier, w = 1 if opaque), precomputed per material; w = 2 marks
// a material with an albedo texture, whose response is evaluated at the hit's texture coordinates (kTex variant only,
// optix_get_albedo_for_shadowing, optix_common.cuh:48-66).
// The visible part of the contribution is added to the accumulator of the ray's NEE slot (the reference's RMW of
// DeviceTaskResult, direct_lighting.cuh:445-669). A path has at most one ray per slot and bounce, so the plain
// read-modify-write is race-free and the result is deterministic.
// ---------------------------------------------------------------------------------------------
// ray k of the concatenated shadow-queue regions -> queue entry (region s starts at s * capacity)
__device__ __forceinline__ uint32_t lb_shadow_queue_entry(uint32_t k, uint32_t n0, uint32_t n01, uint32_t n012, uint32_t capacity) {
  return (k < n0) ? k : ((k < n01) ? (k - n0) + capacity : ((k < n012) ? (k - n01) + 2u * capacity : (k - n012) + 3u * capacity));
}

template <bool kTex>
struct LbShadowPolicy {
  LbPaths P;
  LbTexScene T;
  const uint16_t* __restrict__ prim_material;
  const float4* __restrict__ shadow_tab;
  uint32_t k, acc, ignore_prim, target_prim;  // acc = LB_NEE_SLOTS * path + slot
  float vr, vg, vb;

  uint32_t n0, n01, n012;  // entries of region 0, of regions 0 + 1, of regions 0 + 1 + 2

  __device__ __forceinline__ uint32_t peek(uint32_t k_) const { return lb_shadow_queue_entry(k_, n0, n01, n012, P.capacity); }
  __device__ __forceinline__ void prefetch(uint32_t entry) const {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.sq_org + entry));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.sq_dir + entry));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.sq_col + entry));
  }
  __device__ __forceinline__ void begin(uint32_t k_, LbRay& r) {
    // ray k_ of the concatenated regions -> queue entry
    k              = lb_shadow_queue_entry(k_, n0, n01, n012, P.capacity);
    const float4 o = P.sq_org[k];
    const float4 d = P.sq_dir[k];
    const uint32_t path = __float_as_uint(o.w) & 0x3FFFFFFFu;
    acc                 = LB_NEE_SLOTS * path + (__float_as_uint(o.w) >> 30);
    r.ox = o.x, r.oy = o.y, r.oz = o.z;
    r.dx = d.x, r.dy = d.y, r.dz = d.z;
    r.tmin      = FLT_EPSILON;
    r.tmax      = d.w;
    ignore_prim = P.prim[path];
    target_prim = __float_as_uint(P.sq_col[k].w);
    vr = vg = vb = 1.0f;
  }
  __device__ __forceinline__ bool hit(uint32_t prim, float t, float u, float v, float& tmax) {
    if (prim == ignore_prim || prim == target_prim)
      return false;
    if (!(t < tmax))  // tmax never shrinks for shadow rays: it is the distance to the light
      return false;
    const uint32_t mid = __ldg(prim_material + prim);
    float4 m           = __ldg(shadow_tab + mid);
    if (kTex && m.w == 2.0f) {
      const uint4 m0 = __ldg(T.materials + 2 * mid), m1 = __ldg(T.materials + 2 * mid + 1);
      m              = lb_shadow_response(lb_shadow_albedo(T, prim, m1.z & 0xFFFFu, u, v), (m0.x & 0x10u) != 0);
    }
    if (m.w != 0.0f) {
      vr = vg = vb = 0.0f;
      return true;
    }
    vr *= m.x;
    vg *= m.y;
    vb *= m.z;
    return false;
  }
  __device__ __forceinline__ bool end() {
    if (vr != 0.0f || vg != 0.0f || vb != 0.0f) {
      const float4 c = P.sq_col[k];
      float4 a       = P.nee[acc];
      a.x += c.x * vr;
      a.y += c.y * vg;
      a.z += c.z * vb;
      P.nee[acc] = a;
    }
    return false;
  }
};

#ifndef LB_SHADOW_MIN_BLOCKS
#define LB_SHADOW_MIN_BLOCKS 8
#endif
template <bool kCount, bool kTex>
__global__ void __launch_bounds__(TRACE_THREADS, LB_SHADOW_MIN_BLOCKS) k_trace_shadow(Bvh8 bvh, LbPaths P, LbCounters* C, const uint16_t* __restrict__ prim_material,
                                                                const float4* __restrict__ shadow_tab, LbTraceTuning tune, LbTexScene T) {
  LbTraversalCount cnt;
  cnt.nodes = 0, cnt.tris = 0;
  const uint32_t n0 = C->n_shadow[0], n1 = C->n_shadow[1], n2 = C->n_shadow[2], n3 = C->n_shadow[3];
  const uint32_t n  = n0 + n1 + n2 + n3;
  LbShadowPolicy<kTex> pol;
  pol.n0            = n0;
  pol.n01           = n0 + n1;
  pol.n012          = n0 + n1 + n2;
  pol.P             = P;
  pol.T             = T;
  pol.prim_material = prim_material;
  pol.shadow_tab    = shadow_tab;
  lb_trace_warp<LbShadowPolicy<kTex>, kCount>(bvh, n, &C->fetch_shadow, pol, cnt, tune, &C->stack_overflow);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(&C->shadow_rays, (unsigned long long) n);
  if (kCount) {
    atomicAdd(&C->shadow_nodes, (unsigned long long) cnt.nodes);
    atomicAdd(&C->shadow_tris, (unsigned long long) cnt.tris);
  }
}

// ---------------------------------------------------------------------------------------------
// emitter enumeration along the BSDF-sampled NEE direction: the reference's light_bsdf_trace any-hit program
// (optix_anyhit.cuh:145-205) reservoir-samples ONE of the emitters a ray pierces, in OptiX's unspecified any-hit order; here (and
// in the oracle) the order is fixed to ascending distance, ties by light id. Round 1 walked the emitter BVH inside k_shade, one
// divergent stack-based traversal per thread in the least occupied kernel; now k_shade only queues the ray and the persistent
// warp loop traces it like any other: every step is a closest-hit query restricted to hits behind the previous one (policy.end()
// asks the loop to restart the ray until an opaque emitter or the end of the ray is reached). The emitter BVH's primitive ids are
// light ids. Results: eq_dir.w = selected light id (bits), eq_hits = emitters counted.
// ---------------------------------------------------------------------------------------------
template <bool kTex>
struct LbEnumPolicy {
  LbPaths P;
  LbTexScene T;
  const uint32_t* __restrict__ light_prims;  // light id -> flattened primitive
  uint32_t k, ignore_prim;
  float prev_t, best_t, best_u, best_v, random;
  uint32_t prev_light, best_light, num_hits, selected, guard;

  __device__ __forceinline__ uint32_t peek(uint32_t) const { return 0xFFFFFFFFu; }
  __device__ __forceinline__ void prefetch(uint32_t) const {}
  __device__ __forceinline__ void begin(uint32_t k_, LbRay& r) {
    k              = k_;
    const float4 o = P.eq_org[k];
    const float4 d = P.eq_dir[k];
    r.ox = o.x, r.oy = o.y, r.oz = o.z;
    r.dx = d.x, r.dy = d.y, r.dz = d.z;
    r.tmin      = FLT_EPSILON;
    r.tmax      = FLT_MAX;
    ignore_prim = P.prim[__float_as_uint(o.w)];
    random      = d.w;
    prev_t      = -1.0f;
    prev_light  = LB_LIGHT_ID_INVALID;
    best_t      = FLT_MAX;
    best_light  = LB_LIGHT_ID_INVALID;
    best_u = best_v = 0.0f;
    num_hits    = 0;
    selected    = LB_LIGHT_ID_INVALID;
    guard       = 0;
  }
  __device__ __forceinline__ bool hit(uint32_t light, float t, float u, float v, float& tmax) {
    const bool after_prev = (t > prev_t) || (t == prev_t && prev_light != LB_LIGHT_ID_INVALID && light > prev_light);
    if (!after_prev)
      return false;
    if (t < best_t || (t == best_t && light < best_light)) {
      best_t     = t;
      best_light = light;
      best_u     = u;
      best_v     = v;
      tmax       = t;
    }
    return false;
  }
  // true = trace the ray again for the next emitter behind the one just processed
  __device__ __forceinline__ bool end() {
    bool again = false;
    if (best_light != LB_LIGHT_ID_INVALID && guard < 64u) {
      guard++;
      again      = true;
      prev_t     = best_t;
      prev_light = best_light;
      const uint32_t prim = __ldg(light_prims + best_light);
      if (prim != ignore_prim) {
        const uint32_t mid = __ldg(T.prim_material + prim);
        const uint4 m0     = __ldg(T.materials + 2 * mid);
        float alpha        = (m0.w >> 16) * (1.0f / 0xFFFF);
        if (kTex) {  // optix_get_albedo_for_shadowing with the barycentrics of the emitter-BVH hit
          const uint32_t atex = __ldg(&T.materials[2 * mid + 1].z) & 0xFFFFu;
          if (atex != LB_TEXTURE_NONE)
            alpha = lb_shadow_albedo(T, prim, atex, best_u, best_v).w;
        }
        const bool colored = (m0.x & 0x10u) != 0;
        if (!(alpha == 0.0f && !colored)) {
          num_hits++;
          bool accepted = true;
          if (num_hits > 1) {
            const float prob  = 1.0f / num_hits;
            accepted          = random < prob;
            const float shift = accepted ? 0.0f : prob;
            const float scale = accepted ? prob : 1.0f - prob;
            random            = lbrng::saturate_random((random - shift) / scale);
          }
          if (accepted)
            selected = best_light;
          if (alpha == 1.0f)
            again = false;  // an opaque emitter culls everything behind it
        }
      }
      best_t     = FLT_MAX;
      best_light = LB_LIGHT_ID_INVALID;
    }
    if (!again) {
      P.eq_dir[k].w = __uint_as_float(selected);
      P.eq_hits[k]  = num_hits;
    }
    return again;
  }
};

template <bool kTex>
__global__ void __launch_bounds__(TRACE_THREADS, 6) k_trace_enum(Bvh8 light_bvh, LbPaths P, LbCounters* C, const uint32_t* __restrict__ light_prims,
                                                                  LbTraceTuning tune, LbTexScene T) {
  LbTraversalCount cnt;
  cnt.nodes = 0, cnt.tris = 0;
  const uint32_t n = C->n_enum;
  LbEnumPolicy<kTex> pol;
  pol.P           = P;
  pol.T           = T;
  pol.light_prims = light_prims;
  lb_trace_warp<LbEnumPolicy<kTex>, false>(light_bvh, n, &C->fetch_enum, pol, cnt, tune, &C->stack_overflow);
  if (blockIdx.x == 0 && threadIdx.x == 0)
    atomicAdd(&C->light_rays, (unsigned long long) n);
}

// ---------------------------------------------------------------------------------------------
// queue sort: counting sort of the active queue by material id (misses last)
// ---------------------------------------------------------------------------------------------
// sort_rank[material] = bin of the material: its rank in (class, material id) order (device_api.cu: upload_materials), so that the
// hits of one material class are contiguous after the sort. nullptr = unsorted mode: every hit goes to bin 0.
__device__ __forceinline__ uint32_t sort_key(uint32_t prim, const uint16_t* __restrict__ prim_material, const uint16_t* __restrict__ sort_rank) {
  if (prim == LB_HIT_SKY)
    return LB_SORT_KEY_SKY;
  if (!sort_rank)
    return 0;
  return __ldg(sort_rank + __ldg(prim_material + prim));
}

__global__ void __launch_bounds__(256) k_sort_count(LbPaths P, const uint32_t* __restrict__ queue, const LbCounters* C,
                                                    const uint16_t* __restrict__ prim_material, const uint16_t* __restrict__ by_material,
                                                    uint32_t* __restrict__ bins) {
  __shared__ uint32_t local[LB_SORT_BINS];
  for (uint32_t b = threadIdx.x; b < LB_SORT_BINS; b += blockDim.x)
    local[b] = 0;
  __syncthreads();
  const uint32_t n    = C->n_active;
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t k0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); k0 < n; k0 += gridDim.x * blockDim.x) {
    const uint32_t k     = k0 + lane;
    const bool valid     = k < n;
    const uint32_t key   = valid ? sort_key(P.prim[queue[k]], prim_material, by_material) : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
    if (valid && lane == (uint32_t) (__ffs(peers) - 1))
      atomicAdd(&local[key], (uint32_t) __popc(peers));
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < LB_SORT_BINS; b += blockDim.x)
    if (local[b])
      atomicAdd(&bins[b], local[b]);
}

// single block: exclusive scan of the bins into bins[LB_SORT_BINS ..], publishes n_hits, resets cursors
__global__ void __launch_bounds__(LB_SORT_BINS) k_sort_scan(uint32_t* __restrict__ bins, LbCounters* C, LbSortClasses classes) {
  __shared__ uint32_t tmp[LB_SORT_BINS];
  const uint32_t t = threadIdx.x;
  const uint32_t v = bins[t];
  bins[t]          = 0;  // ready for the next k_sort_count: the count bins are zero at allocation and after every scan (no clear launch)
  tmp[t]           = v;
  __syncthreads();
  for (uint32_t off = 1; off < LB_SORT_BINS; off <<= 1) {
    const uint32_t add = (t >= off) ? tmp[t - off] : 0u;
    __syncthreads();
    tmp[t] += add;
    __syncthreads();
  }
  const uint32_t excl     = tmp[t] - v;
  bins[LB_SORT_BINS + t]  = excl;  // running cursor per bin
  if (t == LB_SORT_KEY_SKY)
    C->n_hits = excl;
  // class ranges of the sorted queue: class c starts where its first bin starts
#pragma unroll
  for (int c = 0; c <= LB_NUM_CLASSES; c++)
    if (t == classes.first_rank[c])
      C->class_begin[c] = excl;
}

// Block-aggregated scatter: a block ranks one chunk of the queue in shared memory (one shared atomic per distinct key
// per warp via __match_any_sync), reserves its range of every non-empty bin with ONE global atomic per bin, then
// writes. 2M paths cost ~20k global atomics instead of 2M on a handful of hot addresses.
#define SORT_ITEMS 8
__global__ void __launch_bounds__(256) k_sort_scatter(LbPaths P, const uint32_t* __restrict__ queue_in, uint32_t* __restrict__ queue_out,
                                                      const LbCounters* C, const uint16_t* __restrict__ prim_material,
                                                      const uint16_t* __restrict__ by_material, uint32_t* __restrict__ bins) {
  __shared__ uint32_t local[LB_SORT_BINS];
  const uint32_t n     = C->n_active;
  const uint32_t chunk = 256u * SORT_ITEMS;
  const uint32_t lane  = threadIdx.x & 31u;
  for (uint32_t start = blockIdx.x * chunk; start < n; start += gridDim.x * chunk) {
    for (uint32_t b = threadIdx.x; b < LB_SORT_BINS; b += 256u)
      local[b] = 0;
    __syncthreads();
    uint32_t idx[SORT_ITEMS], key[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; j++) {
      const uint32_t k = start + j * 256u + threadIdx.x;
      const bool valid = k < n;
      idx[j]           = valid ? queue_in[k] : 0u;
      key[j]           = valid ? sort_key(P.prim[idx[j]], prim_material, by_material) : 0xFFFFFFFFu;
      const uint32_t peers  = __match_any_sync(0xFFFFFFFFu, key[j]);
      const uint32_t leader = __ffs(peers) - 1u;
      uint32_t base         = 0;
      if (valid && lane == leader)
        base = atomicAdd(&local[key[j]], (uint32_t) __popc(peers));
      base    = __shfl_sync(0xFFFFFFFFu, base, leader);
      rank[j] = base + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < LB_SORT_BINS; b += 256u) {
      const uint32_t c = local[b];
      if (c)
        local[b] = atomicAdd(&bins[LB_SORT_BINS + b], c);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; j++)
      if (key[j] != 0xFFFFFFFFu)
        queue_out[local[key[j]] + rank[j]] = idx[j];
    __syncthreads();
  }
}

// rotate counters between bounces: next queue becomes the active one
__global__ void k_next_bounce(LbCounters* C) {
  C->n_active = C->n_next;
  C->n_next   = 0;
  reset_bounce_counters(C);
}

// ---------------------------------------------------------------------------------------------
// explicit ray batches (C-ABI lumb200_device_trace_rays) and result extraction for the parity hooks
// ---------------------------------------------------------------------------------------------
__global__ void k_load_rays(LbPaths P, const float* __restrict__ origins, const float* __restrict__ dirs, uint32_t n, uint32_t* queue,
                            LbCounters* C) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    P.org[i]  = make_float4(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2], 0.0f);
    P.dir[i]  = make_float4(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], FLT_MAX);
    P.prim[i] = LB_PRIM_NONE;
    queue[i]  = i;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    C->n_active = n;
    C->n_next   = 0;
    reset_fetch_cursors(C);
  }
}

__global__ void k_extract_hits(LbPaths P, const uint2* __restrict__ prim_handle, const float2* __restrict__ uv, uint32_t n,
                               uint32_t* __restrict__ inst_out, uint32_t* __restrict__ tri_out, float* __restrict__ t_out, float* __restrict__ u_out,
                               float* __restrict__ v_out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t prim = P.prim[i];
    uint2 h             = make_uint2(LB_HIT_SKY, 0u);
    if (prim != LB_HIT_SKY)
      h = prim_handle[prim];
    inst_out[i] = h.x;
    tri_out[i]  = h.y;
    t_out[i]    = P.dir[i].w;
    u_out[i]    = uv[i].x;
    v_out[i]    = uv[i].y;
  }
}

// ---------------------------------------------------------------------------------------------
// per-vertex / per-ray parity hooks (C-ABI lumb200_device_shade_vertices, lumb200_device_trace_shadow_rays): the surface stages
// run on caller-supplied vertices instead of the output of the closest-hit stage
// ---------------------------------------------------------------------------------------------
__global__ void k_load_vertices(LbPaths P, const Lumb200VertexIn* __restrict__ in, uint32_t n, uint32_t width, uint32_t sample_id,
                                uint32_t* __restrict__ queue, LbCounters* C) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Lumb200VertexIn v = in[i];
    P.org[i]       = make_float4(v.origin[0], v.origin[1], v.origin[2], 0.0f);
    P.dir[i]       = make_float4(v.ray[0], v.ray[1], v.ray[2], v.t);
    P.prim[i]      = (v.prim == 0xFFFFFFFFu) ? LB_HIT_SKY : v.prim;  // a miss: sorted to the tail, shaded by the miss kernel
    P.record[i]    = make_uint2(v.record[0], v.record[1]);
    P.pixel[i]     = v.pixel_x + v.pixel_y * width;
    P.state[i]     = v.state;
    P.medium[i]    = v.medium;
    P.sample_id[i] = sample_id;
    P.result[i]    = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int slot = 0; slot < LB_NEE_SLOTS; slot++)
      P.nee[LB_NEE_SLOTS * (size_t) i + slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    queue[i]       = i;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    C->n_active = n;
    C->n_next   = 0;
    reset_bounce_counters(C);
  }
}

// shadow-queue entries -> the segment records of their vertices (before k_trace_shadow adds the visible part)
__global__ void k_extract_segments(LbPaths P, const LbCounters* C, Lumb200VertexOut* __restrict__ out) {
  const uint32_t n0 = C->n_shadow[0], n01 = n0 + C->n_shadow[1], n012 = n01 + C->n_shadow[2], n = n012 + C->n_shadow[3];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint32_t k = lb_shadow_queue_entry(j, n0, n01, n012, P.capacity);
    const float4 o = P.sq_org[k], d = P.sq_dir[k], c = P.sq_col[k];
    const uint32_t tag = __float_as_uint(o.w);
    Lumb200NeeSegment& s = out[tag & 0x3FFFFFFFu].nee[tag >> 30];
    s.valid = 1;
    s.ray[0] = d.x, s.ray[1] = d.y, s.ray[2] = d.z, s.dist = d.w;
    s.color[0] = c.x, s.color[1] = c.y, s.color[2] = c.z;
    s.target_prim = __float_as_uint(c.w);
  }
}

__global__ void k_extract_vertices(LbPaths P, uint32_t n, const uint32_t* __restrict__ queue_out, const LbCounters* C,
                                   Lumb200VertexOut* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Lumb200VertexOut& v = out[i];
    const float4 r      = P.result[i];
    v.emission[0] = r.x, v.emission[1] = r.y, v.emission[2] = r.z;
#pragma unroll
    for (int s = 0; s < LB_NEE_SLOTS; s++) {
      const float4 a = P.nee[LB_NEE_SLOTS * (size_t) i + s];
      v.nee[s].visible[0] = a.x, v.nee[s].visible[1] = a.y, v.nee[s].visible[2] = a.z;
    }
  }
  const uint32_t alive = C->n_next;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < alive; k += gridDim.x * blockDim.x) {
    const uint32_t i    = queue_out[k];
    Lumb200VertexOut& v = out[i];
    const float4 o = P.org[i], d = P.dir[i];
    const uint2 rec = P.record[i];
    v.alive = 1;
    v.state = P.state[i];
    v.origin[0] = o.x, v.origin[1] = o.y, v.origin[2] = o.z;
    v.ray[0] = d.x, v.ray[1] = d.y, v.ray[2] = d.z;
    v.record[0] = rec.x, v.record[1] = rec.y;
    v.medium = P.medium[i];
  }
}

__global__ void k_load_shadow_rays(LbPaths P, const float* __restrict__ origins, const float* __restrict__ dirs, const float* __restrict__ max_dist,
                                   const uint32_t* __restrict__ ignore_prims, const uint32_t* __restrict__ target_prims, uint32_t n, LbCounters* C) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    P.prim[i]   = ignore_prims[i];
    P.sq_org[i] = make_float4(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2], __uint_as_float(i));  // path i, NEE slot 0
    P.sq_dir[i] = make_float4(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2], max_dist[i]);
    P.sq_col[i] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(target_prims[i]));
    P.nee[LB_NEE_SLOTS * (size_t) i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    C->n_shadow[0] = n;
    C->n_shadow[1] = C->n_shadow[2] = C->n_shadow[3] = 0;
    reset_fetch_cursors(C);
  }
}

__global__ void k_extract_visibility(LbPaths P, uint32_t n, float* __restrict__ out) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 a = P.nee[LB_NEE_SLOTS * (size_t) i];
    out[3 * i + 0] = a.x, out[3 * i + 1] = a.y, out[3 * i + 2] = a.z;
  }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
extern "C++" {

void lb_launch_raygen(const LbPaths& P, const LbFrame& F, const LbCameraDev& cam, const uint32_t* bluenoise, uint32_t sample_id, uint32_t* queue,
                      LbCounters* C, int grid, cudaStream_t s, bool tile_order) {
  k_raygen<<<grid, 256, 0, s>>>(P, F, cam, bluenoise, sample_id, queue, C, tile_order);
}

static LbTraceTuning tuning() {
  // LUMB200_FETCH_THRESHOLD / LUMB200_TRI_THRESHOLD override the defaults (parameter sweeps only)
  static LbTraceTuning t = [] {
    LbTraceTuning v = {LB_FETCH_THRESHOLD_DEFAULT, LB_TRI_THRESHOLD_DEFAULT, LB_NODE_BIAS_BITS};
    if (const char* e = getenv("LUMB200_FETCH_THRESHOLD"))
      v.fetch_threshold = (uint32_t) atoi(e);
    if (const char* e = getenv("LUMB200_TRI_THRESHOLD"))
      v.tri_threshold = (uint32_t) atoi(e);
    return v;
  }();
  return t;
}

// tex == nullptr: the scene has no albedo textures, the untextured kernel variants run
void lb_launch_trace_closest(const Bvh8& bvh, const LbPaths& P, const uint32_t* queue, LbCounters* C, float2* uv, int grid, cudaStream_t s,
                             bool count, const LbTexScene* tex) {
  const LbTexScene T = tex ? *tex : LbTexScene{};
  if (tex) {
    if (count)
      k_trace_closest<true, true><<<grid, TRACE_THREADS, 0, s>>>(bvh, P, queue, C, uv, tuning(), T);
    else
      k_trace_closest<false, true><<<grid, TRACE_THREADS, 0, s>>>(bvh, P, queue, C, uv, tuning(), T);
  }
  else if (count)
    k_trace_closest<true, false><<<grid, TRACE_THREADS, 0, s>>>(bvh, P, queue, C, uv, tuning(), T);
  else
    k_trace_closest<false, false><<<grid, TRACE_THREADS, 0, s>>>(bvh, P, queue, C, uv, tuning(), T);
}

void lb_launch_trace_shadow(const Bvh8& bvh, const LbPaths& P, LbCounters* C, const uint16_t* prim_material, const float4* shadow_tab, int grid,
                            cudaStream_t s, bool count, const LbTexScene* tex) {
  const LbTexScene T = tex ? *tex : LbTexScene{};
  if (tex) {
    if (count)
      k_trace_shadow<true, true><<<grid, TRACE_THREADS, 0, s>>>(bvh, P, C, prim_material, shadow_tab, tuning(), T);
    else
      k_trace_shadow<false, true><<<grid, TRACE_THREADS, 0, s>>>(bvh, P, C, prim_material, shadow_tab, tuning(), T);
  }
  else if (count)
    k_trace_shadow<true, false><<<grid, TRACE_THREADS, 0, s>>>(bvh, P, C, prim_material, shadow_tab, tuning(), T);
  else
    k_trace_shadow<false, false><<<grid, TRACE_THREADS, 0, s>>>(bvh, P, C, prim_material, shadow_tab, tuning(), T);
}

// sort_rank == nullptr: unsorted mode (hits / misses only); `classes` must then put every hit into the GENERIC class
// tex == nullptr: no albedo-textured material; the enumeration still needs the material table (emitter alpha)
void lb_launch_trace_enum(const Bvh8& light_bvh, const LbPaths& P, LbCounters* C, const uint32_t* light_prims, const LbTexScene& T, bool textured,
                          int grid, cudaStream_t s) {
  if (textured)
    k_trace_enum<true><<<grid, TRACE_THREADS, 0, s>>>(light_bvh, P, C, light_prims, tuning(), T);
  else
    k_trace_enum<false><<<grid, TRACE_THREADS, 0, s>>>(light_bvh, P, C, light_prims, tuning(), T);
}

void lb_launch_sort(const LbPaths& P, const uint32_t* queue_in, uint32_t* queue_out, LbCounters* C, const uint16_t* prim_material,
                    const uint16_t* sort_rank, const LbSortClasses& classes, uint32_t* bins, int grid, cudaStream_t s) {
  k_sort_count<<<grid, 256, 0, s>>>(P, queue_in, C, prim_material, sort_rank, bins);
  k_sort_scan<<<1, LB_SORT_BINS, 0, s>>>(bins, C, classes);
  k_sort_scatter<<<grid, 256, 0, s>>>(P, queue_in, queue_out, C, prim_material, sort_rank, bins);
}

void lb_launch_next_bounce(LbCounters* C, cudaStream_t s) { k_next_bounce<<<1, 1, 0, s>>>(C); }

void lb_launch_load_rays(const LbPaths& P, const float* origins, const float* dirs, uint32_t n, uint32_t* queue, LbCounters* C, int grid,
                         cudaStream_t s) {
  k_load_rays<<<grid, 256, 0, s>>>(P, origins, dirs, n, queue, C);
}

void lb_launch_raygen_pixel(const LbPaths& P, const LbFrame& F, const LbCameraDev& cam, const uint32_t* bluenoise, uint32_t x, uint32_t y,
                            uint32_t sample_id, uint32_t* queue, LbCounters* C, cudaStream_t s) {
  k_raygen_pixel<<<1, 1, 0, s>>>(P, F, cam, bluenoise, x, y, sample_id, queue, C);
}

void lb_launch_load_vertices(const LbPaths& P, const Lumb200VertexIn* in, uint32_t n, uint32_t width, uint32_t sample_id, uint32_t* queue,
                             LbCounters* C, int grid, cudaStream_t s) {
  k_load_vertices<<<grid, 256, 0, s>>>(P, in, n, width, sample_id, queue, C);
}

void lb_launch_extract_segments(const LbPaths& P, const LbCounters* C, Lumb200VertexOut* out, int grid, cudaStream_t s) {
  k_extract_segments<<<grid, 256, 0, s>>>(P, C, out);
}

void lb_launch_extract_vertices(const LbPaths& P, uint32_t n, const uint32_t* queue_out, const LbCounters* C, Lumb200VertexOut* out, int grid,
                                cudaStream_t s) {
  k_extract_vertices<<<grid, 256, 0, s>>>(P, n, queue_out, C, out);
}

void lb_launch_load_shadow_rays(const LbPaths& P, const float* origins, const float* dirs, const float* max_dist, const uint32_t* ignore_prims,
                                const uint32_t* target_prims, uint32_t n, LbCounters* C, int grid, cudaStream_t s) {
  k_load_shadow_rays<<<grid, 256, 0, s>>>(P, origins, dirs, max_dist, ignore_prims, target_prims, n, C);
}

void lb_launch_extract_visibility(const LbPaths& P, uint32_t n, float* out, int grid, cudaStream_t s) {
  k_extract_visibility<<<grid, 256, 0, s>>>(P, n, out);
}

void lb_launch_extract_hits(const LbPaths& P, const uint2* prim_handle, const float2* uv, uint32_t n, uint32_t* inst, uint32_t* tri, float* t,
                            float* u, float* v, int grid, cudaStream_t s) {
  k_extract_hits<<<grid, 256, 0, s>>>(P, prim_handle, uv, n, inst, tri, t, u, v);
}
}
