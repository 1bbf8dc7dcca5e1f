// wavefront.cuh - per-path state of the wavefront (struct of arrays, 16-byte lanes) and launch parameters.
//
// The reference keeps 80-byte DeviceTaskState records per thread in a warp-interleaved AoSoA
// (device_utils.h:359-410, cuda/memory.cuh:114-197). Here every path of a sample pass owns one slot `i` in
// global SoA arrays, and stages communicate through index queues (compaction instead of per-thread lists).
#pragma once

#include "lumb200_internal.cuh"

// StateFlag, reference cuda/utils.cuh:113-120
#define LB_STATE_DELTA_PATH 0x01u
#define LB_STATE_CAMERA_DIRECTION 0x02u
#define LB_STATE_VOLUME_SCATTERED 0x04u
#define LB_STATE_ALLOW_EMISSION 0x08u
#define LB_STATE_ALLOW_AMBIENT 0x10u
#define LB_STATE_USE_IGNORE_HANDLE 0x20u

#define LB_SORT_BINS 1024u
#define LB_SORT_KEY_SKY (LB_SORT_BINS - 1u)

struct LbPaths {
  float4* org;       // xyz origin of the current ray
  float4* dir;       // xyz direction, w = hit distance written by the closest-hit kernel
  uint32_t* prim;    // in: primitive to ignore (LB_PRIM_NONE for none); out: closest primitive or LB_HIT_SKY
  uint2* record;     // throughput, 3 x 21-bit floats (record_pack, reference cuda/math.cuh:1609-1619)
  uint32_t* pixel;   // pixel index x + y * width
  uint32_t* state;   // state flags
  uint32_t* medium;  // IOR stack (DeviceTaskMediumStack.ior, device_utils.h:383-389)
  float4* result;    // radiance gathered by this path during the pass (emission, sky)
  uint32_t* sample_id;  // per path, written by k_raygen_adaptive only (adaptive executions mix sample ids in one launch)
  float4* nee;       // [3 * capacity] radiance gathered through the three NEE slots; one shadow ray per slot and bounce adds to
                     // its own accumulator, so the sum is deterministic without atomics; folded in by k_accumulate
  // shadow-ray queue of the current bounce, [3 * capacity], appended by k_shade (n_shadow entries)
  float4* sq_org;    // xyz origin (raw hit point), w = path slot | NEE slot << 30 (bits)
  float4* sq_dir;    // xyz direction, w = max distance
  float4* sq_col;    // rgb contribution (already multiplied by the throughput), w = target light prim (bits)
  uint32_t capacity;
};

// device-side counters of one pass
struct LbCounters {
  uint32_t n_active;      // entries in the current queue
  uint32_t n_next;        // entries appended to the next queue
  uint32_t fetch;         // work fetch cursor of the persistent kernels
  uint32_t n_hits;        // after sorting: queue[0 .. n_hits) are surface hits, the rest misses
  unsigned long long closest_rays;
  unsigned long long shadow_rays;
  unsigned long long light_rays;
  uint32_t stack_overflow;
  uint32_t n_shadow;      // entries in the shadow-ray queue of this bounce
  // filled by the instrumented kernel variants only (lumb200_device_measure_traversal)
  unsigned long long closest_nodes, closest_tris, shadow_nodes, shadow_tris;
};

// adaptive sampler state as the kernels see it (DeviceSampleAllocation + adaptive_sampling_accumulated_stages, device_utils.h:333-338,527)
#define LB_ADAPTIVE_STAGES 4
struct LbAdaptive {
  const uint32_t* words;  // per 4 x 4 block: byte s = samples per pixel and execution of stage s + 1, minus one
  uint32_t bw;            // blocks per row
  uint32_t executions[LB_ADAPTIVE_STAGES + 1];  // finished executions per stage
};

__device__ __forceinline__ uint32_t as_stage_count(uint32_t word, uint32_t stage) { return ((word >> (stage * 8u)) & 0xFFu) + 1u; }

// adaptive_sampling_get_sample_count / adapative_sampling_get_sample_offset: samples every pixel of the block has received
__device__ __forceinline__ uint32_t as_block_samples(uint32_t word, const LbAdaptive& A) {
  uint32_t count = A.executions[0];
#pragma unroll
  for (uint32_t st = 0; st < LB_ADAPTIVE_STAGES; st++)
    count += A.executions[st + 1] * as_stage_count(word, st);
  return min(count, 1u << 20);
}


struct LbCameraDev {  // DeviceCamera thin-lens subset, device_structs.h:38-83
  float px, py, pz;
  float qx, qy, qz, qw;
  float fov, aperture_size, object_distance, camera_scale, rr_threshold;
  uint32_t aperture_shape, aperture_blade_count;
};

struct LbFrame {
  uint32_t width, height;
  uint32_t max_depth;
  uint32_t sky_mode;
  float sky_r, sky_g, sky_b;
};
