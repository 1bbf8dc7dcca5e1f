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

// Material classes of the sorted hit queue. The counting sort keys a hit by the RANK of its material in (class, material id)
// order, so every class is one contiguous range of the sorted queue and is shaded by a kernel instantiation that only carries the
// code of that class (north_star: "rays compacted and sorted by material ... GGX / transmission / emission shading kernels"):
//   GENERIC     anything: translucent substrates (refraction lobe, medium stack), opacity < 1 or alpha from an albedo texture
//   DIELECTRIC  opaque substrate, not metallic, opacity 1: diffuse + glossy lobes, no refraction / Fresnel-transmission code
//   METAL       opaque substrate, metallic, opacity 1: conductor lobe only
// Scenes with more materials than sort bins, and the unsorted mode (sort_by_material = 0), run everything as GENERIC.
#define LB_CLASS_GENERIC 0
#define LB_CLASS_DIELECTRIC 1
#define LB_CLASS_METAL 2
#define LB_NUM_CLASSES 3

// NEE slots of a path vertex: every slot owns one shadow ray per bounce, one region of the shadow-ray queue and one accumulator
//   0 light-tree light, 1 BSDF-sampled light, 2 ambient (constant-colour sky), 3 sun (procedural sky)
#define LB_NEE_SLOTS 4
#define LB_NEE_SLOT_SUN 3

struct LbSortClasses {
  uint32_t first_rank[LB_NUM_CLASSES + 1];  // first sort bin of every class; [LB_NUM_CLASSES] = LB_SORT_KEY_SKY
};

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
  float4* nee;       // [LB_NEE_SLOTS * capacity] radiance gathered through the NEE slots; one shadow ray per slot and bounce adds to
                     // its own accumulator, so the sum is deterministic without atomics; folded in by k_accumulate
  // shadow-ray queues of the current bounce: one region of `capacity` entries per NEE slot (light-tree light, BSDF-sampled
  // light, ambient, sun), region s starts at s * capacity and holds n_shadow[s] entries. k_trace_shadow walks the regions back
  // to back, so a warp traces rays of ONE kind (bounded segments towards emitters / unbounded ambient rays) from neighbouring paths.
  float4* sq_org;    // xyz origin (raw hit point), w = path slot | NEE slot << 30 (bits)
  float4* sq_dir;    // xyz direction, w = max distance
  float4* sq_col;    // rgb contribution (already multiplied by the throughput), w = target light prim (bits)
  // emitter-enumeration queue of the BSDF-sampled NEE direction (direct_lighting.cuh:601-669), [capacity], n_enum entries:
  // written by k_shade, traced against the emitter BVH by k_trace_enum, turned into slot-1 shadow segments by k_enum_finish
  float4* eq_org;     // xyz raw hit point, w = path slot (bits)
  float4* eq_dir;     // xyz direction, w = the any-hit reservoir's random number; k_trace_enum overwrites w with the selected light id (bits)
  float4* eq_weight;  // rgb BSDF weight of the direction, w = its sampling probability
  float4* eq_rec;     // rgb path throughput, w = light_tree_root_sum
  uint32_t* eq_hits;  // emitters counted along the ray (written by k_trace_enum)
  uint32_t capacity;
};

// device-side counters of one pass
struct LbCounters {
  uint32_t n_active;      // entries in the current queue
  uint32_t n_next;        // entries appended to the next queue
  uint32_t fetch;         // work fetch cursor of k_trace_closest (one cursor per persistent kernel of a bounce: all are reset together by
                          // the kernel that ends the previous bounce instead of by launches of their own)
  uint32_t n_hits;        // after sorting: queue[0 .. n_hits) are surface hits, the rest misses
  unsigned long long closest_rays;
  unsigned long long shadow_rays;
  unsigned long long light_rays;
  uint32_t stack_overflow;
  uint32_t n_shadow[LB_NEE_SLOTS];  // entries in the shadow-ray queue regions of this bounce
  uint32_t n_enum;        // entries in the emitter-enumeration queue of this bounce
  uint32_t class_begin[LB_NUM_CLASSES + 1];  // after sorting: queue[class_begin[c] .. class_begin[c + 1]) are the hits of class c
  // filled by the instrumented kernel variants only (lumb200_device_measure_traversal)
  unsigned long long closest_nodes, closest_tris, shadow_nodes, shadow_tris;
  unsigned long long light_tree_nodes, shaded_vertices;
  uint32_t nonfinite_samples;  // path samples whose radiance was NaN / Inf and was dropped by the accumulation kernels
  uint32_t nonfinite_pixel;    // pixel index of the last one (diagnostics)
  uint32_t fetch_enum;         // work fetch cursor of k_trace_enum
  uint32_t fetch_shadow;       // work fetch cursor of k_trace_shadow
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
