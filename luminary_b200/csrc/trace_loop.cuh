// trace_loop.cuh - warp-synchronous persistent traversal loop with per-lane ray replacement and triangle postponing.
//
// Replaces the closed-source traversal behind the reference's two OptiX launches per bounce
// (optix/optix_kernel_raytrace.cu:147-183, optix/optix_kernel_shadow.cu:15-100). B200 has no RT cores, so the loop is
// organised around what limits a software traversal on SIMT hardware - lane utilisation:
//   * every lane owns one ray and an explicit state machine (node group, pending triangle group, stack);
//   * the warp votes (__ballot_sync) on what to do next: one BVH8 node step or one triangle test. Triangle groups are
//     POSTPONED (kept in the lane, older ones pushed to the stack) until enough lanes have one, so that the triangle
//     test executes with many lanes active instead of running a 3-lane loop after every node step;
//   * lanes whose ray has terminated stay idle only until the number of active lanes drops below a threshold; then
//     the warp fetches replacement rays with ONE atomic (ballot + popc ranks), Aila/Laine-style dynamic fetch.
// Results are independent of the traversal order: closest hits resolve equal-t ties to the smaller primitive index
// and box culling is conservative, so ids / t / u / v stay bit-identical to the oracle.
#pragma once

#include "traverse.cuh"

// Tunables (warp-uniform kernel arguments; defaults chosen from sweeps on B200, see profiles/):
struct LbTraceTuning {
  uint32_t fetch_threshold;  // fetch replacement rays once <= this many lanes are active
  uint32_t tri_threshold;    // run a triangle step once >= this many lanes hold a pending triangle
  uint32_t one_bits;         // LB_NODE_BIAS_BITS, passed as data so that it lives in a register (see lb_u8_biased)
};
#define LB_FETCH_THRESHOLD_DEFAULT 22
#define LB_TRI_THRESHOLD_DEFAULT 8
#ifndef LB_LOOP_STACK
#define LB_LOOP_STACK 32
#endif
// Short stack: the first LB_SMEM_STACK entries of every lane's stack live in shared memory ([entry][thread]: a warp's accesses to
// one entry are 256 contiguous bytes, conflict-free however the lanes' stack pointers differ), deeper entries in local memory.
// 0 keeps the whole stack in local memory (round 1). Measured on B200: profiles/r2_variants.md.
#ifndef LB_SMEM_STACK
#define LB_SMEM_STACK 0
#endif
#ifndef LB_TRACE_THREADS
#define LB_TRACE_THREADS 128
#endif
#ifndef LB_STACK_TOP_REG
#define LB_STACK_TOP_REG 0
#endif
// 1: after every fetch the warp looks at the 32 queue entries behind its batch (Policy::peek) and, one loop iteration later, issues L2
// prefetches for their ray records (Policy::prefetch)
#ifndef LB_TRACE_PREFETCH
#define LB_TRACE_PREFETCH 0
#endif

// Policy interface:
//   void begin(uint32_t k, LbRay& r)                       load ray k of the queue, reset the per-ray result
//   bool hit(uint32_t prim, float t, float u, float v, float& tmax)   as the visitors of traverse.cuh; true = terminate
//   bool end()                                             the traversal of the ray is complete: write its result and return false, or
//                                                          return true to traverse the SAME ray again from the root with r.tmax restored
//                                                          (emitter enumeration: one closest-hit query per emitter, k_trace_enum)
// The stack holds sibling node groups AND postponed triangle groups: one descent pushes at most two entries per BVH8 level, so a tree of
// depth D needs at most 2 D entries. lumb200_device_build_accel refuses trees with 2 D > LB_LOOP_STACK (the level-synchronous collapse
// knows D); should an entry ever not fit all the same, the drop is COUNTED in *overflow (LbCounters.stack_overflow, Lumb200Stats) -
// tests and bench.py assert that it stays 0 - instead of silently losing a subtree.
template <typename Policy, bool kCount>
__device__ __forceinline__ void lb_trace_warp(const Bvh8& bvh, const uint32_t n, uint32_t* fetch_cursor, Policy& pol, LbTraversalCount& cnt,
                                              const LbTraceTuning tune, uint32_t* overflow) {
  const uint32_t FULL    = 0xFFFFFFFFu;
  const uint32_t lane    = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;

  LbRay r;
  r.ox = r.oy = r.oz = 0.0f, r.dx = r.dy = 0.0f, r.dz = 1.0f, r.tmin = 0.0f, r.tmax = 0.0f;
  float idx = 0.0f, idy = 0.0f, idz = 0.0f, tmax = 0.0f;
  uint32_t octinv = 0;
  LbShear shear;
  shear.Sx = shear.Sy = shear.Sz = 0.0f, shear.kz = 2, shear.swap = false;
  uint2 group = make_uint2(0u, 0u);  // inner-node group: x = child base, y = hit bits (31..24) | imask (7..0)
  uint2 tris  = make_uint2(0u, 0u);  // pending triangle group: x = triangle base, y = 24 hit bits
  int sp = 0;
#if LB_SMEM_STACK > 0
  __shared__ uint2 s_stack[LB_SMEM_STACK][LB_TRACE_THREADS];
  uint2 stack[LB_LOOP_STACK - LB_SMEM_STACK];
#define LB_STACK_PUSH(e)                                   \
  do {                                                     \
    if (sp < LB_SMEM_STACK)                                \
      s_stack[sp][threadIdx.x] = (e);                      \
    else                                                   \
      stack[sp - LB_SMEM_STACK] = (e);                     \
    sp++;                                                  \
  } while (0)
#define LB_STACK_POP() ((--sp < LB_SMEM_STACK) ? s_stack[sp][threadIdx.x] : stack[sp - LB_SMEM_STACK])
#elif LB_STACK_TOP_REG
  // the top of the stack lives in registers: a pop returns it at once and issues the load of the entry below, whose latency then
  // overlaps the next node / triangle step instead of stalling the refill (ncu: 14.6 % of k_trace_closest's samples sit there)
  uint2 stack[LB_LOOP_STACK];
  uint2 top = make_uint2(0u, 0u);
#define LB_STACK_PUSH(e)      \
  do {                        \
    if (sp > 0)               \
      stack[sp - 1] = top;    \
    top = (e);                \
    sp++;                     \
  } while (0)
  auto lb_stack_pop = [&]() {
    const uint2 e = top;
    sp--;
    if (sp > 0)
      top = stack[sp - 1];
    return e;
  };
#define LB_STACK_POP() lb_stack_pop()
#else
  uint2 stack[LB_LOOP_STACK];
#define LB_STACK_PUSH(e) (stack[sp++] = (e))
#define LB_STACK_POP() (stack[--sp])
#endif
  bool active    = false;
  bool exhausted = false;
#if LB_TRACE_PREFETCH
  uint32_t pf = 0xFFFFFFFFu;
#endif

  for (;;) {
    // ---------------- dynamic fetch ----------------
    uint32_t act_mask = __ballot_sync(FULL, active);
    if (!exhausted && (uint32_t) __popc(act_mask) <= tune.fetch_threshold) {
      const uint32_t idle = ~act_mask;
      const uint32_t want = __popc(idle);
      uint32_t base       = 0;
      if (lane == 0)
        base = atomicAdd(fetch_cursor, want);
      base = __shfl_sync(FULL, base, 0);
      if (!active) {
        const uint32_t k = base + __popc(idle & lt_mask);
        if (k < n) {
          pol.begin(k, r);
          const float tiny = 8.271806125530277e-25f;  // 2^-80
          idx    = 1.0f / ((fabsf(r.dx) > tiny) ? r.dx : copysignf(tiny, r.dx));
          idy    = 1.0f / ((fabsf(r.dy) > tiny) ? r.dy : copysignf(tiny, r.dy));
          idz    = 1.0f / ((fabsf(r.dz) > tiny) ? r.dz : copysignf(tiny, r.dz));
          octinv = ((r.dx >= 0.0f) ? 1u : 0u) | ((r.dy >= 0.0f) ? 2u : 0u) | ((r.dz >= 0.0f) ? 4u : 0u);
          shear  = lb_shear(r);
          tmax   = r.tmax;
          group  = make_uint2(0u, 0x01000000u);  // selects node 0: imask 0 gives relative index 0 for any slot
          tris   = make_uint2(0u, 0u);
          sp     = 0;
          active = true;
        }
      }
      if (base + want >= n)
        exhausted = true;
#if LB_TRACE_PREFETCH
      {  // the rays behind this batch are fetched next (by whichever warp runs dry first): pull them towards the L2 now
        const uint32_t kp = base + want + lane;
        pf                = (kp < n) ? pol.peek(kp) : 0xFFFFFFFFu;
      }
#endif
      act_mask = __ballot_sync(FULL, active);
    }
#if LB_TRACE_PREFETCH
    else if (pf != 0xFFFFFFFFu) {  // one iteration after the peek: its load has returned, the prefetches cost no stall
      pol.prefetch(pf);
      pf = 0xFFFFFFFFu;
    }
#endif
    if (act_mask == 0)
      break;

    // ---------------- vote: node step or triangle step ----------------
    const bool has_node   = active && (group.y & 0xFF000000u);
    const bool has_tri    = active && (tris.y != 0u);
    const uint32_t m_node = __ballot_sync(FULL, has_node);
    const uint32_t m_tri  = __ballot_sync(FULL, has_tri);

    if (m_node != 0 && (uint32_t) __popc(m_tri) < tune.tri_threshold) {
      if (has_node) {
        const uint32_t hits = group.y;
        const uint32_t bit  = 31u - __clz(hits);
        group.y &= ~(1u << bit);
        if (group.y & 0xFF000000u) {
          if (sp < LB_LOOP_STACK)
            LB_STACK_PUSH(group);
          else
            atomicAdd(overflow, 1u);
        }
        const uint32_t slot       = (bit - 24u) ^ octinv;
        const uint32_t imask      = hits & 0xFFu;
        const uint32_t rel        = __popc(imask & ~(0xFFFFFFFFu << slot));
        const uint32_t node_index = group.x + rel;

        const uint4* np = bvh.nodes + 5 * (size_t) node_index;
        const uint4 n0  = __ldg(np + 0);
        const uint4 n1  = __ldg(np + 1);
        const uint4 n2  = __ldg(np + 2);
        const uint4 n3  = __ldg(np + 3);
        const uint4 n4  = __ldg(np + 4);

        const uint32_t hitmask = lb_node_hits(n0, n1, n2, n3, n4, r, idx, idy, idz, octinv * 0x01010101u, tmax, tune.one_bits);
        if (kCount)
          cnt.nodes++;

        group.x = n1.x;
        group.y = (hitmask & 0xFF000000u) | (n0.w >> 24);
        if (hitmask & 0x00FFFFFFu) {
          if (tris.y != 0u) {  // postpone the older triangle group
            if (sp < LB_LOOP_STACK)
              LB_STACK_PUSH(tris);
            else
              atomicAdd(overflow, 1u);
          }
          tris.x = n1.y;
          tris.y = hitmask & 0x00FFFFFFu;
        }
      }
    }
    else {
      if (has_tri) {
        const uint32_t i = __ffs(tris.y) - 1u;
        tris.y &= tris.y - 1u;
        const float4* tp = bvh.tris + 3 * (size_t) (tris.x + i);
        const float4 v0  = __ldg(tp + 0);
        const float4 v1  = __ldg(tp + 1);
        const float4 v2  = __ldg(tp + 2);
        if (kCount)
          cnt.tris++;
        float t, u, v;
        if (lb_tri_watertight(r, shear, v0, v1, v2, t, u, v)) {
          if (t >= r.tmin && t <= tmax) {
            if (pol.hit(__float_as_uint(v0.w), t, u, v, tmax)) {
              if (pol.end()) {
                tmax  = r.tmax;
                group = make_uint2(0u, 0x01000000u);
                tris  = make_uint2(0u, 0u);
                sp    = 0;
              }
              else
                active = false;
            }
          }
        }
      }
    }

    // ---------------- refill from the stack / terminate ----------------
    if (active && (group.y & 0xFF000000u) == 0u && tris.y == 0u) {
      if (sp == 0) {
        if (pol.end()) {
          tmax  = r.tmax;
          group = make_uint2(0u, 0x01000000u);
          tris  = make_uint2(0u, 0u);
        }
        else
          active = false;
      }
      else {
        const uint2 e = LB_STACK_POP();
        if (e.y & 0xFF000000u)
          group = e;
        else
          tris = e;
      }
    }
  }
}
#undef LB_STACK_PUSH
#undef LB_STACK_POP
