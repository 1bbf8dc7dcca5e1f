// traverse.cuh - stack-based traversal of the compressed 8-wide BVH with a watertight triangle test.
//
// Replaces the closed-source OptiX traversal behind the reference's optixTrace call sites
// (cuda/optix_utils.cuh:43-94). Node format and octant-ordered traversal after Ylitie, Karras, Laine 2017;
// triangle test after Woop, Benthin, Wald, "Watertight Ray/Triangle Intersection", JCGT 2013, evaluated in
// exactly the operation order of the CPU oracle (oracle/orc_trace.c: tri_wt) - the including translation
// unit is compiled with -fmad=false so results are bit-identical. Equal-t ties resolve to the smaller
// flattened primitive index, which makes closest-hit ids independent of traversal order.
#pragma once

#include "lumb200_internal.cuh"

#define LB_STACK_SIZE 40

struct LbRay {
  float ox, oy, oz;
  float dx, dy, dz;
  float tmin, tmax;
};

struct LbHit {
  uint32_t prim;  // flattened primitive index, LB_HIT_SKY on miss
  float t, u, v;
};

// Ray constants of the watertight test.
struct LbShear {
  float Sx, Sy, Sz;
  int kz;      // dominant axis
  bool swap;   // dir[kz] < 0: swap kx, ky to preserve winding
};

__device__ __forceinline__ LbShear lb_shear(const LbRay& r) {
  LbShear s;
  const float ax = fabsf(r.dx), ay = fabsf(r.dy), az = fabsf(r.dz);
  s.kz = (ax >= ay && ax >= az) ? 0 : ((ay >= az) ? 1 : 2);
  float dkx, dky, dkz;
  if (s.kz == 0) {
    dkx = r.dy, dky = r.dz, dkz = r.dx;
  }
  else if (s.kz == 1) {
    dkx = r.dz, dky = r.dx, dkz = r.dy;
  }
  else {
    dkx = r.dx, dky = r.dy, dkz = r.dz;
  }
  s.swap = dkz < 0.0f;
  if (s.swap) {
    const float tmp = dkx;
    dkx             = dky;
    dky             = tmp;
  }
  s.Sx = dkx / dkz;
  s.Sy = dky / dkz;
  s.Sz = 1.0f / dkz;
  return s;
}

// Axis permutation (kx, ky, kz) of the watertight test, written as selects: the dominant axis differs from lane to
// lane, and the branchy form made the triangle test run with ~4 of 32 lanes active (ncu source page, round 1).
// Selection is exact, so the result is bit-identical to the oracle's indexed form.
__device__ __forceinline__ void lb_permute(const LbShear& s, float x, float y, float z, float& px, float& py, float& pz) {
  const bool k0 = s.kz == 0, k1 = s.kz == 1;
  const float a = k0 ? y : (k1 ? z : x);
  const float b = k0 ? z : (k1 ? x : y);
  pz            = k0 ? x : (k1 ? y : z);
  px            = s.swap ? b : a;
  py            = s.swap ? a : b;
}

// Returns true and (t, u, v) when the supporting plane is hit inside the triangle. The caller applies the
// t-interval. u, v weight v1, v2 (reference convention: coords.x * edge1 + coords.y * edge2).
__device__ __forceinline__ bool lb_tri_watertight(const LbRay& r, const LbShear& s, const float4 v0, const float4 v1, const float4 v2, float& t,
                                                  float& u, float& v) {
  float Akx, Aky, Akz, Bkx, Bky, Bkz, Ckx, Cky, Ckz;
  lb_permute(s, v0.x - r.ox, v0.y - r.oy, v0.z - r.oz, Akx, Aky, Akz);
  lb_permute(s, v1.x - r.ox, v1.y - r.oy, v1.z - r.oz, Bkx, Bky, Bkz);
  lb_permute(s, v2.x - r.ox, v2.y - r.oy, v2.z - r.oz, Ckx, Cky, Ckz);

  const float Ax = Akx - s.Sx * Akz;
  const float Ay = Aky - s.Sy * Akz;
  const float Bx = Bkx - s.Sx * Bkz;
  const float By = Bky - s.Sy * Bkz;
  const float Cx = Ckx - s.Sx * Ckz;
  const float Cy = Cky - s.Sy * Ckz;

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

  const float Az = s.Sz * Akz;
  const float Bz = s.Sz * Bkz;
  const float Cz = s.Sz * Ckz;
  const float T  = U * Az + V * Bz + W * Cz;

  const float rcp = 1.0f / det;
  t               = T * rcp;
  u               = V * rcp;
  v               = W * rcp;
  return true;
}

// 0x10 per byte -> 0xFF per byte. (__byte_perm masks its selector nibbles to 3 bits, so the sign-replicating
// mode of prmt is not reachable through it; a per-byte multiply cannot carry since 1 * 255 < 256.)
__device__ __forceinline__ uint32_t lb_expand_flag_bytes(uint32_t flags_0x10) { return (flags_0x10 >> 4) * 0xFFu; }

// Byte j of `packed` as the float 1 + b * 2^-15, built with one PRMT: (0x3F800000 | b << 8). An I2F per quantised
// plane (48 per node) made the XU pipe the busiest unit of the traversal kernels (ncu: 71 %); the byte permute runs on
// the ALU pipe instead and the conversion offset is folded into the per-node constants of lb_node_hits.
// `one` is the bit pattern of 1.0f held in a REGISTER that ptxas cannot constant-fold (the trace kernels pass it as a kernel
// argument): with both PRMT inputs known at compile time ptxas kept 0x3F800000 as the immediate and re-materialised the
// selector from a uniform register before almost every PRMT (about 40 extra IMAD.U32 per node step, cuobjdump -sass).
__device__ __forceinline__ float lb_u8_biased(uint32_t packed, int byte, uint32_t one) {
  return __uint_as_float(__byte_perm(packed, one, 0x7604u | ((uint32_t) byte << 4)));
}

// LB_NODE_PACKED (default 1): the byte -> float conversion of the 48 quantised planes of a node goes through half precision and
// the plane distances through the packed FP32 FMA of sm_100:
//   one PRMT builds a half2 {0x6400 | b_j, 0x6400 | b_j+1} = {1024 + b_j, 1024 + b_j+1} (exact: half has 11 significant bits),
//   HADD2.F32 widens each half (FMA pipe, exact), and one FFMA2 (fma.rn.f32x2) evaluates t = w * adj + org for both children.
// Per node 24 PRMT + 48 HADD2.F32 + 24 FFMA2 instead of 48 PRMT + 48 FFMA: the ALU pipe (PRMT, FMNMX, LOP3, SHF, SEL, ISETP; 16
// lanes per cycle and SM sub-partition) is the busiest unit of both traversal kernels (ncu: 77 %), the FMA pipe has room.
// 0 keeps the round-1 form (one PRMT per plane into the mantissa of a float in [1, 2)).
#ifndef LB_NODE_PACKED
#define LB_NODE_PACKED 1
#endif
#if LB_NODE_PACKED
#define LB_NODE_BIAS_BITS 0x64006400u
#else
#define LB_NODE_BIAS_BITS 0x3F800000u
#endif

// LB_HITMASK_V2 (default 1): hit-mask assembly of lb_node_hits with one PRMT per extracted byte and a modulo-32 shift (4 ALU-pipe
// instructions per child that is hit instead of 6 - 7: ptxas had sunk the inner-mask expansion into every predicated child block).
#ifndef LB_HITMASK_V2
#define LB_HITMASK_V2 1
#endif
#if LB_HITMASK_V2
#define LB_CHILD_CONTRIB(bits4, index4, j) (__byte_perm((bits4), 0u, 0x4440u | (uint32_t) (j)) << (((index4) >> (8 * (j))) & 31u))
#else
#define LB_CHILD_CONTRIB(bits4, index4, j) ((((bits4) >> (8 * (j))) & 0xFFu) << (((index4) >> (8 * (j))) & 0xFFu))
#endif

__device__ __forceinline__ float2 lb_ffma2(float2 a, float sb, float sc) {  // {a.x * sb + sc, a.y * sb + sc}, each one rounding
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(rb) : "f"(sb));
  asm("mov.b64 %0, {%1, %1};" : "=l"(rc) : "f"(sc));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}

// bytes (2 pair, 2 pair + 1) of `packed` as the floats 1024 + b; `bias` = 0x64006400 held in a register (see lb_u8_biased)
__device__ __forceinline__ float2 lb_u8x2_biased(uint32_t packed, int pair, uint32_t bias) {
  const uint32_t h2 = __byte_perm(packed, bias, pair ? 0x7372u : 0x7170u);  // bytes {b_lo, 0x64, b_hi, 0x64}: halves 0x6400 | b
  float2 f;
  asm("{\n\t"
      ".reg .b16 lo, hi;\n\t"
      "mov.b32 {lo, hi}, %2;\n\t"
      "cvt.f32.f16 %0, lo;\n\t"
      "cvt.f32.f16 %1, hi;\n\t"
      "}"
      : "=f"(f.x), "=f"(f.y)
      : "r"(h2));
  return f;
}

// Intersects the 8 quantised child boxes of one node. Returns the hit mask: bits 24..31 inner children in
// octant priority order, bits 0..23 triangle slots.
__device__ __forceinline__ uint32_t lb_node_hits(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, const uint4 n4, const LbRay& r,
                                                 const float idx, const float idy, const float idz, const uint32_t octinv4, const float tmax,
                                                 const uint32_t one = LB_NODE_BIAS_BITS) {
  // plane distance t = q * cell * id + (p - o) * id, evaluated as t = w * adj + org with
  //   LB_NODE_PACKED: w = 1024 + q,         adj = cell * id,         org = (p - o) * id - 1024 * adj
  //   else:           w = 1 + q * 2^-15,    adj = 2^15 * cell * id,  org = (p - o) * id - adj      (w = lb_u8_biased(q))
  // Folding costs one extra rounding of org, at most 2^-9 of a cell in t (2^-14 in the packed form); the builder rounds every child
  // box outwards with a margin of at least 1 / 64 of a cell (LB_QUANT_MARGIN, bvh_build.cu), so the test stays conservative.
  // Rounding errors that scale with the distance to the node are relative to t and covered by the 1 + 3.4 ulp factor of the
  // comparison below.
  const uint32_t ebits = n0.w;
#if LB_NODE_PACKED
  const float adjx = __uint_as_float((ebits & 0xFFu) << 23) * idx;
  const float adjy = __uint_as_float(((ebits >> 8) & 0xFFu) << 23) * idy;
  const float adjz = __uint_as_float(((ebits >> 16) & 0xFFu) << 23) * idz;
  const float orgx = fmaf(-1024.0f, adjx, (__uint_as_float(n0.x) - r.ox) * idx);
  const float orgy = fmaf(-1024.0f, adjy, (__uint_as_float(n0.y) - r.oy) * idy);
  const float orgz = fmaf(-1024.0f, adjz, (__uint_as_float(n0.z) - r.oz) * idz);
#else
  const float adjx     = __uint_as_float(((ebits & 0xFFu) + 15u) << 23) * idx;
  const float adjy     = __uint_as_float((((ebits >> 8) & 0xFFu) + 15u) << 23) * idy;
  const float adjz     = __uint_as_float((((ebits >> 16) & 0xFFu) + 15u) << 23) * idz;
  const float orgx     = (__uint_as_float(n0.x) - r.ox) * idx - adjx;
  const float orgy     = (__uint_as_float(n0.y) - r.oy) * idy - adjy;
  const float orgz     = (__uint_as_float(n0.z) - r.oz) * idz - adjz;
#endif

  uint32_t hitmask = 0;

#pragma unroll
  for (int half = 0; half < 2; half++) {
    const uint32_t meta4 = half ? n1.w : n1.z;
    const uint32_t qlox  = half ? n2.y : n2.x;
    const uint32_t qloy  = half ? n2.w : n2.z;
    const uint32_t qloz  = half ? n3.y : n3.x;
    const uint32_t qhix  = half ? n3.w : n3.z;
    const uint32_t qhiy  = half ? n4.y : n4.x;
    const uint32_t qhiz  = half ? n4.w : n4.z;

#if LB_HITMASK_V2
    // 0x10 per inner child -> 0x01 -> times the octant (0..7, no carries between bytes) = the XOR term of the slot index. Only the low
    // five bits of a byte of bit_index4 are the index; the shift below is taken modulo 32 (SHF.L.W), so they need no masking.
    const uint32_t is_inner4   = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t bit_index4  = meta4 ^ ((is_inner4 >> 4) * (octinv4 & 0xFFu));
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
#else
    const uint32_t is_inner4   = (meta4 & (meta4 << 1)) & 0x10101010u;
    const uint32_t inner_mask4 = lb_expand_flag_bytes(is_inner4);
    const uint32_t bit_index4  = (meta4 ^ (octinv4 & inner_mask4)) & 0x1F1F1F1Fu;
    const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
#endif

    // near / far planes per axis depend on the sign of the direction
    const uint32_t nearx = (idx < 0.0f) ? qhix : qlox;
    const uint32_t farx  = (idx < 0.0f) ? qlox : qhix;
    const uint32_t neary = (idy < 0.0f) ? qhiy : qloy;
    const uint32_t fary  = (idy < 0.0f) ? qloy : qhiy;
    const uint32_t nearz = (idz < 0.0f) ? qhiz : qloz;
    const uint32_t farz  = (idz < 0.0f) ? qloz : qhiz;

#if LB_NODE_PACKED
#pragma unroll
    for (int pair = 0; pair < 2; pair++) {
      const float2 tnx = lb_ffma2(lb_u8x2_biased(nearx, pair, one), adjx, orgx);
      const float2 tny = lb_ffma2(lb_u8x2_biased(neary, pair, one), adjy, orgy);
      const float2 tnz = lb_ffma2(lb_u8x2_biased(nearz, pair, one), adjz, orgz);
      const float2 tfx = lb_ffma2(lb_u8x2_biased(farx, pair, one), adjx, orgx);
      const float2 tfy = lb_ffma2(lb_u8x2_biased(fary, pair, one), adjy, orgy);
      const float2 tfz = lb_ffma2(lb_u8x2_biased(farz, pair, one), adjz, orgz);
      {
        const float tn = fmaxf(fmaxf(tnx.x, tny.x), fmaxf(tnz.x, r.tmin));
        const float tf = fminf(fminf(tfx.x, tfy.x), fminf(tfz.x, tmax));
        if (tn <= tf * 1.0000004f) {
          const int j = 2 * pair;
          hitmask |= LB_CHILD_CONTRIB(child_bits4, bit_index4, j);
        }
      }
      {
        const float tn = fmaxf(fmaxf(tnx.y, tny.y), fmaxf(tnz.y, r.tmin));
        const float tf = fminf(fminf(tfx.y, tfy.y), fminf(tfz.y, tmax));
        if (tn <= tf * 1.0000004f) {
          const int j = 2 * pair + 1;
          hitmask |= LB_CHILD_CONTRIB(child_bits4, bit_index4, j);
        }
      }
    }
#else
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const float tnx = fmaf(lb_u8_biased(nearx, j, one), adjx, orgx);
      const float tny = fmaf(lb_u8_biased(neary, j, one), adjy, orgy);
      const float tnz = fmaf(lb_u8_biased(nearz, j, one), adjz, orgz);
      const float tfx = fmaf(lb_u8_biased(farx, j, one), adjx, orgx);
      const float tfy = fmaf(lb_u8_biased(fary, j, one), adjy, orgy);
      const float tfz = fmaf(lb_u8_biased(farz, j, one), adjz, orgz);

      const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, r.tmin));
      const float tf = fminf(fminf(tfx, tfy), fminf(tfz, tmax));

      if (tn <= tf * 1.0000004f) {
        const uint32_t bits  = (child_bits4 >> (8 * j)) & 0xFFu;
        const uint32_t index = (bit_index4 >> (8 * j)) & 0xFFu;
        hitmask |= bits << index;
      }
    }
#endif
  }
  return hitmask;
}

// Generic traversal. `Visitor::hit(prim, t, u, v, tmax)` is called for every triangle whose watertight test
// passes with t in [tmin, tmax]; it may shrink tmax (closest hit) and returns true to terminate the ray.
struct LbTraversalCount {
  uint32_t nodes;
  uint32_t tris;
};

// Stack: sibling node groups only (triangle groups are tested immediately), one per level: depth <= LB_STACK_SIZE is enforced at build
// time; a drop is counted in *overflow (see trace_loop.cuh).
template <typename Visitor, bool kCount = false>
__device__ __forceinline__ void lb_traverse(const Bvh8& bvh, const LbRay& r, Visitor& vis, LbTraversalCount* count = nullptr,
                                            uint32_t* overflow = nullptr) {
  const float tiny = 8.271806125530277e-25f;  // 2^-80
  const float idx  = 1.0f / ((fabsf(r.dx) > tiny) ? r.dx : copysignf(tiny, r.dx));
  const float idy  = 1.0f / ((fabsf(r.dy) > tiny) ? r.dy : copysignf(tiny, r.dy));
  const float idz  = 1.0f / ((fabsf(r.dz) > tiny) ? r.dz : copysignf(tiny, r.dz));
  const uint32_t octinv  = ((r.dx >= 0.0f) ? 1u : 0u) | ((r.dy >= 0.0f) ? 2u : 0u) | ((r.dz >= 0.0f) ? 4u : 0u);
  const uint32_t octinv4 = octinv * 0x01010101u;
  const LbShear shear    = lb_shear(r);

  float tmax = r.tmax;

  uint2 stack[LB_STACK_SIZE];
  int sp = 0;

  uint2 group = make_uint2(0u, 0x80000000u);  // root: pretend a parent with the root in slot 7's priority bit
  // root is node 0: child_base 0, imask chosen so that the relative index is 0
  uint32_t root_pending = 1;

  for (;;) {
    uint2 tri_group = make_uint2(0u, 0u);

    if (root_pending || (group.y & 0xFF000000u)) {
      uint32_t node_index;
      if (root_pending) {
        root_pending = 0;
        node_index   = 0;
        group.y      = 0;
      }
      else {
        const uint32_t hits = group.y;
        const uint32_t bit  = 31u - __clz(hits);
        group.y &= ~(1u << bit);
        if (group.y & 0xFF000000u) {
          if (sp < LB_STACK_SIZE)
            stack[sp++] = group;
          else if (overflow)
            atomicAdd(overflow, 1u);
        }
        const uint32_t slot  = (bit - 24u) ^ octinv;
        const uint32_t imask = hits & 0xFFu;
        const uint32_t rel   = __popc(imask & ~(0xFFFFFFFFu << slot));
        node_index           = group.x + rel;
      }

      const uint4* np = bvh.nodes + 5 * (size_t) node_index;
      const uint4 n0  = __ldg(np + 0);
      const uint4 n1  = __ldg(np + 1);
      const uint4 n2  = __ldg(np + 2);
      const uint4 n3  = __ldg(np + 3);
      const uint4 n4  = __ldg(np + 4);

      const uint32_t hitmask = lb_node_hits(n0, n1, n2, n3, n4, r, idx, idy, idz, octinv4, tmax);
      if (kCount)
        count->nodes++;

      group.x     = n1.x;
      group.y     = (hitmask & 0xFF000000u) | (n0.w >> 24);
      tri_group.x = n1.y;
      tri_group.y = hitmask & 0x00FFFFFFu;
    }
    else {
      // no inner work left in the current group: pop
      if (sp == 0)
        return;
      group = stack[--sp];
      continue;
    }

    while (tri_group.y) {
      const uint32_t i = __ffs(tri_group.y) - 1u;
      tri_group.y &= tri_group.y - 1u;
      const float4* tp = bvh.tris + 3 * (size_t) (tri_group.x + i);
      const float4 v0  = __ldg(tp + 0);
      const float4 v1  = __ldg(tp + 1);
      const float4 v2  = __ldg(tp + 2);
      float t, u, v;
      if (kCount)
        count->tris++;
      if (lb_tri_watertight(r, shear, v0, v1, v2, t, u, v)) {
        if (t >= r.tmin && t <= tmax) {
          if (vis.hit(__float_as_uint(v0.w), t, u, v, tmax))
            return;
        }
      }
    }

    if ((group.y & 0xFF000000u) == 0) {
      if (sp == 0)
        return;
      group = stack[--sp];
    }
  }
}

// Closest hit with the reference's semantics (optix_kernel_raytrace.cu:82-95, optix_anyhit.cuh:15-31):
// the ignore handle is rejected, nothing else is (textures / alpha cut-outs are a "next" row).
struct LbClosestVisitor {
  uint32_t ignore_prim;
  LbHit best;

  __device__ __forceinline__ bool hit(uint32_t prim, float t, float u, float v, float& tmax) {
    if (prim == ignore_prim)
      return false;
    if (t < best.t || (t == best.t && best.prim != LB_HIT_SKY && prim < best.prim)) {
      best.prim = prim;
      best.t    = t;
      best.u    = u;
      best.v    = v;
      tmax      = t;
    }
    return false;
  }
};

template <bool kCount = false>
__device__ __forceinline__ LbHit lb_closest_hit(const Bvh8& bvh, const LbRay& r, uint32_t ignore_prim, LbTraversalCount* count = nullptr) {
  LbClosestVisitor vis;
  vis.ignore_prim = ignore_prim;
  vis.best.prim   = LB_HIT_SKY;
  vis.best.t      = r.tmax;
  vis.best.u      = 0.0f;
  vis.best.v      = 0.0f;
  lb_traverse<LbClosestVisitor, kCount>(bvh, r, vis, count);
  if (vis.best.prim == LB_HIT_SKY)
    vis.best.t = 3.402823466e+38f;
  return vis.best;
}
