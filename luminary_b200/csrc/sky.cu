// sky.cu - kernels of the procedural atmosphere: LUT generation and the miss shader of LUMINARY_SKY_MODE_DEFAULT.
//
// Replaces, on the reference's path:
//   sky_compute_transmittance_lut     cuda/sky.cuh:109-178   ([Bru17]: 2500-step optical depth per texel of the 256 x 64 table)
//   sky_compute_multiscattering_lut   cuda/sky.cuh:185-330   ([Hil20]: 256 directions x 500 steps per texel of the 32 x 32 table, block reduction)
//   sky_process_tasks (DEFAULT / HDRI) cuda/sky.cuh:609-633 -> sky_color_main -> sky_compute_atmosphere | sky_hdri_sample
//   sky_compute_hdri                  cuda/sky_hdri.cuh:60-158 (the HDRI mode's bake)
// Launch geometry of the LUT kernels follows device_sky.c:80-117 (one thread per transmittance texel, one 256-thread block per
// multiscattering texel). Compiled with --use_fast_math like the reference's kernels.
#include "rng.cuh"
#include "shade_api.cuh"
#include "sky.cuh"

using namespace lbsky;

// [Bru17] sky_compute_transmittance_optical_depth, sky.cuh:110-141
__device__ static Spectrum transmittance_optical_depth(const LbSkyDev& S, float r, float mu) {
  const int steps       = 2500;
  const float disc      = r * r * (mu * mu - 1.0f) + LB_SKY_ATMO_RADIUS * LB_SKY_ATMO_RADIUS;
  const float dist      = fmaxf(-r * mu + sqrtf(fmaxf(0.0f, disc)), 0.0f);
  const float step_size = dist / steps;
  Spectrum depth        = s_set1(0.0f);
#pragma unroll 1
  for (int i = 0; i <= steps; i++) {
    const float reach  = i * step_size;
    const float height = sqrtf(reach * reach + 2.0f * r * mu * reach + r * r) - LB_SKY_EARTH_RADIUS;
    const Medium m     = medium_at(S, height);
    const float w      = (i == 0 || i == steps) ? 0.5f : 1.0f;
    depth              = s_add(depth, s_scale(m.extinction, w * step_size));
  }
  return depth;
}

__global__ void __launch_bounds__(128) k_sky_transmittance_lut(LbSkyDev S, float4* __restrict__ dst_low, float4* __restrict__ dst_high) {
  const int amount = LB_SKY_TM_TEX_WIDTH * LB_SKY_TM_TEX_HEIGHT;
  for (unsigned int id = blockIdx.x * blockDim.x + threadIdx.x; id < amount; id += blockDim.x * gridDim.x) {
    const int y = id / LB_SKY_TM_TEX_WIDTH;
    const int x = id - y * LB_SKY_TM_TEX_WIDTH;
    float fx    = ((float) x + 0.5f) / LB_SKY_TM_TEX_WIDTH;
    float fy    = ((float) y + 0.5f) / LB_SKY_TM_TEX_HEIGHT;
    fx          = sub_to_unit_uv(fx, LB_SKY_TM_TEX_WIDTH);
    fy          = sub_to_unit_uv(fy, LB_SKY_TM_TEX_HEIGHT);

    const float H   = sqrtf(LB_SKY_ATMO_RADIUS * LB_SKY_ATMO_RADIUS - LB_SKY_EARTH_RADIUS * LB_SKY_EARTH_RADIUS);
    const float rho = H * fy;
    const float r   = sqrtf(rho * rho + LB_SKY_EARTH_RADIUS * LB_SKY_EARTH_RADIUS);

    const float d_min = LB_SKY_ATMO_RADIUS - r;
    const float d_max = rho + H;
    const float d     = d_min + fx * (d_max - d_min);

    float mu = (d == 0.0f) ? 1.0f : (H * H - rho * rho - d * d) / (2.0f * r * d);
    mu       = fminf(1.0f, fmaxf(-1.0f, mu));

    const Spectrum transmittance = s_exp(s_scale(transmittance_optical_depth(S, r, mu), -1.0f));
    dst_low[x + y * LB_SKY_TM_TEX_WIDTH]  = s_low(transmittance);
    dst_high[x + y * LB_SKY_TM_TEX_WIDTH] = s_high(transmittance);
  }
}

// [Hil20] sky_compute_multiscattering_integration, sky.cuh:186-272
__device__ static void multiscattering_integration(const LbSkyDev& S, V3 origin, V3 ray, V3 sun, Spectrum& L, Spectrum& as1) {
  L   = s_set1(0.0f);
  as1 = s_set1(0.0f);
  const float2 path = compute_path(origin, ray, LB_SKY_EARTH_RADIUS, LB_SKY_ATMO_RADIUS);
  if (path.y == -FLT_MAX)
    return;
  const float start    = path.x;
  const float distance = path.y;
  if (distance > 0.0f) {
    const int steps = 500;
    float reach     = start;
    float step_size;
    const float light_angle = sample_sphere_solid_angle(sun, LB_SKY_SUN_RADIUS, origin);
    Spectrum transmittance  = s_set1(1.0f);
    const JendersieEon mie  = jendersie_eon_parameters(S.mie_diameter);
#pragma unroll 1
    for (int i = 0; i < steps; i++) {
      const float new_reach = start + distance * (i + 0.3f) / steps;
      step_size             = new_reach - reach;
      reach                 = new_reach;

      const V3 pos       = origin + ray * reach;
      const float height = sky_height(pos);

      const V3 ray_scatter         = normalize3(sun - pos);
      const float cos_angle        = dot3(ray, ray_scatter);
      const float phase_rayleigh   = rayleigh_phase(cos_angle);
      const float phase_mie        = jendersie_eon_phase(cos_angle, mie);
      const float zenith_cos_angle = dot3(normalize3(pos), ray_scatter);

      const float2 tm_uv            = transmittance_lut_uv(height, zenith_cos_angle);
      const Spectrum extinction_sun = s_merge(tex2D<float4>(S.tm_low, tm_uv.x, tm_uv.y), tex2D<float4>(S.tm_high, tm_uv.x, tm_uv.y));

      const Medium m            = medium_at(S, height);
      const Spectrum scattering = s_add(m.scattering_rayleigh, s_set1(m.scattering_mie));
      const Spectrum phase_times_scattering = s_add(s_scale(m.scattering_rayleigh, phase_rayleigh), s_set1(m.scattering_mie * phase_mie));

      const float shadow   = sph_ray_hit_p0(ray_scatter, pos, LB_SKY_EARTH_RADIUS) ? 0.0f : 1.0f;
      const Spectrum Sterm = s_scale(s_mul(extinction_sun, phase_times_scattering), shadow * light_angle);

      const Spectrum step_transmittance = s_exp(s_scale(m.extinction, -step_size));
      const Spectrum ss_int = s_mul(s_sub(Sterm, s_mul(Sterm, step_transmittance)), s_inv(m.extinction));
      const Spectrum ms_int = s_mul(s_sub(scattering, s_mul(scattering, step_transmittance)), s_inv(m.extinction));

      L             = s_add(L, s_mul(ss_int, transmittance));
      as1           = s_add(as1, s_mul(ms_int, transmittance));
      transmittance = s_mul(transmittance, step_transmittance);
    }
  }
}

// one block of LB_SKY_MS_ITER threads per texel; the tree reduction keeps the reference's summation order
__global__ void __launch_bounds__(LB_SKY_MS_ITER) k_sky_multiscattering_lut(LbSkyDev S, float4* __restrict__ dst_low, float4* __restrict__ dst_high) {
  const int x = blockIdx.x;
  const int y = blockIdx.y;
  float fx    = ((float) x + 0.5f) / LB_SKY_MS_TEX_SIZE;
  float fy    = ((float) y + 0.5f) / LB_SKY_MS_TEX_SIZE;
  fx          = sub_to_unit_uv(fx, LB_SKY_MS_TEX_SIZE);
  fy          = sub_to_unit_uv(fy, LB_SKY_MS_TEX_SIZE);

  __shared__ Spectrum luminance_shared[LB_SKY_MS_ITER];
  __shared__ Spectrum multiscattering_shared[LB_SKY_MS_ITER];

  const float cos_angle = fx * 2.0f - 1.0f;
  const V3 sun_dir      = v3(0.0f, cos_angle, sqrtf(__saturatef(1.0f - cos_angle * cos_angle)));
  const float height    = LB_SKY_EARTH_RADIUS + __saturatef(fy + LB_SKY_HEIGHT_OFFSET) * (LB_SKY_ATMO_HEIGHT - LB_SKY_HEIGHT_OFFSET);
  const V3 pos          = v3(0.0f, height, 0.0f);
  const V3 sun          = sun_dir * LB_SKY_SUN_DISTANCE;

  const float sqrt_sample = (float) LB_SKY_MS_BASE;
  const float a           = threadIdx.x / LB_SKY_MS_BASE;
  const float b           = (threadIdx.x - ((threadIdx.x / LB_SKY_MS_BASE) * LB_SKY_MS_BASE));
  const float randA       = a / sqrt_sample;
  const float randB       = b / sqrt_sample;
  const V3 ray            = sample_ray_sphere(2.0f * randA - 1.0f, randB);

  Spectrum L, as1;
  multiscattering_integration(S, pos, ray, sun, L, as1);
  luminance_shared[threadIdx.x]       = L;
  multiscattering_shared[threadIdx.x] = as1;

  for (int i = LB_SKY_MS_ITER >> 1; i > 0; i = i >> 1) {
    __syncthreads();
    if (threadIdx.x < i) {
      luminance_shared[threadIdx.x]       = s_add(luminance_shared[threadIdx.x], luminance_shared[threadIdx.x + i]);
      multiscattering_shared[threadIdx.x] = s_add(multiscattering_shared[threadIdx.x], multiscattering_shared[threadIdx.x + i]);
    }
  }
  if (threadIdx.x > 0)
    return;

  const Spectrum luminance       = s_scale(luminance_shared[0], 1.0f / (sqrt_sample * sqrt_sample));
  const Spectrum multiscattering = s_scale(multiscattering_shared[0], 1.0f / (sqrt_sample * sqrt_sample));
  const Spectrum contribution    = s_inv(s_sub(s_set1(1.0f), multiscattering));
  const Spectrum out             = s_scale(s_mul(luminance, contribution), S.multiscattering_factor);
  dst_low[x + y * LB_SKY_MS_TEX_SIZE]  = s_low(out);
  dst_high[x + y * LB_SKY_MS_TEX_SIZE] = s_high(out);
}

void lb_launch_sky_transmittance_lut(const LbSkyDev& sky, float4* dst_low, float4* dst_high, cudaStream_t s) {
  const int amount = LB_SKY_TM_TEX_WIDTH * LB_SKY_TM_TEX_HEIGHT;
  k_sky_transmittance_lut<<<(amount + 127) / 128, 128, 0, s>>>(sky, dst_low, dst_high);
}

void lb_launch_sky_multiscattering_lut(const LbSkyDev& sky, float4* dst_low, float4* dst_high, cudaStream_t s) {
  k_sky_multiscattering_lut<<<dim3(LB_SKY_MS_TEX_SIZE, LB_SKY_MS_TEX_SIZE), LB_SKY_MS_ITER, 0, s>>>(sky, dst_low, dst_high);
}

// sky_process_tasks in DEFAULT mode: the misses are the tail [n_hits, n_active) of the sorted queue. One thread marches one ray
// (sky.steps steps, four float4 LUT fetches each); every miss of every bounce is shaded (geometry.cuh:123-126 keeps ALLOW_AMBIENT set).
template <bool kAdaptive, bool kHdri>
__global__ void __launch_bounds__(128) k_shade_miss_sky(LbShadeParams P) {
  const uint32_t n_active = P.counters->n_active;
  const uint32_t n_hits   = P.counters->n_hits;
  for (uint32_t k = n_hits + blockIdx.x * blockDim.x + threadIdx.x; k < n_active; k += gridDim.x * blockDim.x) {
    const uint32_t i     = P.queue_in[k];
    const uint32_t state = P.paths.state[i];
    if (!(state & LB_STATE_ALLOW_AMBIENT))
      continue;
    const float4 o4      = P.paths.org[i];
    const float4 d4      = P.paths.dir[i];
    const uint32_t pixel = P.paths.pixel[i];
    const uint32_t py    = pixel / P.frame.width;
    const uint32_t px    = pixel - py * P.frame.width;
    float random_offset = 0.0f;
    if constexpr (kHdri) {
    }
    else if constexpr (kAdaptive) {
      lbrng::Sampler smp;
      smp.bluenoise = P.bluenoise, smp.px = px, smp.py = py, smp.sample_id = P.paths.sample_id[i], smp.depth = P.rng_depth;
      random_offset = smp.get1(lbrng::T_SKY_STEP_OFFSET);
    }
    else {
      lbrng::TabSampler smp;
      smp.bluenoise = P.bluenoise, smp.table = P.rng_table + P.rng_depth * lbrng::T_COUNT, smp.px = px, smp.py = py;
      random_offset = smp.get1(lbrng::T_SKY_STEP_OFFSET);
    }
    const bool include_sun = (state & (LB_STATE_CAMERA_DIRECTION | LB_STATE_ALLOW_EMISSION)) != 0;
    const float3 sky       = kHdri ? sky_color_hdri(P.sky, v3(o4.x, o4.y, o4.z), v3(d4.x, d4.y, d4.z), include_sun)
                                   : sky_color(P.sky, v3(o4.x, o4.y, o4.z), v3(d4.x, d4.y, d4.z), include_sun, random_offset);
    // record_unpack, math.cuh:1595-1607
    const uint2 rec = P.paths.record[i];
    const float rr  = __uint_as_float((rec.x & 0x1FFFFFu) << 11);
    const float rg  = __uint_as_float(((rec.x >> 21) | ((rec.y & 0x3FFu) << 11)) << 11);
    const float rb  = __uint_as_float((rec.y >> 10) << 11);
    const float sr = sky.x * rr, sg = sky.y * rg, sb = sky.z * rb;
    if (sr != 0.0f || sg != 0.0f || sb != 0.0f) {
      float4 res = P.paths.result[i];
      res.x += sr, res.y += sg, res.z += sb;
      P.paths.result[i] = res;
    }
  }
}

// sky_process_inscattering_events (cuda/kernels.cuh:356-389) + sky_trace_inscattering (cuda/sky.cuh:517-532): aerial perspective. Runs
// between the closest-hit trace and the sort, over the UNSORTED queue: every hit adds the light scattered into its segment
// (x throughput) to the path's result and attenuates the throughput by the segment's transmittance.
template <bool kAdaptive>
__global__ void __launch_bounds__(128) k_sky_inscattering(LbShadeParams P) {
  const uint32_t n_active = P.counters->n_active;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_active; k += gridDim.x * blockDim.x) {
    const uint32_t i = P.queue_in[k];
    if (P.paths.prim[i] == LB_HIT_SKY)
      continue;
    const float4 o4      = P.paths.org[i];
    const float4 d4      = P.paths.dir[i];
    const uint32_t pixel = P.paths.pixel[i];
    const uint32_t py    = pixel / P.frame.width;
    const uint32_t px    = pixel - py * P.frame.width;
    float rnd_steps, rnd_offset;
    if constexpr (kAdaptive) {
      lbrng::Sampler smp;
      smp.bluenoise = P.bluenoise, smp.px = px, smp.py = py, smp.sample_id = P.paths.sample_id[i], smp.depth = P.rng_depth;
      rnd_steps = smp.get1(lbrng::T_SKY_INSCATTERING_STEP), rnd_offset = smp.get1(lbrng::T_SKY_STEP_OFFSET);
    }
    else {
      lbrng::TabSampler smp;
      smp.bluenoise = P.bluenoise, smp.table = P.rng_table + P.rng_depth * lbrng::T_COUNT, smp.px = px, smp.py = py;
      rnd_steps = smp.get1(lbrng::T_SKY_INSCATTERING_STEP), rnd_offset = smp.get1(lbrng::T_SKY_STEP_OFFSET);
    }
    const V3 sky_origin    = world_to_sky(P.sky, v3(o4.x, o4.y, o4.z));
    const float limit      = d4.w * 0.001f;                           // world_to_sky_scale(trace.depth)
    const float base_range = (P.rng_depth == 0) ? 40.0f : 80.0f;      // IS_PRIMARY_RAY = device.state.depth == 0
    const int steps        = fminf(fmaxf(0.5f, limit / base_range), 2.0f) * (P.sky.steps / 6) + rnd_steps - 0.5f;
    Spectrum transmittance = s_set1(1.0f);
    const Spectrum radiance = compute_atmosphere(P.sky, sky_origin, v3(d4.x, d4.y, d4.z), limit, false, steps, rnd_offset, &transmittance);
    const float3 ins        = color_from_spectrum(radiance);
    const float3 tr         = color_from_spectrum(transmittance);
    // record_unpack / record_pack, math.cuh:1580-1607
    const uint2 rec = P.paths.record[i];
    float rr        = __uint_as_float((rec.x & 0x1FFFFFu) << 11);
    float rg        = __uint_as_float(((rec.x >> 21) | ((rec.y & 0x3FFu) << 11)) << 11);
    float rb        = __uint_as_float((rec.y >> 10) << 11);
    const float sr = ins.x * rr, sg = ins.y * rg, sb = ins.z * rb;
    if (sr > 0.0f || sg > 0.0f || sb > 0.0f) {  // write_beauty_buffer: color_any
      float4 res = P.paths.result[i];
      res.x += sr, res.y += sg, res.z += sb;
      P.paths.result[i] = res;
    }
    rr *= tr.x, rg *= tr.y, rb *= tr.z;
    const uint32_t br = __float_as_uint(rr) >> 11, bg = __float_as_uint(rg) >> 11, bb = __float_as_uint(rb) >> 11;
    P.paths.record[i] = make_uint2(br | (bg << 21), (bg >> 11) | (bb << 10));
  }
}

void lb_launch_sky_inscattering(const LbShadeParams& sp, int grid, cudaStream_t s) {
  if (sp.adaptive)
    k_sky_inscattering<true><<<grid, 128, 0, s>>>(sp);
  else
    k_sky_inscattering<false><<<grid, 128, 0, s>>>(sp);
}

void lb_launch_shade_miss_sky(const LbShadeParams& sp, int grid, cudaStream_t s) {
  if (sp.sky.mode == 1)
    k_shade_miss_sky<false, true><<<grid, 128, 0, s>>>(sp);
  else if (sp.adaptive)
    k_shade_miss_sky<true, false><<<grid, 128, 0, s>>>(sp);
  else
    k_shade_miss_sky<false, false><<<grid, 128, 0, s>>>(sp);
}

// ---------------------------------------------------------------------------------------------------------------------------
// HDRI bake: sky_compute_hdri (cuda/sky_hdri.cuh:60-158, no clouds). One warp per texel; lane l integrates the samples l, l + 32, ...;
// the per-lane means are combined by sky_hdri_warp_apply_median_of_means (sky_hdri.cuh:13-58), which lane 0 runs on the warp's 32
// shared-memory slots, channel by channel, exactly as the reference does (insertion sort, Gini-weighted trimmed mean).
// ---------------------------------------------------------------------------------------------------------------------------
__device__ static float hdri_median_of_means(float* buckets, const uint32_t num_buckets) {
  for (uint32_t i = 1; i < num_buckets; i++) {
    const float x = buckets[i];
    uint32_t j    = i;
    while (j > 0 && buckets[j - 1] > x) {
      buckets[j] = buckets[j - 1];
      j--;
    }
    buckets[j] = x;
  }
  float num = 0.0f, denom = 0.0f;
  for (uint32_t b = 0; b < num_buckets; b++) {
    const float value = buckets[b];
    num += b * value;
    denom += value;
  }
  num *= 2.0f;
  denom *= num_buckets;
  const float G    = __saturatef((num / denom) - (num_buckets + 1.0f) / num_buckets);
  const uint32_t k = num_buckets >> 1;
  const uint32_t c = k - (1.0f - G) * k;
  float output     = 0.0f;
  for (uint32_t b = c; b < num_buckets - c; b++)
    output += buckets[b];
  output /= num_buckets - 2 * c;
  return output;
}

__global__ void __launch_bounds__(128) k_sky_hdri(LbSkyDev S, const uint32_t* __restrict__ bluenoise, float ox, float oy, float oz, uint32_t dim,
                                                  uint32_t sample_count, float4* __restrict__ dst) {
  __shared__ float s_values[128];
  const uint32_t lane     = threadIdx.x & 31u;
  const uint32_t pixel_id = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (pixel_id >= dim * dim)  // warp-uniform
    return;
  const uint32_t y      = pixel_id / dim;
  const uint32_t x      = pixel_id - y * dim;
  const float step_size = 1.0f / (dim - 1);
  float cr = 0.0f, cg = 0.0f, cb = 0.0f;
  uint32_t num_samples = 0;
  const V3 sky_origin  = world_to_sky(S, v3(ox, oy, oz));
  for (uint32_t sample_id = lane; sample_id < sample_count; sample_id += 32) {
    lbrng::Sampler smp;
    smp.bluenoise = bluenoise, smp.px = x, smp.py = y, smp.sample_id = sample_id, smp.depth = 0;
    const float2 jitter  = smp.get2(lbrng::T_CAMERA_JITTER);
    const float u        = (((float) x) + jitter.x) * step_size;
    const float v        = 1.0f - (((float) y) + jitter.y) * step_size;
    const float altitude = LB_SKY_PI * v - 0.5f * LB_SKY_PI;
    const float azimuth  = 2.0f * LB_SKY_PI * u - LB_SKY_PI;
    const V3 ray         = v3(cosf(azimuth) * cosf(altitude), sinf(altitude), sinf(azimuth) * cosf(altitude));
    const float3 sky     = color_from_spectrum(compute_atmosphere(S, sky_origin, ray, FLT_MAX, false, (int) S.steps, smp.get1(lbrng::T_SKY_STEP_OFFSET)));
    cr += sky.x, cg += sky.y, cb += sky.z;
    num_samples++;
  }
  const uint32_t bucket_count = min(32u, sample_count);
  float* buckets              = s_values + (threadIdx.x & ~31u);
  float out[3];
  const float mean[3] = {num_samples ? cr / num_samples : 0.0f, num_samples ? cg / num_samples : 0.0f, num_samples ? cb / num_samples : 0.0f};
#pragma unroll
  for (int c = 0; c < 3; c++) {
    __syncwarp();
    buckets[lane] = mean[c];
    __syncwarp();
    out[c] = (lane == 0) ? hdri_median_of_means(buckets, bucket_count) : 0.0f;
  }
  if (lane == 0)
    dst[x + y * dim] = make_float4(out[0], out[1], out[2], 0.0f);
}

void lb_launch_sky_hdri(const LbSkyDev& sky, const uint32_t* bluenoise, const float origin[3], uint32_t dim, uint32_t sample_count, float4* dst,
                        cudaStream_t s) {
  const uint64_t threads = (uint64_t) dim * dim * 32u;
  k_sky_hdri<<<(uint32_t) ((threads + 127u) / 128u), 128, 0, s>>>(sky, bluenoise, origin[0], origin[1], origin[2], dim, sample_count, dst);
}
