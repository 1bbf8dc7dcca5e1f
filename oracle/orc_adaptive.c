/* orc_adaptive.c - CPU restatement of the reference's adaptive sampler. TEST INFRASTRUCTURE ONLY.
 *
 * Reference: cuda/adaptive_sampling.cuh (block variance :166-199, stage sample counts :201-221, sample offsets / counts :54-102),
 * device/device_adaptive_sampler.c (allocation :58-71, stage build :91-215), device_renderer.c:350-376 (a stage is rebuilt after
 * update_interval << stage_id executions), cuda/kernels.cuh:195-356 (tasks_create_adaptive_sampling), cuda/accumulation.cuh:86-190.
 *
 * The image is tiled into 4 x 4 pixel blocks. Stage 0 renders one sample per pixel and execution. After update_interval << s
 * executions of stage s the per-block sample counts of stage s + 1 are computed from the variance estimate (byte s of the block's
 * 32-bit word holds count - 1); an execution of stage s + 1 then renders that many samples for every pixel of the block. A pixel's
 * sample ids are consecutive over its whole history, and means divide by the pixel's own sample count.
 * The reference builds a stage asynchronously (the switch happens when the main device's event has fired); this restatement and the
 * product switch exactly at the threshold, which is when the reference queues the build. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "orc_internal.h"

void orc_tonemap_rgb(float rgb[3], float exposure, uint32_t tonemap, float agx_slope, float agx_power, float agx_saturation);

uint32_t orc_adaptive_stage_count(uint32_t word, uint32_t stage /* 0..3 = byte */) { return ((word >> (stage * 8)) & 0xFFu) + 1u; }

/* adaptive_sampling_get_sample_count_from_block_index for a finished execution: samples every pixel of the block has received */
uint32_t orc_adaptive_block_samples(uint32_t word, const uint32_t executions[ORC_ADAPTIVE_STAGES + 1]) {
  uint32_t count = executions[0];
  for (uint32_t s = 0; s < ORC_ADAPTIVE_STAGES; s++)
    count += executions[s + 1] * orc_adaptive_stage_count(word, s);
  return (count < (1u << 20)) ? count : (1u << 20);
}

/* adaptive_sampling_compute_tonemap_compression_factor, adaptive_sampling.cuh:9-17 */
static float compression(const OrcAdaptiveParams* p, float r, float g, float b) {
  float exposed[3] = {r * p->exposure, g * p->exposure, b * p->exposure};
  const float ev   = 0.212655f * exposed[0] + 0.715158f * exposed[1] + 0.072187f * exposed[2];
  float tm[3]      = {r, g, b};
  orc_tonemap_rgb(tm, p->exposure, p->tonemap, p->agx_slope, p->agx_power, p->agx_saturation);
  const float tv = 0.212655f * tm[0] + 0.715158f * tm[1] + 0.072187f * tm[2];
  return (ev > 0.0f) ? tv / ev : 1.0f;
}

/* adaptive_sampling_block_reduce_variance: max over the block's pixels of max(E[lum(x^2)] - lum(E[x]^2), 0), times the squared
 * tone-map compression when exposure-aware. Returns the sum over all blocks. */
float orc_adaptive_block_variance(const float* planes, uint32_t width, uint32_t height, const uint32_t* words,
                                  const uint32_t executions[ORC_ADAPTIVE_STAGES + 1], const OrcAdaptiveParams* p, float* block_variance) {
  const uint32_t bw = (width + 3) >> 2, bh = (height + 3) >> 2;
  const size_t n    = (size_t) width * height;
  double sum        = 0.0;
  for (uint32_t by = 0; by < bh; by++)
    for (uint32_t bx = 0; bx < bw; bx++) {
      const uint32_t b = bx + by * bw;
      const float inv  = 1.0f / (float) orc_adaptive_block_samples(words[b], executions);
      float best       = 0.0f;
      for (uint32_t ly = 0; ly < 4; ly++)
        for (uint32_t lx = 0; lx < 4; lx++) {
          const uint32_t x = 4 * bx + lx, y = 4 * by + ly;
          if (x >= width || y >= height)
            continue;
          const size_t i = x + (size_t) y * width;
          const float r = planes[i] * inv, g = planes[n + i] * inv, bl = planes[2 * n + i] * inv;
          const float l2 = planes[3 * n + i] * inv;
          const float ls = 0.212655f * (r * r) + 0.715158f * (g * g) + 0.072187f * (bl * bl);
          float var      = fmaxf(l2 - ls, 0.0f);
          if (p->exposure != 0.0f) {
            const float c = compression(p, r, g, bl);
            var *= c * c;
          }
          best = fmaxf(best, var);
        }
      block_variance[b] = fabsf(best);
      sum += block_variance[b];
    }
  return (float) sum;
}

/* adaptive_sampling_compute_stage_sample_counts: byte `stage` (0..3) of every word */
void orc_adaptive_stage_counts(const float* block_variance, float sum_variance, uint32_t num_blocks, uint32_t stage, const OrcAdaptiveParams* p,
                               uint32_t* words) {
  const float avg = sum_variance / (float) num_blocks;
  for (uint32_t b = 0; b < num_blocks; b++) {
    uint32_t w    = words[b] & ((1u << (stage * 8)) - 1u);
    const float v = block_variance[b] / avg * (float) p->avg_sampling_rate + 0.5f; /* remap(variance, 0, avg, 0, rate) + 0.5 */
    uint32_t c    = (v >= 0.0f && v < 4294967296.0f) ? (uint32_t) v : ((v >= 4294967296.0f) ? 0xFFFFFFFFu : 0u); /* NaN -> 0 like cvt.rzi.u32 */
    if (c < 1)
      c = 1;
    if (c > p->max_sampling_rate)
      c = p->max_sampling_rate;
    words[b] = w | ((c - 1u) << (stage * 8));
  }
}

/* accumulation_generate_result (cuda/accumulation.cuh:86-190): every output mode + the camera's local error minimisation.
 * words == NULL: adaptive sampling off, every pixel has uniform_count samples. rgb = three planes. */
typedef struct {
  float r, g, b, variance, inv_n;
} ResolvePixel;

static ResolvePixel resolve_pixel(const float* planes, uint32_t width, uint32_t height, uint32_t x, uint32_t y, const uint32_t* words,
                                  const uint32_t* executions, uint32_t uniform_count) {
  const size_t n     = (size_t) width * height;
  const size_t i     = x + (size_t) y * width;
  const uint32_t bw  = (width + 3) >> 2;
  const uint32_t cnt = words ? orc_adaptive_block_samples(words[(x >> 2) + (y >> 2) * bw], executions) : uniform_count;
  ResolvePixel p;
  p.inv_n = 1.0f / (float) cnt;
  p.r = planes[i] * p.inv_n, p.g = planes[n + i] * p.inv_n, p.b = planes[2 * n + i] * p.inv_n;
  const float ls = 0.212655f * (p.r * p.r) + 0.715158f * (p.g * p.g) + 0.072187f * (p.b * p.b);
  p.variance     = fmaxf(planes[3 * n + i] * p.inv_n - ls, 0.0f);
  return p;
}

void orc_resolve(const float* planes, uint32_t width, uint32_t height, const uint32_t* words, const uint32_t* executions, uint32_t uniform_count,
                 uint32_t mode, int local_error_minimization, uint32_t stage, const OrcAdaptiveParams* p, float* rgb) {
  const size_t n    = (size_t) width * height;
  const uint32_t bw = (width + 3) >> 2;
  for (uint32_t y = 0; y < height; y++)
    for (uint32_t x = 0; x < width; x++) {
      const ResolvePixel c = resolve_pixel(planes, width, height, x, y, words, executions, uniform_count);
      float out[3]         = {c.r, c.g, c.b};
      if (mode == 0 && local_error_minimization) {
        const float center_error = c.variance * c.inv_n;
        const uint32_t x0 = (x > 1 ? x : 1) - 1, x1 = (x < width - 1 ? x : width - 1) + 1;
        const uint32_t y0 = (y > 1 ? y : 1) - 1, y1 = (y < height - 1 ? y : height - 1) + 1;
        float nm[3] = {0.0f, 0.0f, 0.0f}, nerr = 0.0f;
        for (uint32_t yi = y0; yi <= y1; yi++)
          for (uint32_t xi = x0; xi <= x1; xi++) {
            if ((xi == x && yi == y) || xi >= width || yi >= height)
              continue; /* out-of-frame neighbours contribute 0 but are counted (accumulation.cuh:113-135) */
            const ResolvePixel q = resolve_pixel(planes, width, height, xi, yi, words, executions, uniform_count);
            nm[0] += q.r, nm[1] += q.g, nm[2] += q.b;
            nerr += q.variance * q.inv_n;
          }
        const float nn = 1.0f / (float) ((x1 - x0 + 1) * (y1 - y0 + 1) - 1);
        nm[0] *= nn, nm[1] *= nn, nm[2] *= nn;
        nerr *= nn;
        float t = center_error / (8.0f * nerr);
        t       = (t != t) ? 0.0f : fminf(fmaxf(t, 0.0f), 1.0f); /* __saturatef(NaN) = 0 */
        for (int k = 0; k < 3; k++)
          out[k] = out[k] + t * (nm[k] - out[k]);
      }
      else if (mode == 1)
        out[0] = out[1] = out[2] = 128.0f * c.variance;
      else if (mode == 2) {
        const float value = 1024.0f * (sqrtf(c.variance * c.inv_n) * compression(p, c.r, c.g, c.b));
        out[0]            = orc_saturate(2.0f * value);
        out[1]            = orc_saturate(2.0f * (value - 0.5f));
        out[2]            = orc_saturate((value > 0.5f) ? 4.0f * (0.25f - fabsf(value - 1.0f)) : 4.0f * (0.25f - fabsf(value - 0.25f)));
      }
      else if (mode == 3) {
        const uint32_t tpp = (words && stage > 0) ? orc_adaptive_stage_count(words[(x >> 2) + (y >> 2) * bw], stage - 1) : 1u;
        out[0] = out[1] = out[2] = (float) tpp / 256.0f;
      }
      const size_t i = x + (size_t) y * width;
      rgb[i] = out[0], rgb[n + i] = out[1], rgb[2 * n + i] = out[2];
    }
}
