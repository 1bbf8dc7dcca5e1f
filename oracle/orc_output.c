/* orc_output.c - CPU restatement of the reference's output chain for parity tests. TEST INFRASTRUCTURE ONLY.
 *
 * Follows generate_final_image + convert_RGBF_to_ARGB8 (device/cuda/kernels.cuh:503-644) with tonemap_apply
 * (device/cuda/tonemap.cuh:7-246), linearRGB_to_SRGB / SRGB_to_linearRGB (cuda/math.cuh:1044-1060), color_luminance
 * (:1062-1064) and random_dither_mask (cuda/random.cuh:150-154,370-375), for supersampling 0, undersampling 0, filter
 * NONE, purkinje / colour correction / film grain off. exposure is the linear factor expf(camera.exposure)
 * (device_structs.c:77). Parity unpinned against a running reference (GPU-only); the formulas are closed form. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "lum_oracle.h"

typedef struct {
  float r, g, b;
} Col;

static float lum(Col v) { return 0.212655f * v.r + 0.715158f * v.g + 0.072187f * v.b; }
static float to_srgb(float v) { return (v <= 0.0031308f) ? 12.92f * v : 1.055f * powf(v, 0.416666666667f) - 0.055f; }
static float to_linear(float v) { return (v <= 0.04045f) ? v / 12.92f : powf((v + 0.055f) / 1.055f, 2.4f); }

static Col aces(Col p) { /* tonemap.cuh:9-35 */
  Col c = {0.59719f * p.r + 0.35458f * p.g + 0.04823f * p.b, 0.07600f * p.r + 0.90834f * p.g + 0.01566f * p.b,
           0.02840f * p.r + 0.13383f * p.g + 0.83777f * p.b};
  float ch[3] = {c.r, c.g, c.b};
  for (int k = 0; k < 3; k++) {
    const float a = ch[k] * (ch[k] + 0.0245786f) - 0.000090537f;
    const float b = ch[k] * (ch[k] * 0.983729f + 0.432951f) + 0.238081f;
    ch[k]         = a / b;
  }
  Col o = {1.60475f * ch[0] - 0.53108f * ch[1] - 0.07367f * ch[2], -0.10208f * ch[0] + 1.10813f * ch[1] - 0.00605f * ch[2],
           -0.00327f * ch[0] - 0.07276f * ch[1] + 1.07602f * ch[2]};
  return o;
}

static float u2(float x) { /* uncharted2_partial, tonemap.cuh:37-52 */
  const float a = 0.15f, b = 0.50f, c = 0.10f, d = 0.20f, e = 0.02f, f = 0.30f;
  return ((x * (a * x + c * b) + d * e) / (x * (a * x + b) + d * f)) - e / f;
}

static float agx_poly(float v) { /* tonemap.cuh:80-85 */
  const float v2 = v * v, v4 = v2 * v2;
  return 15.5f * v4 * v2 - 40.14f * v4 * v + 31.96f * v4 - 6.868f * v2 * v + 0.4298f * v2 + 0.1191f * v - 0.00232f;
}

static Col agx_forward(Col p) { /* agx_conversion, tonemap.cuh:95-121 */
  float a[3] = {0.842479062253094f * p.r + 0.0784335999999992f * p.g + 0.0792237451477643f * p.b,
                0.0423282422610123f * p.r + 0.878468636469772f * p.g + 0.0791661274605434f * p.b,
                0.0423756549057051f * p.r + 0.0784336f * p.g + 0.879142973793104f * p.b};
  const float lo = -12.47393f, hi = 4.026069f;
  for (int k = 0; k < 3; k++) {
    float v = log2f(fmaxf(a[k], 0.00017578139f));
    v       = fminf(fmaxf(v, lo), hi);
    a[k]    = agx_poly((v - lo) / (hi - lo));
  }
  Col o = {a[0], a[1], a[2]};
  return o;
}

static Col agx_inverse(Col p) { /* agx_inv_conversion, tonemap.cuh:123-141 */
  Col a = {1.19687900512017f * p.r - 0.0980208811401368f * p.g - 0.0990297440797205f * p.b,
           -0.0528968517574562f * p.r + 1.15190312990417f * p.g - 0.0989611768448433f * p.b,
           -0.0529716355144438f * p.r - 0.0980434501171241f * p.g + 1.15107367264116f * p.b};
  Col o = {to_linear(fmaxf(a.r, 0.0f)), to_linear(fmaxf(a.g, 0.0f)), to_linear(fmaxf(a.b, 0.0f))};
  return o;
}

static Col agx_look(Col p, float slope, float power, float sat) { /* tonemap.cuh:143-157 */
  const float l = lum(p);
  Col q         = {powf(p.r * slope, power), powf(p.g * slope, power), powf(p.b * slope, power)};
  Col o         = {l + sat * (q.r - l), l + sat * (q.g - l), l + sat * (q.b - l)};
  return o;
}

static Col purkinje(Col pixel, float kappa1, float kappa2) { /* purkinje_shift, cuda/purkinje.cuh:19-90 */
  const float strength = 5000.0f;
  if (lum(pixel) >= (1.0f / strength))
    return pixel;
  const float lc  = 0.096869562190332f * pixel.r + 0.318940374720484f * pixel.g - 0.188428411786113f * pixel.b;
  const float mc  = 0.020208210904239f * pixel.r + 0.291385283197581f * pixel.g - 0.090918262127325f * pixel.b;
  const float sc  = 0.002760510899553f * pixel.r - 0.008341563564118f * pixel.g + 0.067213551661950f * pixel.b;
  const float rod = -0.007607045462440f * pixel.r + 0.122492925567539f * pixel.g + 0.022445835141881f * pixel.b;
  const float lm = 1.0f / 0.63721f, mm = 1.0f / 0.39242f, sm = 1.0f / 1.6064f;
  const float eps = 1.1920929e-7f;
  const float sr  = 1.0f / sqrtf(fmaxf(1.0f + (1.0f / 3.0f) * lm * (lc + kappa1 * rod), eps));
  const float sg  = 1.0f / sqrtf(fmaxf(1.0f + (1.0f / 3.0f) * mm * (mc + kappa1 * rod), eps));
  const float sb  = 1.0f / sqrtf(fmaxf(1.0f + (1.0f / 3.0f) * sm * (sc + kappa2 * rod), eps));
  const float K = 45.0f, S = 10.0f, k3 = 0.6f, rw = 0.139f, p = 0.6189f;
  float o_r = ((-k3 - rw) * sr + (1.0f + k3 * rw) * sg) * kappa1 * lm;
  float o_g = (p * k3 * sr + (1.0f - p) * k3 * sg + sb) * kappa1 * mm;
  float o_b = (p * S * sr + (1.0f - p) * S * sg) * kappa2 * sm;
  const float f = (K / S) * rod;
  o_r *= f, o_g *= f, o_b *= f;
  const float L = lc + 0.5f * (o_b - o_r), M = mc + 0.5f * (o_b + o_r), Sx = sc + o_g + o_b;
  const float X = 1.9102f * L - 1.1121f * M + 0.2019f * Sx, Y = 0.3710f * L + 0.6291f * M, Z = Sx;
  Col rgb = {3.2405f * X - 1.5371f * Y - 0.4985f * Z, -0.9693f * X + 1.876f * Y + 0.0416f * Z, 0.0556f * X - 0.2040f * Y + 1.0572f * Z};
  float blend = fminf(fmaxf(1.0f - strength * lum(pixel), 0.0f), 1.0f);
  blend *= blend;
  Col o = {pixel.r * (1.0f - blend) + rgb.r * blend, pixel.g * (1.0f - blend) + rgb.g * blend, pixel.b * (1.0f - blend) + rgb.b * blend};
  return o;
}

static Col tonemap_pixel(Col p, float exposure, uint32_t tonemap, float agx_slope, float agx_power, float agx_saturation, int use_purkinje,
                         float kappa1, float kappa2) { /* tonemap_apply, cuda/tonemap.cuh:205-246 */
  if (use_purkinje)
    p = purkinje(p, kappa1, kappa2);
  p.r = fmaxf(p.r * exposure, 0.0f), p.g = fmaxf(p.g * exposure, 0.0f), p.b = fmaxf(p.b * exposure, 0.0f);
  switch (tonemap) {
    case 1: p = aces(p); break;
    case 2: {
      const float f = 1.0f / (1.0f + lum(p));
      p.r *= f, p.g *= f, p.b *= f;
    } break;
    case 3: {
      const float s = 1.0f / u2(11.2f);
      p.r = u2(2.0f * p.r) * s, p.g = u2(2.0f * p.g) * s, p.b = u2(2.0f * p.b) * s;
    } break;
    case 4: p = agx_inverse(agx_forward(p)); break;
    case 5: p = agx_inverse(agx_look(agx_forward(p), 1.0f, 1.35f, 1.4f)); break;
    case 6: p = agx_inverse(agx_look(agx_forward(p), agx_slope, agx_power, agx_saturation)); break;
    default: break;
  }
  return p;
}

static Col rgb_to_hsv(Col rgb) { /* math.cuh:1483-1511 */
  const float mx = fmaxf(rgb.r, fmaxf(rgb.g, rgb.b)), mn = fminf(rgb.r, fminf(rgb.g, rgb.b));
  const float s  = (mx - mn) / mx;
  float h        = 0.0f;
  if (s != 0.0f) {
    const float delta = mx - mn;
    if (mx == rgb.r)
      h = (rgb.g - rgb.b) / delta;
    else if (mx == rgb.g)
      h = 2.0f + (rgb.b - rgb.r) / delta;
    else
      h = 4.0f + (rgb.r - rgb.g) / delta;
    h *= 1.0f / 6.0f;
    if (h < 0.0f)
      h += 1.0f;
  }
  Col o = {h, s, mx};
  return o;
}
static float sat01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
static Col hsv_to_rgb(Col hsv) { /* math.cuh:1516-1541 */
  const float s = hsv.g, v = hsv.b;
  if (s == 0.0f) {
    Col g = {v, v, v};
    return g;
  }
  const float h = hsv.r * 6.0f;
  const float hr = sat01(fabsf(fmodf(h, 6.0f) - 3.0f) - 1.0f), hg = sat01(fabsf(fmodf(h + 4.0f, 6.0f) - 3.0f) - 1.0f),
              hb = sat01(fabsf(fmodf(h + 2.0f, 6.0f) - 3.0f) - 1.0f);
  Col o = {((1.0f - s) + hr * s) * v, ((1.0f - s) + hg * s) * v, ((1.0f - s) + hb * s) * v};
  return o;
}
uint16_t orc_squares16(uint32_t key, uint32_t counter);
static float white_noise_offset(uint32_t offset) { /* random.cuh:297-307 */
  union {
    uint32_t u;
    float f;
  } c;
  c.u = 0x3F800000u | ((uint32_t) orc_squares16(0xfcbd6e15u, offset) << 7);
  return c.f - 1.0f;
}

/* the whole output chain incl. colour correction, film grain (tonemap_apply, tonemap.cuh:205-246) and the image filters
 * (convert_RGBF_to_ARGB8, kernels.cuh:615-637; math.cuh:1081-1168) */
void orc_output_argb8_full(const float* planes, uint32_t width, uint32_t height, uint32_t sample_count, const OrcOutputParams* op,
                           const uint16_t* bluenoise_1d, uint8_t* dst) {
  const size_t n       = (size_t) width * height;
  const float norm     = 1.0f / (float) sample_count;
  const uint32_t scale = 1u << op->supersampling;
  const uint32_t ow = width >> op->supersampling, oh = height >> op->supersampling;
  for (uint32_t y = 0; y < oh; y++) {
    for (uint32_t x = 0; x < ow; x++) {
      Col acc = {0.0f, 0.0f, 0.0f};
      for (uint32_t yi = 0; yi < scale; yi++)
        for (uint32_t xi = 0; xi < scale; xi++) {
          const uint32_t px = (x * scale + xi < width) ? x * scale + xi : width - 1;
          const uint32_t py = (y * scale + yi < height) ? y * scale + yi : height - 1;
          const size_t k    = px + (size_t) py * width;
          Col p             = {planes[k] * norm, planes[n + k] * norm, planes[2 * n + k] * norm};
          if (op->purkinje)
            p = purkinje(p, op->purkinje_kappa1, op->purkinje_kappa2);
          if (op->use_color_correction) {
            Col hsv = rgb_to_hsv(p);
            hsv.r += op->color_correction[0], hsv.g += op->color_correction[1], hsv.b += op->color_correction[2];
            if (hsv.r < 0.0f)
              hsv.r += 1.0f;
            if (hsv.r > 1.0f)
              hsv.r -= 1.0f;
            hsv.g = sat01(hsv.g);
            if (hsv.b < 0.0f)
              hsv.b = 0.0f;
            p = hsv_to_rgb(hsv);
          }
          p.r *= op->exposure, p.g *= op->exposure, p.b *= op->exposure;
          const float grain = op->film_grain * (white_noise_offset(px + py * width) - 0.5f);
          p.r = fmaxf(0.0f, p.r + grain), p.g = fmaxf(0.0f, p.g + grain), p.b = fmaxf(0.0f, p.b + grain);
          p = tonemap_pixel(p, 1.0f, op->tonemap, op->agx_slope, op->agx_power, op->agx_saturation, 0, 0.0f, 0.0f);
          acc.r += p.r, acc.g += p.g, acc.b += p.b;
        }
      const float inv = 1.0f / (float) (scale * scale);
      acc.r *= inv, acc.g *= inv, acc.b *= inv;
      float mask = 0.5f;
      if (bluenoise_1d) {
        union {
          uint32_t u;
          float f;
        } c;
        c.u  = 0x3F800000u | ((uint32_t) bluenoise_1d[(x & 0xFFu) + (y & 0xFFu) * 256u] << 7);
        mask = c.f - 1.0f;
      }
      switch (op->filter) {
        case 1: {
          const float v = lum(acc);
          acc.r = acc.g = acc.b = v;
        } break;
        case 2: {
          const Col q = {acc.r * 0.393f + acc.g * 0.769f + acc.b * 0.189f, acc.r * 0.349f + acc.g * 0.686f + acc.b * 0.168f,
                         acc.r * 0.272f + acc.g * 0.534f + acc.b * 0.131f};
          acc         = q;
        } break;
        case 3: {
          const int tone = (int) (4.0f * lum(acc) + mask);
          const Col t0 = {15.0f / 255.0f, 56.0f / 255.0f, 15.0f / 255.0f}, t1 = {48.0f / 255.0f, 98.0f / 255.0f, 48.0f / 255.0f},
                    t2 = {139.0f / 255.0f, 172.0f / 255.0f, 15.0f / 255.0f}, t3 = {155.0f / 255.0f, 188.0f / 255.0f, 15.0f / 255.0f};
          acc = (tone == 0) ? t0 : (tone == 1) ? t1 : (tone == 2) ? t2 : t3;
        } break;
        case 4: {
          const int tone = (int) (4.0f * lum(acc) + mask);
          const float v  = (tone == 0) ? 0.0f : (tone == 1) ? 1.0f / 3.0f : (tone == 2) ? 2.0f / 3.0f : 1.0f;
          acc.r = acc.g = acc.b = v;
        } break;
        case 5: {
          acc.r *= 1.5f, acc.g *= 1.5f, acc.b *= 1.5f;
          const uint32_t row = y % 3u;
          if (row == 0)
            acc.r = acc.g = 0.0f;
          else if (row == 1)
            acc.g = acc.b = 0.0f;
          else
            acc.r = acc.b = 0.0f;
        } break;
        case 6: {
          const int tone = (int) (2.0f * lum(acc) + mask);
          acc.r = acc.g = acc.b = (tone == 0) ? 0.0f : 1.0f;
        } break;
        default: break;
      }
      const float dither = op->dithering ? mask : 0.5f;
      const size_t i     = x + (size_t) y * ow;
      dst[4 * i + 0]     = (uint8_t) fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * to_srgb(acc.b)));
      dst[4 * i + 1]     = (uint8_t) fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * to_srgb(acc.g)));
      dst[4 * i + 2]     = (uint8_t) fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * to_srgb(acc.r)));
      dst[4 * i + 3]     = 0xFFu;
    }
  }
}

/* width / height: internal (rendered) resolution; the image written is (width >> supersampling) x (height >> supersampling) */
void orc_output_argb8_ex(const float* planes, uint32_t width, uint32_t height, uint32_t sample_count, float exposure, uint32_t tonemap,
                         float agx_slope, float agx_power, float agx_saturation, const uint16_t* bluenoise_1d, int use_purkinje, float kappa1,
                         float kappa2, uint32_t supersampling, uint8_t* dst) {
  const size_t n       = (size_t) width * height;
  const float norm     = 1.0f / (float) sample_count;
  const uint32_t scale = 1u << supersampling;
  const uint32_t ow = width >> supersampling, oh = height >> supersampling;
  for (uint32_t y = 0; y < oh; y++) {
    for (uint32_t x = 0; x < ow; x++) {
      Col acc = {0.0f, 0.0f, 0.0f};
      for (uint32_t yi = 0; yi < scale; yi++)
        for (uint32_t xi = 0; xi < scale; xi++) {
          const uint32_t px = (x * scale + xi < width) ? x * scale + xi : width - 1;
          const uint32_t py = (y * scale + yi < height) ? y * scale + yi : height - 1;
          const size_t k    = px + (size_t) py * width;
          Col p             = {planes[k] * norm, planes[n + k] * norm, planes[2 * n + k] * norm};
          p                 = tonemap_pixel(p, exposure, tonemap, agx_slope, agx_power, agx_saturation, use_purkinje, kappa1, kappa2);
          acc.r += p.r, acc.g += p.g, acc.b += p.b;
        }
      const float inv = 1.0f / (float) (scale * scale);
      acc.r *= inv, acc.g *= inv, acc.b *= inv;
      float dither = 0.5f;
      if (bluenoise_1d) {
        union {
          uint32_t u;
          float f;
        } c;
        c.u    = 0x3F800000u | ((uint32_t) bluenoise_1d[(x & 0xFFu) + (y & 0xFFu) * 256u] << 7);
        dither = c.f - 1.0f;
      }
      const size_t i = x + (size_t) y * ow;
      dst[4 * i + 0] = (uint8_t) fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * to_srgb(acc.b)));
      dst[4 * i + 1] = (uint8_t) fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * to_srgb(acc.g)));
      dst[4 * i + 2] = (uint8_t) fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * to_srgb(acc.r)));
      dst[4 * i + 3] = 0xFFu;
    }
  }
}

void orc_output_argb8(const float* planes, uint32_t width, uint32_t height, uint32_t sample_count, float exposure, uint32_t tonemap,
                      float agx_slope, float agx_power, float agx_saturation, const uint16_t* bluenoise_1d, uint8_t* dst) {
  orc_output_argb8_ex(planes, width, height, sample_count, exposure, tonemap, agx_slope, agx_power, agx_saturation, bluenoise_1d, 0, 0.0f, 0.0f, 0,
                      dst);
}

/* ------------------------------------------------------------------ */
/* bloom: _device_post_bloom_apply (device/device_post.c:62-140) with post_image_downsample / post_image_upsample        */
/* (cuda/post_common.cuh:8-143). Acts in place on the three mean-radiance planes, before the tone map. Multiply-adds of the */
/* accumulation chains are written as fmaf: that is what nvcc contracts them to under --use_fast_math.                    */
/* ------------------------------------------------------------------ */
static float post_sample(const float* buffer, float x, float y, uint32_t width, uint32_t height) {
  const float source_x = fmaxf(0.0f, x * (float) (width - 1));
  const float source_y = fmaxf(0.0f, y * (float) (height - 1));
  const uint32_t x0 = (uint32_t) source_x, y0 = (uint32_t) source_y;
  uint32_t x1 = (uint32_t) (source_x + 1.0f), y1 = (uint32_t) (source_y + 1.0f);
  if (x1 > width - 1)
    x1 = width - 1;
  if (y1 > height - 1)
    y1 = height - 1;
  const float p00 = buffer[x0 + y0 * width], p01 = buffer[x0 + y1 * width], p10 = buffer[x1 + y0 * width], p11 = buffer[x1 + y1 * width];
  const float fx = source_x - (float) x0, ifx = 1.0f - fx, fy = source_y - (float) y0, ify = 1.0f - fy;
  float result = p00 * (ifx * ify);
  result       = fmaf(p01, ifx * fy, result);
  result       = fmaf(p10, fx * ify, result);
  result       = fmaf(p11, fx * fy, result);
  return result;
}

static float post_sample_border(const float* image, float x, float y, uint32_t width, uint32_t height) {
  const float below_one = 0.99999994f; /* 0x3F7FFFFF */
  if (x > below_one || x < 0.0f || y > below_one || y < 0.0f)
    return 0.0f;
  return post_sample(image, x, y, width, height);
}

static void post_downsample(const float* src, uint32_t sw, uint32_t sh, float* dst, uint32_t tw, uint32_t th) {
  const float scale_x = 1.0f / (float) (tw - 1), scale_y = 1.0f / (float) (th - 1);
  const float step_x = 1.0f / (float) (sw - 1), step_y = 1.0f / (float) (sh - 1);
  for (uint32_t y = 0; y < th; y++)
    for (uint32_t x = 0; x < tw; x++) {
      const float sx = scale_x * (float) x, sy = scale_y * (float) y;
      const float hx = 0.5f * step_x, hy = 0.5f * step_y;
      float pixel = 0.0f;
      pixel += post_sample_border(src, sx - hx, sy - hy, sw, sh);
      pixel += post_sample_border(src, sx + hx, sy - hy, sw, sh);
      pixel += post_sample_border(src, sx - hx, sy + hy, sw, sh);
      pixel += post_sample_border(src, sx + hx, sy + hy, sw, sh);
      pixel += post_sample_border(src, sx, sy, sw, sh);
      pixel = fmaf(post_sample_border(src, sx, sy - step_y, sw, sh), 0.5f, pixel);
      pixel = fmaf(post_sample_border(src, sx - step_x, sy, sw, sh), 0.5f, pixel);
      pixel = fmaf(post_sample_border(src, sx + step_x, sy, sw, sh), 0.5f, pixel);
      pixel = fmaf(post_sample_border(src, sx, sy + step_y, sw, sh), 0.5f, pixel);
      pixel = fmaf(post_sample_border(src, sx - step_x, sy - step_y, sw, sh), 0.25f, pixel);
      pixel = fmaf(post_sample_border(src, sx + step_x, sy - step_y, sw, sh), 0.25f, pixel);
      pixel = fmaf(post_sample_border(src, sx - step_x, sy + step_y, sw, sh), 0.25f, pixel);
      pixel = fmaf(post_sample_border(src, sx + step_x, sy + step_y, sw, sh), 0.25f, pixel);
      pixel *= 1.0f / 8.0f;
      dst[x + y * tw] = fmaxf(pixel, 0.0f);
    }
}

static void post_upsample(const float* src, uint32_t sw, uint32_t sh, float* dst, uint32_t tw, uint32_t th, float sa, float sb) {
  const float scale_x = 1.0f / (float) (tw - 1), scale_y = 1.0f / (float) (th - 1);
  const float step_x = 1.0f / (float) (sw - 1), step_y = 1.0f / (float) (sh - 1);
  for (uint32_t y = 0; y < th; y++)
    for (uint32_t x = 0; x < tw; x++) {
      const float sx = scale_x * (float) x, sy = scale_y * (float) y;
      float pixel = post_sample_border(src, sx - step_x, sy - step_y, sw, sh);
      pixel       = fmaf(post_sample_border(src, sx, sy - step_y, sw, sh), 2.0f, pixel);
      pixel += post_sample_border(src, sx + step_x, sy - step_y, sw, sh);
      pixel = fmaf(post_sample_border(src, sx - step_x, sy, sw, sh), 2.0f, pixel);
      pixel = fmaf(post_sample_border(src, sx, sy, sw, sh), 4.0f, pixel);
      pixel = fmaf(post_sample_border(src, sx + step_x, sy, sw, sh), 2.0f, pixel);
      pixel += post_sample_border(src, sx - step_x, sy + step_y, sw, sh);
      pixel = fmaf(post_sample_border(src, sx, sy + step_y, sw, sh), 2.0f, pixel);
      pixel += post_sample_border(src, sx + step_x, sy + step_y, sw, sh);
      pixel *= 1.0f / 20.0f;
      pixel *= sa;
      dst[x + y * tw] = fmaf(dst[x + y * tw], sb, pixel); /* dst == base */
    }
}

/* rgb: three planes of width * height mean radiance, modified in place */
void orc_bloom_apply(float* rgb, uint32_t width, uint32_t height, float blend) {
  uint32_t min_dim = width < height ? width : height, mip_count = 0;
  if (min_dim == 0 || !(blend > 0.0f))
    return;
  while (min_dim != 1) {
    mip_count++;
    min_dim >>= 1;
  }
  if (mip_count <= 1)
    return;
  float** mips = (float**) malloc(sizeof(float*) * mip_count);
  for (uint32_t i = 0; i < mip_count; i++)
    mips[i] = (float*) malloc(sizeof(float) * (size_t) ((width >> (i + 1)) * (height >> (i + 1)) + 1));
  for (uint32_t ch = 0; ch < 3; ch++) {
    float* plane = rgb + (size_t) ch * width * height;
    post_downsample(plane, width, height, mips[0], width >> 1, height >> 1);
    for (uint32_t i = 0; i + 1 < mip_count; i++)
      post_downsample(mips[i], width >> (i + 1), height >> (i + 1), mips[i + 1], width >> (i + 2), height >> (i + 2));
    for (uint32_t i = mip_count - 1; i > 0; i--)
      post_upsample(mips[i], width >> (i + 1), height >> (i + 1), mips[i - 1], width >> i, height >> i, 1.0f, 1.0f);
    post_upsample(mips[0], width >> 1, height >> 1, plane, width, height, blend / (float) mip_count, 1.0f - blend);
  }
  for (uint32_t i = 0; i < mip_count; i++)
    free(mips[i]);
  free(mips);
}

/* exposure + tone-map transform of one colour (tonemap_apply_transform, cuda/tonemap.cuh:175-203), for orc_adaptive.c */
void orc_tonemap_rgb(float rgb[3], float exposure, uint32_t tonemap, float agx_slope, float agx_power, float agx_saturation) {
  Col p = {rgb[0], rgb[1], rgb[2]};
  p     = tonemap_pixel(p, exposure, tonemap, agx_slope, agx_power, agx_saturation, 0, 0.0f, 0.0f);
  rgb[0] = p.r, rgb[1] = p.g, rgb[2] = p.b;
}
