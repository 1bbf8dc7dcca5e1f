/* orc_output.c - CPU restatement of the reference's output chain for parity tests. TEST INFRASTRUCTURE ONLY.
 *
 * Follows generate_final_image + convert_RGBF_to_ARGB8 (device/cuda/kernels.cuh:503-644) with tonemap_apply
 * (device/cuda/tonemap.cuh:7-246), linearRGB_to_SRGB / SRGB_to_linearRGB (cuda/math.cuh:1044-1060), color_luminance
 * (:1062-1064) and random_dither_mask (cuda/random.cuh:150-154,370-375), for supersampling 0, undersampling 0, filter
 * NONE, purkinje / colour correction / film grain off. exposure is the linear factor expf(camera.exposure)
 * (device_structs.c:77). Parity unpinned against a running reference (GPU-only); the formulas are closed form. */
#include <math.h>
#include <stdint.h>

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

void orc_output_argb8(const float* planes, uint32_t width, uint32_t height, uint32_t sample_count, float exposure, uint32_t tonemap,
                      float agx_slope, float agx_power, float agx_saturation, const uint16_t* bluenoise_1d, uint8_t* dst) {
  const size_t n   = (size_t) width * height;
  const float norm = 1.0f / (float) sample_count;
  for (uint32_t y = 0; y < height; y++) {
    for (uint32_t x = 0; x < width; x++) {
      const size_t i = x + (size_t) y * width;
      Col p          = {fmaxf(planes[i] * norm * exposure, 0.0f), fmaxf(planes[n + i] * norm * exposure, 0.0f),
                        fmaxf(planes[2 * n + i] * norm * exposure, 0.0f)};
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
      float dither = 0.5f;
      if (bluenoise_1d) {
        union {
          uint32_t u;
          float f;
        } c;
        c.u    = 0x3F800000u | ((uint32_t) bluenoise_1d[(x & 0xFFu) + (y & 0xFFu) * 256u] << 7);
        dither = c.f - 1.0f;
      }
      dst[4 * i + 0] = (uint8_t) fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * to_srgb(p.b)));
      dst[4 * i + 1] = (uint8_t) fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * to_srgb(p.g)));
      dst[4 * i + 2] = (uint8_t) fmaxf(0.0f, fminf(255.9999f, dither + 255.0f * to_srgb(p.r)));
      dst[4 * i + 3] = 0xFFu;
    }
  }
}
