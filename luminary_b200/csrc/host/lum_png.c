/* lum_png.c - minimal PNG writer for luminary_host_save_png (reference host/png.c:167-412,754-784 writes 8-bit RGBA
 * through zlib; this one emits stored deflate blocks, which every decoder accepts, and needs no zlib). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lum_host_internal.h"

static uint32_t crc_table[256];
static int crc_ready = 0;

static void crc_init(void) {
  for (uint32_t n = 0; n < 256; n++) {
    uint32_t c = n;
    for (int k = 0; k < 8; k++)
      c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    crc_table[n] = c;
  }
  crc_ready = 1;
}

static uint32_t crc_update(uint32_t crc, const uint8_t* buf, size_t len) {
  for (size_t i = 0; i < len; i++)
    crc = crc_table[(crc ^ buf[i]) & 0xFFu] ^ (crc >> 8);
  return crc;
}

static void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t) (v >> 24), p[1] = (uint8_t) (v >> 16), p[2] = (uint8_t) (v >> 8), p[3] = (uint8_t) v; }

static int write_chunk(FILE* f, const char type[4], const uint8_t* data, uint32_t len) {
  uint8_t hdr[8];
  put_be32(hdr, len);
  memcpy(hdr + 4, type, 4);
  uint32_t crc = crc_update(0xFFFFFFFFu, hdr + 4, 4);
  if (len)
    crc = crc_update(crc, data, len);
  uint8_t tail[4];
  put_be32(tail, crc ^ 0xFFFFFFFFu);
  return fwrite(hdr, 1, 8, f) == 8 && (!len || fwrite(data, 1, len, f) == len) && fwrite(tail, 1, 4, f) == 4;
}

LuminaryResult lum_png_write_argb8(const char* path, const uint8_t* argb8, uint32_t width, uint32_t height, size_t ld) {
  LUM_CHECK_NULL(path);
  LUM_CHECK_NULL(argb8);
  if (!width || !height)
    LUM_RETURN_ERROR(LUMINARY_ERROR_INVALID_API_ARGUMENT, "cannot store an empty image");
  if (!crc_ready)
    crc_init();

  /* raw scanlines: filter byte 0 + RGBA */
  const size_t row = 1 + 4 * (size_t) width;
  const size_t raw = row * height;
  /* zlib stream: 2 header bytes, stored blocks of <= 65535 bytes (5 bytes of header each), adler32 */
  const size_t blocks = (raw + 65534) / 65535;
  const size_t zsize  = 2 + raw + 5 * blocks + 4;
  uint8_t* z          = (uint8_t*) malloc(zsize);
  uint8_t* scan       = (uint8_t*) malloc(raw);
  if (!z || !scan) {
    free(z), free(scan);
    LUM_RETURN_ERROR(LUMINARY_ERROR_OUT_OF_MEMORY, "out of host memory writing %s", path);
  }
  for (uint32_t y = 0; y < height; y++) {
    uint8_t* d = scan + row * y;
    *d++       = 0;
    for (uint32_t x = 0; x < width; x++) {
      const uint8_t* s = argb8 + 4 * (x + (size_t) y * ld); /* b, g, r, a */
      *d++ = s[2], *d++ = s[1], *d++ = s[0], *d++ = s[3];
    }
  }
  size_t o = 0;
  z[o++]   = 0x78, z[o++] = 0x01;
  uint32_t a = 1, b = 0;
  for (size_t off = 0; off < raw; off += 65535) {
    const size_t n = (raw - off < 65535) ? raw - off : 65535;
    z[o++]         = (off + n >= raw) ? 1 : 0;
    z[o++] = (uint8_t) (n & 0xFF), z[o++] = (uint8_t) (n >> 8);
    z[o++] = (uint8_t) (~n & 0xFF), z[o++] = (uint8_t) ((~n >> 8) & 0xFF);
    memcpy(z + o, scan + off, n);
    o += n;
    for (size_t i = 0; i < n; i++) {
      a = (a + scan[off + i]) % 65521u;
      b = (b + a) % 65521u;
    }
  }
  put_be32(z + o, (b << 16) | a);
  o += 4;

  FILE* f = fopen(path, "wb");
  if (!f) {
    free(z), free(scan);
    LUM_RETURN_ERROR(LUMINARY_ERROR_C_STD, "File %s could not be opened for writing.", path);
  }
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  uint8_t ihdr[13];
  put_be32(ihdr, width);
  put_be32(ihdr + 4, height);
  ihdr[8] = 8, ihdr[9] = 6, ihdr[10] = 0, ihdr[11] = 0, ihdr[12] = 0;
  int ok = fwrite(sig, 1, 8, f) == 8 && write_chunk(f, "IHDR", ihdr, 13) && write_chunk(f, "IDAT", z, (uint32_t) o) && write_chunk(f, "IEND", NULL, 0);
  fclose(f);
  free(z), free(scan);
  if (!ok)
    LUM_RETURN_ERROR(LUMINARY_ERROR_C_STD, "Failed to write %s", path);
  return LUMINARY_SUCCESS;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* PNG reader for scene textures (reference host/png.c:415-712): 8 / 16 bit, grayscale / grayscale-alpha /         */
/* truecolor / truecolor-alpha, not interlaced, expanded to RGBA8 / RGBA16 (little endian); gAMA -> gamma.         */
/* The reference inflates with zlib; this is a self-contained RFC 1951 decoder.                                     */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
  const uint8_t* src;
  size_t len, pos;
  uint32_t bitbuf;
  int bitcnt;
  uint8_t* out;
  size_t out_len, out_pos;
  int error;
} Inflate;

static uint32_t inf_bits(Inflate* s, int n) {
  while (s->bitcnt < n) {
    if (s->pos >= s->len) {
      s->error = 1;
      return 0;
    }
    s->bitbuf |= (uint32_t) s->src[s->pos++] << s->bitcnt;
    s->bitcnt += 8;
  }
  const uint32_t v = s->bitbuf & ((n == 32) ? 0xFFFFFFFFu : ((1u << n) - 1u));
  s->bitbuf >>= n;
  s->bitcnt -= n;
  return v;
}

typedef struct {
  uint16_t count[16];
  uint16_t symbol[288];
} Huffman;

static void huff_build(Huffman* h, const uint8_t* lengths, int n) {
  uint16_t offs[16];
  memset(h->count, 0, sizeof(h->count));
  for (int i = 0; i < n; i++)
    h->count[lengths[i]]++;
  h->count[0] = 0;
  offs[1]     = 0;
  for (int i = 1; i < 15; i++)
    offs[i + 1] = offs[i] + h->count[i];
  for (int i = 0; i < n; i++)
    if (lengths[i])
      h->symbol[offs[lengths[i]]++] = (uint16_t) i;
}

static int huff_decode(Inflate* s, const Huffman* h) {
  int code = 0, first = 0, index = 0;
  for (int len = 1; len <= 15; len++) {
    code |= (int) inf_bits(s, 1);
    if (s->error)
      return -1;
    const int count = h->count[len];
    if (code - count < first)
      return h->symbol[index + (code - first)];
    index += count;
    first += count;
    first <<= 1;
    code <<= 1;
  }
  s->error = 1;
  return -1;
}

static void inf_codes(Inflate* s, const Huffman* lit, const Huffman* dist) {
  static const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  static const uint16_t lext[29]  = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  static const uint16_t dbase[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,   33,   49,   65,    97,    129,
                                     193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  static const uint16_t dext[30]  = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  while (!s->error) {
    int sym = huff_decode(s, lit);
    if (sym < 0)
      return;
    if (sym < 256) {
      if (s->out_pos >= s->out_len) {
        s->error = 1;
        return;
      }
      s->out[s->out_pos++] = (uint8_t) sym;
    }
    else if (sym == 256)
      return;
    else {
      sym -= 257;
      if (sym >= 29) {
        s->error = 1;
        return;
      }
      const uint32_t len = lbase[sym] + inf_bits(s, lext[sym]);
      const int ds       = huff_decode(s, dist);
      if (ds < 0 || ds >= 30) {
        s->error = 1;
        return;
      }
      const uint32_t d = dbase[ds] + inf_bits(s, dext[ds]);
      if (d > s->out_pos || s->out_pos + len > s->out_len) {
        s->error = 1;
        return;
      }
      for (uint32_t i = 0; i < len; i++, s->out_pos++)
        s->out[s->out_pos] = s->out[s->out_pos - d];
    }
  }
}

/* zlib stream (2-byte header, deflate blocks, adler32) -> exactly out_len bytes. Returns 0 on success. */
static int zlib_inflate(const uint8_t* src, size_t len, uint8_t* out, size_t out_len) {
  if (len < 6 || (src[0] & 0x0F) != 8 || ((src[0] << 8) | src[1]) % 31 != 0 || (src[1] & 0x20))
    return 1;
  Inflate s;
  memset(&s, 0, sizeof(s));
  s.src = src + 2, s.len = len - 2, s.out = out, s.out_len = out_len;
  int last = 0;
  while (!last && !s.error) {
    last           = (int) inf_bits(&s, 1);
    const int type = (int) inf_bits(&s, 2);
    if (type == 0) {
      s.bitbuf = 0, s.bitcnt = 0;
      if (s.pos + 4 > s.len)
        return 1;
      const uint32_t n = s.src[s.pos] | (s.src[s.pos + 1] << 8), nn = s.src[s.pos + 2] | (s.src[s.pos + 3] << 8);
      s.pos += 4;
      if ((n ^ 0xFFFFu) != nn || s.pos + n > s.len || s.out_pos + n > s.out_len)
        return 1;
      memcpy(s.out + s.out_pos, s.src + s.pos, n);
      s.pos += n, s.out_pos += n;
    }
    else if (type == 1) {
      uint8_t l[288];
      Huffman lit, dist;
      for (int i = 0; i < 144; i++) l[i] = 8;
      for (int i = 144; i < 256; i++) l[i] = 9;
      for (int i = 256; i < 280; i++) l[i] = 7;
      for (int i = 280; i < 288; i++) l[i] = 8;
      huff_build(&lit, l, 288);
      for (int i = 0; i < 30; i++) l[i] = 5;
      huff_build(&dist, l, 30);
      inf_codes(&s, &lit, &dist);
    }
    else if (type == 2) {
      static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      const int nlen = (int) inf_bits(&s, 5) + 257, ndist = (int) inf_bits(&s, 5) + 1, ncode = (int) inf_bits(&s, 4) + 4;
      if (nlen > 286 || ndist > 30)
        return 1;
      uint8_t l[320];
      memset(l, 0, sizeof(l));
      for (int i = 0; i < ncode; i++)
        l[order[i]] = (uint8_t) inf_bits(&s, 3);
      Huffman lencode, lit, dist;
      huff_build(&lencode, l, 19);
      memset(l, 0, sizeof(l));
      int idx = 0;
      while (idx < nlen + ndist && !s.error) {
        const int sym = huff_decode(&s, &lencode);
        if (sym < 0)
          return 1;
        if (sym < 16)
          l[idx++] = (uint8_t) sym;
        else {
          int rep, val = 0;
          if (sym == 16) {
            if (idx == 0)
              return 1;
            val = l[idx - 1];
            rep = 3 + (int) inf_bits(&s, 2);
          }
          else if (sym == 17)
            rep = 3 + (int) inf_bits(&s, 3);
          else
            rep = 11 + (int) inf_bits(&s, 7);
          if (idx + rep > nlen + ndist)
            return 1;
          while (rep--)
            l[idx++] = (uint8_t) val;
        }
      }
      huff_build(&lit, l, nlen);
      huff_build(&dist, l + nlen, ndist);
      inf_codes(&s, &lit, &dist);
    }
    else
      return 1;
  }
  return (s.error || s.out_pos != out_len) ? 1 : 0;
}

static uint32_t get_be32(const uint8_t* p) { return ((uint32_t) p[0] << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3]; }

static uint8_t paeth(uint8_t a, uint8_t b, uint8_t c) { /* reference png.c:367-383 */
  const int p = (int) a + (int) b - (int) c;
  const int pa = abs(p - (int) a), pb = abs(p - (int) b), pc = abs(p - (int) c);
  return (pa <= pb && pa <= pc) ? a : ((pb <= pc) ? b : c);
}

void lum_host_texture_free(LumHostTexture* tex) {
  if (tex)
    free(tex->data);
  if (tex)
    memset(tex, 0, sizeof(*tex));
}

LuminaryResult lum_png_read(const char* path, LumHostTexture* tex) {
  LUM_CHECK_NULL(path);
  LUM_CHECK_NULL(tex);
  memset(tex, 0, sizeof(*tex));
  tex->gamma = 1.0f; /* texture_create, texture.c:86 */
  if (!crc_ready)
    crc_init();
  FILE* f = fopen(path, "rb");
  if (!f)
    LUM_RETURN_ERROR(LUMINARY_ERROR_API_EXCEPTION, "Texture %s could not be opened.", path);
  long flen = -1; /* an unseekable path (FIFO, device) is treated like an empty file */
  if (fseek(f, 0, SEEK_END) == 0) {
    flen = ftell(f);
    if (fseek(f, 0, SEEK_SET) != 0)
      flen = -1;
  }
  uint8_t* file = (flen > 0) ? (uint8_t*) malloc((size_t) flen) : NULL;
  const bool read_ok = file && fread(file, 1, (size_t) flen, f) == (size_t) flen;
  fclose(f);
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  LuminaryResult result       = LUMINARY_SUCCESS;
  uint8_t *z = NULL, *raw = NULL;
#define PNG_FAIL(...)                               \
  do {                                              \
    lum_set_error(__VA_ARGS__);                     \
    result = LUMINARY_ERROR_API_EXCEPTION;          \
    goto done;                                      \
  } while (0)
  if (!read_ok || flen < 8 + 25 || memcmp(file, sig, 8))
    PNG_FAIL("File header does not correspond to png! (%s)", path);
  const uint8_t* ihdr = file + 8;
  if (get_be32(ihdr) != 13u || memcmp(ihdr + 4, "IHDR", 4))
    PNG_FAIL("Error in IHDR block. (%s)", path);
  if ((crc_update(0xFFFFFFFFu, ihdr + 4, 17) ^ 0xFFFFFFFFu) != get_be32(ihdr + 21))
    PNG_FAIL("Texture %s is corrupted!", path);
  const uint32_t width = get_be32(ihdr + 8), height = get_be32(ihdr + 12);
  const uint8_t depth = ihdr[16], ctype = ihdr[17], interlace = ihdr[20];
  if (ctype != 0 && ctype != 2 && ctype != 4 && ctype != 6)
    PNG_FAIL("Texture %s is either using a color palette or a non standard format!", path);
  if (depth != 8 && depth != 16)
    PNG_FAIL("Texture %s does not have 8 or 16 bit depth!", path);
  if (interlace)
    PNG_FAIL("Texture %s is interlaced, which is not supported.", path);
  if (!width || !height || width > 65535 || height > 65535)
    PNG_FAIL("Texture %s has an invalid size.", path);
  {
    const uint32_t channels = (ctype == 0) ? 1 : (ctype == 4) ? 2 : (ctype == 2) ? 3 : 4;
    const uint32_t bpc      = depth / 8;
    const uint32_t bpp      = channels * bpc;
    const size_t stride     = (size_t) width * bpp;
    const size_t raw_len    = (stride + 1) * height;
    z                       = (uint8_t*) malloc((size_t) flen);
    raw                     = (uint8_t*) malloc(raw_len);
    if (!z || !raw) {
      result = LUMINARY_ERROR_OUT_OF_MEMORY;
      lum_set_error("out of host memory reading %s", path);
      goto done;
    }
    size_t zlen = 0, off = 8 + 25;
    while (off + 12 <= (size_t) flen) {
      const uint32_t len = get_be32(file + off);
      if (off + 12 + (size_t) len > (size_t) flen)
        PNG_FAIL("Texture %s is truncated.", path);
      const uint8_t* type = file + off + 4;
      const bool crc_ok   = (crc_update(0xFFFFFFFFu, type, 4 + (size_t) len) ^ 0xFFFFFFFFu) == get_be32(file + off + 8 + len);
      if (!memcmp(type, "IDAT", 4)) {
        if (!crc_ok)
          PNG_FAIL("CRC Error. (%s)", path);
        memcpy(z + zlen, type + 4, len);
        zlen += len;
      }
      else if (!memcmp(type, "gAMA", 4)) {
        if (len != 4 || !crc_ok)
          lum_log("error", "Texture %s has a broken gAMA chunk. Ignoring it.", path);
        else
          tex->gamma = 100000.0f / ((float) get_be32(type + 4)); /* png.c:541 */
      }
      else if (!memcmp(type, "IEND", 4))
        break;
      off += 12 + (size_t) len;
    }
    if (zlib_inflate(z, zlen, raw, raw_len))
      PNG_FAIL("Texture %s: the compressed image data is invalid.", path);
    /* undo the scanline filters (png.c:306-413) */
    for (uint32_t y = 0; y < height; y++) {
      uint8_t* line       = raw + (stride + 1) * y + 1;
      const uint8_t* prev = y ? raw + (stride + 1) * (y - 1) + 1 : NULL;
      const uint8_t ft    = line[-1];
      for (size_t i = 0; i < stride; i++) {
        const uint8_t a = (i >= bpp) ? line[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= bpp) ? prev[i - bpp] : 0;
        switch (ft) {
          case 1: line[i] = (uint8_t) (line[i] + a); break;
          case 2: line[i] = (uint8_t) (line[i] + b); break;
          case 3: line[i] = (uint8_t) (line[i] + (uint8_t) (((int) a + (int) b) >> 1)); break;
          case 4: line[i] = (uint8_t) (line[i] + paeth(a, b, c)); break;
          default: break;
        }
      }
    }
    /* expand to four components (png.c:613-705) */
    const size_t px = (size_t) width * height;
    tex->data       = malloc(px * 4 * bpc);
    if (!tex->data) {
      result = LUMINARY_ERROR_OUT_OF_MEMORY;
      lum_set_error("out of host memory reading %s", path);
      goto done;
    }
    for (uint32_t y = 0; y < height; y++) {
      const uint8_t* line = raw + (stride + 1) * y + 1;
      for (uint32_t x = 0; x < width; x++) {
        uint16_t ch[4] = {0, 0, 0, (uint16_t) (bpc == 1 ? 255 : 65535)};
        uint16_t in[4] = {0, 0, 0, 0};
        for (uint32_t k = 0; k < channels; k++) {
          const uint8_t* p = line + (size_t) x * bpp + (size_t) k * bpc;
          in[k]            = (bpc == 1) ? p[0] : (uint16_t) ((p[0] << 8) | p[1]);
        }
        if (channels <= 2) {
          ch[0] = ch[1] = ch[2] = in[0];
          if (channels == 2)
            ch[3] = in[1];
        }
        else {
          ch[0] = in[0], ch[1] = in[1], ch[2] = in[2];
          if (channels == 4)
            ch[3] = in[3];
        }
        const size_t o = ((size_t) y * width + x) * 4;
        for (int k = 0; k < 4; k++) {
          if (bpc == 1)
            ((uint8_t*) tex->data)[o + k] = (uint8_t) ch[k];
          else
            ((uint16_t*) tex->data)[o + k] = ch[k];
        }
      }
    }
    tex->width          = width;
    tex->height         = height;
    tex->type           = (bpc == 1) ? LUMB200_TEXTURE_U8 : LUMB200_TEXTURE_U16;
    tex->num_components = 4;
    tex->pitch          = width * 4 * bpc;
    lum_log("log", "PNG (%s) Size: %ux%u Depth: %u Colortype: %u", path, width, height, depth, ctype);
  }
done:
#undef PNG_FAIL
  free(file), free(z), free(raw);
  if (result != LUMINARY_SUCCESS)
    lum_host_texture_free(tex);
  return result;
}
