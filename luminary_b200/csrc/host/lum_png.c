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
