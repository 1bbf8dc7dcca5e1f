/*
 * luminary/api_utils.h - small value types and export macros (reference api_utils.h:23-51)
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/api_utils.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LUMINARY_API_UTILS_H
#define LUMINARY_API_UTILS_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#define LUMINARY_API
#define LUMINARY_DEPRECATED

typedef struct LuminaryVec3 { float x, y, z; } LuminaryVec3;
typedef struct LuminaryRGBF { float r, g, b; } LuminaryRGBF;
typedef struct LuminaryRGBAF { float r, g, b, a; } LuminaryRGBAF;
typedef struct LuminaryARGB8 { uint8_t b, g, r, a; } LuminaryARGB8; /* byte order of the output images */

#endif /* LUMINARY_API_UTILS_H */
