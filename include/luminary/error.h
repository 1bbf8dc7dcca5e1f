/*
 * luminary/error.h - result codes (reference error.h:22-101)
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/error.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LUMINARY_API_ERROR_H
#define LUMINARY_API_ERROR_H

#include <luminary/api_utils.h>
#include <stdint.h>

/* 0 = success, 1..13 = error class, bit 63 = the error was propagated from an inner call */
typedef uint64_t LuminaryResult;

#define LUMINARY_SUCCESS (0ull)
#define LUMINARY_ERROR_ARGUMENT_NULL (1ull)         /* non-optional argument was NULL */
#define LUMINARY_ERROR_NOT_IMPLEMENTED (2ull)       /* feature outside the path this library serves */
#define LUMINARY_ERROR_INVALID_API_ARGUMENT (3ull)
#define LUMINARY_ERROR_MEMORY_LEAK (4ull)
#define LUMINARY_ERROR_OUT_OF_MEMORY (5ull)
#define LUMINARY_ERROR_C_STD (6ull)
#define LUMINARY_ERROR_API_EXCEPTION (7ull)         /* API used in a non-compliant way */
#define LUMINARY_ERROR_CUDA (8ull)
#define LUMINARY_ERROR_OPTIX (9ull)                 /* never produced here: there is no OptiX on this path */
#define LUMINARY_ERROR_PREVIOUS_ERROR (10ull)
#define LUMINARY_ERROR_DEBUG_ASSERT (11ull)
#define LUMINARY_ERROR_MISSING_DATA (12ull)
#define LUMINARY_ERROR_INVALID_DEVICE (13ull)
#define LUMINARY_ERROR_PROPAGATED (0x8000000000000000ull)

#ifdef __cplusplus
extern "C" {
#endif
LUMINARY_API const char* luminary_result_to_string(LuminaryResult result);
#ifdef __cplusplus
}
#endif

#endif /* LUMINARY_API_ERROR_H */
