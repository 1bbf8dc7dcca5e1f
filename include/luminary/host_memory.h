/*
 * luminary/host_memory.h - tracked host allocations (reference host_memory.h:22-30)
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/host_memory.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LUMINARY_API_HOST_MEMORY_H
#define LUMINARY_API_HOST_MEMORY_H

#include <luminary/api_utils.h>
#include <luminary/error.h>

#define host_malloc(ptr, size) _host_malloc((void**) (ptr), (size), (const char*) #ptr, (const char*) __func__, __LINE__)
#define host_realloc(ptr, size) _host_realloc((void**) (ptr), (size), (const char*) #ptr, (const char*) __func__, __LINE__)
#define host_free(ptr) _host_free((void**) (ptr), (const char*) #ptr, (const char*) __func__, __LINE__)

#ifdef __cplusplus
extern "C" {
#endif

LUMINARY_API LuminaryResult _host_malloc(void** ptr, size_t size, const char* buf_name, const char* func, uint32_t line);
LUMINARY_API LuminaryResult _host_realloc(void** ptr, size_t size, const char* buf_name, const char* func, uint32_t line);
LUMINARY_API LuminaryResult _host_free(void** ptr, const char* buf_name, const char* func, uint32_t line);

#ifdef __cplusplus
}
#endif

#endif /* LUMINARY_API_HOST_MEMORY_H */
