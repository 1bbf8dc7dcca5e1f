/*
 * luminary/ringbuffer.h - FIFO byte arena: entries are allocated at the head and released at the tail (reference ringbuffer.h:22-34)
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/ringbuffer.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LUMINARY_RINGBUFFER_H
#define LUMINARY_RINGBUFFER_H

#include <luminary/api_utils.h>
#include <luminary/error.h>

struct LuminaryRingBuffer;
typedef struct LuminaryRingBuffer LuminaryRingBuffer;

#define ringbuffer_create(buffer, size) _ringbuffer_create((buffer), (size), (const char*) #buffer, (const char*) __func__, __LINE__)
#define ringbuffer_destroy(buffer) _ringbuffer_destroy((buffer), (const char*) #buffer, (const char*) __func__, __LINE__)

#ifdef __cplusplus
extern "C" {
#endif

LUMINARY_API LuminaryResult
  _ringbuffer_create(LuminaryRingBuffer** buffer, size_t size, const char* buf_name, const char* func, uint32_t line);
LUMINARY_API LuminaryResult ringbuffer_allocate_entry(LuminaryRingBuffer* buffer, size_t entry_size, void** entry);
LUMINARY_API LuminaryResult ringbuffer_release_entry(LuminaryRingBuffer* buffer, size_t entry_size);
LUMINARY_API LuminaryResult _ringbuffer_destroy(LuminaryRingBuffer** buffer, const char* buf_name, const char* func, uint32_t line);

#ifdef __cplusplus
}
#endif

#endif /* LUMINARY_RINGBUFFER_H */
