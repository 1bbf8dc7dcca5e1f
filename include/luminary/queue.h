/*
 * luminary/queue.h - bounded multi-producer / multi-consumer queue of fixed-size elements (reference queue.h:22-40)
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/queue.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LUMINARY_QUEUE_H
#define LUMINARY_QUEUE_H

#include <luminary/api_utils.h>
#include <luminary/error.h>

struct LuminaryQueue;
typedef struct LuminaryQueue LuminaryQueue;

typedef bool (*LuminaryEqOp)(void* lhs, void* rhs);

#define queue_create(queue, size_of_element, num_elements) \
  _queue_create((queue), (size_of_element), (num_elements), (const char*) #queue, (const char*) __func__, __LINE__)
#define queue_destroy(queue) _queue_destroy((queue), (const char*) #queue, (const char*) __func__, __LINE__)

#ifdef __cplusplus
extern "C" {
#endif

LUMINARY_API LuminaryResult
  _queue_create(LuminaryQueue** queue, size_t size_of_element, size_t num_elements, const char* buf_name, const char* func, uint32_t line);
LUMINARY_API LuminaryResult queue_push(LuminaryQueue* queue, void* object);
LUMINARY_API LuminaryResult queue_push_unique(LuminaryQueue* queue, void* object, LuminaryEqOp equal_operator, bool* already_queued);
LUMINARY_API LuminaryResult queue_pop(LuminaryQueue* queue, void* object, bool* success);
LUMINARY_API LuminaryResult queue_pop_blocking(LuminaryQueue* queue, void* object, bool* success);
LUMINARY_API LuminaryResult queue_set_is_blocking(LuminaryQueue* queue, bool is_blocking);
LUMINARY_API LuminaryResult _queue_destroy(LuminaryQueue** queue, const char* buf_name, const char* func, uint32_t line);

#ifdef __cplusplus
}
#endif

#endif /* LUMINARY_QUEUE_H */
