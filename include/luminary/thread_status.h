/*
 * luminary/thread_status.h - named worker + current task + wall time since the task started (reference thread_status.h:22-34)
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/thread_status.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LUMINARY_WALL_TIME_H
#define LUMINARY_WALL_TIME_H

#include <luminary/api_utils.h>
#include <luminary/error.h>

struct LuminaryThreadStatus;
typedef struct LuminaryThreadStatus LuminaryThreadStatus;

#ifdef __cplusplus
extern "C" {
#endif

LUMINARY_API LuminaryResult thread_status_create(LuminaryThreadStatus** thread_status);
LUMINARY_API LuminaryResult thread_status_set_worker_name(LuminaryThreadStatus* thread_status, const char* name);
LUMINARY_API LuminaryResult thread_status_get_worker_name(LuminaryThreadStatus* thread_status, const char** name);
LUMINARY_API LuminaryResult thread_status_start(LuminaryThreadStatus* thread_status, const char* string);
LUMINARY_API LuminaryResult thread_status_get_time(LuminaryThreadStatus* thread_status, double* time);
LUMINARY_API LuminaryResult thread_status_get_string(LuminaryThreadStatus* thread_status, const char** string);
LUMINARY_API LuminaryResult thread_status_stop(LuminaryThreadStatus* thread_status);
LUMINARY_API LuminaryResult thread_status_destroy(LuminaryThreadStatus** thread_status);

#ifdef __cplusplus
}
#endif

#endif /* LUMINARY_WALL_TIME_H */
