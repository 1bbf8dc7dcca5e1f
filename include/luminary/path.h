/*
 * luminary/path.h - file system paths handed to the loaders and writers (reference path.h:23-28)
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/path.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LUMINARY_PATH_H
#define LUMINARY_PATH_H

#include <luminary/api_utils.h>
#include <luminary/error.h>

#ifdef __cplusplus
extern "C" {
#endif

struct LuminaryPath;
typedef struct LuminaryPath LuminaryPath;

LUMINARY_API LuminaryResult luminary_path_create(LuminaryPath** path);
LUMINARY_API LuminaryResult luminary_path_set_from_string(LuminaryPath* path, const char* string);
LUMINARY_API LuminaryResult luminary_path_destroy(LuminaryPath** path);

#ifdef __cplusplus
}
#endif

#endif /* LUMINARY_PATH_H */
