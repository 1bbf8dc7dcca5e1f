/*
 * luminary/log.h - console + in-memory log (reference log.h:22-83). LUMINARY_INCLUDE_EXTRA_UTILS pulls it into <luminary/luminary.h>
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/log.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LOG_H
#define LOG_H

#include <luminary/api_utils.h>

/* the reference's message macros: every message also lands in the log that luminary_write_log dumps to "luminary.log" */
#define log_message(fmt, ...) luminary_print_log("[%s:%d] " fmt, __func__, __LINE__, ##__VA_ARGS__)
#define info_message(fmt, ...)                                             \
  {                                                                        \
    luminary_print_info(false, fmt, ##__VA_ARGS__);                        \
    luminary_print_log("[%s:%d] " fmt, __func__, __LINE__, ##__VA_ARGS__); \
  }
#define warn_message(fmt, ...) luminary_print_warn("[%s:%d] " fmt, __func__, __LINE__, ##__VA_ARGS__)
#define error_message(fmt, ...) luminary_print_error("[%s:%d] " fmt, __func__, __LINE__, ##__VA_ARGS__)
#define crash_message(fmt, ...) luminary_print_crash("[%s:%d] " fmt, __func__, __LINE__, ##__VA_ARGS__)

#if __cplusplus
extern "C" {
#endif

LUMINARY_API void luminary_print_log(const char* format, ...);                   /* log only */
LUMINARY_API void luminary_print_info(bool log, const char* format, ...);        /* stdout (+ log) */
LUMINARY_API void luminary_print_info_inline(bool log, const char* format, ...); /* no newline; the next message overwrites it */
LUMINARY_API void luminary_print_warn(const char* format, ...);                  /* yellow */
LUMINARY_API void luminary_print_error(const char* format, ...);                 /* red */
LUMINARY_API void luminary_print_crash(const char* format, ...);                 /* purple; writes the log and terminates the program */
LUMINARY_API void luminary_write_log();

#if __cplusplus
}
#endif

#endif /* LOG_H */
