/*
 * luminary/luminary.h - the public C API of Luminary as served by the B200-native path (luminary_b200).
 *
 * Applications written against MilchRatchet/Luminary include <luminary/luminary.h> and call luminary_init,
 * luminary_host_* and luminary_path_*; this header declares the same names, argument meanings, result codes and
 * struct layouts for the part of that API that drives the path-tracing hot path (reference: include/luminary/host.h:29-129,
 * structs.h, error.h:24-101, path.h:26-28, luminary.h:45-50), so such an application re-links against
 * libluminary_b200.so and renders through the CUDA kernels of liblumb200.so without source changes.
 *
 * The API is split over the same files as the reference's include/luminary (api_utils.h, error.h, path.h, structs.h, host.h,
 * name_strings.h and - behind LUMINARY_INCLUDE_EXTRA_UTILS - array.h, host_memory.h, log.h, queue.h, ringbuffer.h,
 * thread_status.h); this file is the umbrella.
 *
 * What is different, on purpose:
 *   - entities that are not on the path (ocean, clouds, fog, particles, sky HDRI) have their real layouts and entry points; they
 *     read back as inactive and answer LUMINARY_ERROR_NOT_IMPLEMENTED when asked to become active;
 *   - adaptive sampling is on by default as in the reference; its stages switch deterministically after update_interval << stage
 *     executions (the reference finishes the stage build asynchronously) and its executions run on the main device;
 *     undersampling (preview passes) is ignored with a warning;
 *   - the host renders until every requested output has been produced and then idles.
 */
#ifndef LUMINARY_H
#define LUMINARY_H

#include <luminary/api_utils.h>
#include <luminary/error.h>
#include <luminary/host.h>
#include <luminary/name_strings.h>
#include <luminary/path.h>
#include <luminary/structs.h>

/* utilities that do not follow the luminary_ naming scheme (the reference's front end uses them) */
#ifdef LUMINARY_INCLUDE_EXTRA_UTILS
#include <luminary/array.h>
#include <luminary/host_memory.h>
#include <luminary/log.h>
#include <luminary/queue.h>
#include <luminary/ringbuffer.h>
#include <luminary/thread_status.h>
#endif /* LUMINARY_INCLUDE_EXTRA_UTILS */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library life time (reference luminary.h:45-50): init once before anything else, shutdown after everything ---- */
LUMINARY_API void luminary_init(void);
LUMINARY_API void luminary_shutdown(void);

#ifdef __cplusplus
}
#endif

#endif /* LUMINARY_H */
