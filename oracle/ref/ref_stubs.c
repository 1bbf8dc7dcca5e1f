/* ref_stubs.c - link-time stand-ins for the reference's device-manager functions that its host-side C files reference
 * but that the host-only oracle never reaches (textured-emitter integration, OptiX light BVH upload).
 * TEST INFRASTRUCTURE ONLY (see oracle/lum_oracle.h). Any call aborts loudly. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define REF_STUB(name)                                                                 \
  uint64_t name(void) {                                                                \
    fprintf(stderr, "oracle/_ref: device function %s is not available here\n", #name); \
    abort();                                                                           \
    return 0;                                                                          \
  }

REF_STUB(_device_free)
REF_STUB(_device_malloc)
REF_STUB(device_download)
REF_STUB(device_upload)
REF_STUB(device_staging_manager_execute)
REF_STUB(device_staging_manager_register)
REF_STUB(kernel_execute_custom)
REF_STUB(kernel_execute_with_args)
REF_STUB(device_download2D)
REF_STUB(device_sync_constant_memory)
REF_STUB(device_texture_create)
REF_STUB(device_texture_destroy)
REF_STUB(texture_create)
REF_STUB(texture_destroy)
REF_STUB(texture_fill)
REF_STUB(optix_bvh_create)
REF_STUB(optix_bvh_destroy)
REF_STUB(optix_bvh_light_build)

/* optix_stubs.h expects the function table symbol; never dereferenced by the host-only path. */
char g_optixFunctionTable_105[4096];
