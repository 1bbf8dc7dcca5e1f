/*
 * luminary/name_strings.h - display names of the enumerators (reference name_strings.h:22-29)
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/name_strings.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LUMINARY_NAME_STRINGS_H
#define LUMINARY_NAME_STRINGS_H

#include "structs.h"

#ifdef __cplusplus
extern "C" {
#endif

extern const char* const luminary_strings_shading_mode[LUMINARY_SHADING_MODE_COUNT];
extern const char* const luminary_strings_adaptive_sampling_output_mode[LUMINARY_ADAPTIVE_SAMPLING_OUTPUT_MODE_COUNT];
extern const char* const luminary_strings_filter[LUMINARY_FILTER_COUNT];
extern const char* const luminary_strings_tonemap[LUMINARY_TONEMAP_COUNT];
extern const char* const luminary_strings_aperture[LUMINARY_APERTURE_COUNT];
extern const char* const luminary_strings_jerlov_water_type[LUMINARY_JERLOV_WATER_TYPE_COUNT];
extern const char* const luminary_strings_sky_mode[LUMINARY_SKY_MODE_COUNT];
extern const char* const luminary_strings_material_base_substrate[LUMINARY_MATERIAL_BASE_SUBSTRATE_COUNT];

#ifdef __cplusplus
}
#endif

#endif /* LUMINARY_NAME_STRINGS_H */
