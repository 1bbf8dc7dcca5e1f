/*
 * luminary/array.h - growable typed arrays with a hidden header in front of the data (reference array.h:22-46)
 *
 * Part of the public C API of MilchRatchet/Luminary as served by the B200-native path (libluminary_b200.so): same file name, same
 * names, argument meanings, result codes and struct layouts as the reference's include/luminary/array.h, so that an application
 * written against Luminary compiles against this directory unchanged (tests/test_reference_frontend.py builds the reference's own
 * command line front end against it). Restated, not copied: see INTEGRATION.md.
 */
#ifndef LUMINARY_API_ARRAY_H
#define LUMINARY_API_ARRAY_H

#include <luminary/api_utils.h>
#include <luminary/error.h>

/* `array` is the ADDRESS of a typed pointer (T**); the macros record the variable name and call site for leak reports */
#define array_create(array, size_of_element, num_elements) \
  _array_create((void**) (array), (size_of_element), (num_elements), (const char*) #array, (const char*) __func__, __LINE__)
#define array_resize(array, size) _array_resize((void**) (array), (size), (const char*) #array, (const char*) __func__, __LINE__)
#define array_push(array, object) _array_push((void**) (array), (void*) (object), (const char*) #array, (const char*) __func__, __LINE__)
#define array_copy(dst, src) _array_copy((void**) (dst), (void**) (src), (const char*) #dst, (const char*) __func__, __LINE__)
#define array_append(dst, src) _array_append((void**) (dst), (const void*) (src), (const char*) #dst, (const char*) __func__, __LINE__)
#define array_set_num_elements(array, num_elements) \
  _array_set_num_elements((void**) (array), (num_elements), (const char*) #array, (const char*) __func__, __LINE__)
#define array_destroy(array) _array_destroy((void**) (array), (const char*) #array, (const char*) __func__, __LINE__)

#ifdef __cplusplus
extern "C" {
#endif

LUMINARY_API LuminaryResult
  _array_create(void** array, size_t size_of_element, uint32_t num_elements, const char* buf_name, const char* func, uint32_t line);
LUMINARY_API LuminaryResult _array_resize(void** array, size_t size, const char* buf_name, const char* func, uint32_t line);
LUMINARY_API LuminaryResult _array_push(void** array, void* object, const char* buf_name, const char* func, uint32_t line);
LUMINARY_API LuminaryResult _array_copy(void** dst, const void* src, const char* buf_name, const char* func, uint32_t line);
LUMINARY_API LuminaryResult _array_append(void** dst, const void* src, const char* buf_name, const char* func, uint32_t line);
LUMINARY_API LuminaryResult _array_destroy(void** array, const char* buf_name, const char* func, uint32_t line);

LUMINARY_API LuminaryResult array_clear(void* array);
LUMINARY_API LuminaryResult array_get_size(const void* array, size_t* size);
LUMINARY_API LuminaryResult array_get_num_elements(const void* array, uint32_t* num_elements);
LUMINARY_API LuminaryResult
  _array_set_num_elements(void** array, uint32_t num_elements, const char* buf_name, const char* func, uint32_t line);

#ifdef __cplusplus
}
#endif

#endif /* LUMINARY_API_ARRAY_H */
