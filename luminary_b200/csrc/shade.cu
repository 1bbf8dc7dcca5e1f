// shade.cu - TEMPORARY STUB (replaced by the real shading kernels)
#include "shade_api.cuh"
void lb_launch_shade(const LbShadeParams&, int, cudaStream_t) {}
void lb_launch_accumulate(const LbPaths&, const LbFrame&, float*, int, cudaStream_t) {}
void lb_launch_generate_result(const float*, float*, uint32_t, uint32_t, int, cudaStream_t) {}
Lumb200Result lb_lut_generate(LbLutTextures*, const uint32_t*, cudaStream_t) { return LUMB200_ERROR_NOT_IMPLEMENTED; }
Lumb200Result lb_lut_upload(LbLutTextures*, const uint16_t*, const uint16_t*, const uint16_t*, const uint16_t*, cudaStream_t) { return LUMB200_ERROR_NOT_IMPLEMENTED; }
Lumb200Result lb_lut_download(LbLutTextures*, uint16_t*, uint16_t*, uint16_t*, uint16_t*, cudaStream_t) { return LUMB200_ERROR_NOT_IMPLEMENTED; }
void lb_lut_destroy(LbLutTextures*) {}
