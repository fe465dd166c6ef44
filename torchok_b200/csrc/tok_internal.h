// tok_internal.h — helpers shared between the translation units of libtokb200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "tok_conv.cuh"

namespace tok {
int set_error(int code, const char* fmt, ...);
int make_tmap_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows);
int make_tmap_im2col(CUtensorMap* tm, const void* base, int n, int h, int w, int c, const PixelSrc& s, int pixels);
int make_tmap_im2col_ex(CUtensorMap* tm, const void* base, const cuuint64_t dims[4], const cuuint64_t strides[3],
                        const PixelSrc& s, int pixels);

#define TOK_CHECK_LAUNCH(name)                                                                  \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) return tok::set_error(-2, name ": %s", cudaGetErrorString(e__));    \
  } while (0)
}  // namespace tok
