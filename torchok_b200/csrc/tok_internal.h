// tok_internal.h — helpers shared between the translation units of libtokb200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "tok_conv.cuh"

namespace tok {
int set_error(int code, const char* fmt, ...);
int make_tmap_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows);
int make_tmap_im2col(CUtensorMap* tm, const void* base, int n, int h, int w, int c, const PixelSrc& s, int pixels);
int make_tmap_im2col_ex(CUtensorMap* tm, const void* base, const cuuint64_t dims[4], const cuuint64_t strides[3],
                        const PixelSrc& s, int pixels);

// Launch with programmatic stream serialization (PDL): the kernel's CTAs may be scheduled — and run their local prologue
// (barrier init, TMEM allocation, descriptor prefetch) — while the previous kernel of the stream drains; the kernel
// calls pdl_wait() before it touches anything its predecessors produced (tok_ptx.cuh).  Works under stream capture
// (programmatic graph edges).  TOK_PDL=0 falls back to plain launches (A/B aid).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define TOK_CHECK_LAUNCH(name)                                                                  \
  do {                                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ != cudaSuccess) return tok::set_error(-2, name ": %s", cudaGetErrorString(e__));    \
  } while (0)
}  // namespace tok
