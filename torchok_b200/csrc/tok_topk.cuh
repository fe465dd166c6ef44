// tok_topk.cuh — running top-KP list of one query row, folded over 32-score chunks of the tcgen05 accumulator.
// Shared by cosine_topk_kernel (tok_retrieval.cu) and its CTA-pair variant (tok_retrieval2.cu).
//
// ncu of the first version (profiles/r2_retrieval_kernel.md): tensor pipe 19 % active, 1 500 instructions per warp and
// 128 x 128 tile, issued by ONE warp per scheduler at 0.15 IPC (compare -> branch chains) = 10 000 clk per tile against
// 2 048 clk of UMMA time: the per-score filter was the bottleneck, not the gallery stream.  Now a chunk costs a 31-deep
// max tree (independent FMNMXs, full ILP) and ONE compare; the element-wise insertion only runs when the chunk's
// maximum beats the row's current KP-th best, which after the first few thousand gallery rows is a rare event.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace tok {

// strict '>' everywhere: of two equal scores the one seen first (the lower gallery index) stays ahead — faiss order
template <int KP>
__device__ __forceinline__ void topk_insert(float (&val)[KP], int (&id)[KP], float s, int col) {
#pragma unroll
  for (int u = KP - 1; u >= 1; --u) {
    const bool above = s > val[u - 1];
    const bool here = !above && s > val[u];
    val[u] = above ? val[u - 1] : (here ? s : val[u]);
    id[u] = above ? id[u - 1] : (here ? col : id[u]);
  }
  if (s > val[0]) {
    val[0] = s;
    id[0] = col;
  }
}

// r: 32 consecutive accumulator columns (inner products) of this thread's query row, gallery rows col0 .. col0+31.
// g_sqnorm != nullptr: L2 metric, ranked by 2*q.g - |g|^2.  Columns >= ng (ragged last tile) never enter the list.
template <int KP>
__device__ __forceinline__ void topk_fold_chunk(float (&val)[KP], int (&id)[KP], const uint32_t (&r)[32], int col0,
                                                int ng, const float* __restrict__ g_sqnorm) {
  float s[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j] = __uint_as_float(r[j]);
  const bool full = col0 + 32 <= ng;   // warp-uniform
  if (g_sqnorm != nullptr) {
    if (full) {
      const float4* np = reinterpret_cast<const float4*>(g_sqnorm + col0);   // col0 % 32 == 0, cudaMalloc-aligned base
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 n4 = __ldg(np + j);
        s[4 * j] = 2.f * s[4 * j] - n4.x;
        s[4 * j + 1] = 2.f * s[4 * j + 1] - n4.y;
        s[4 * j + 2] = 2.f * s[4 * j + 2] - n4.z;
        s[4 * j + 3] = 2.f * s[4 * j + 3] - n4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) s[j] = 2.f * s[j] - (col0 + j < ng ? __ldg(g_sqnorm + col0 + j) : 0.f);
    }
  }
  if (!full) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (col0 + j >= ng) s[j] = -INFINITY;
  }
  float m[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) m[j] = fmaxf(s[2 * j], s[2 * j + 1]);
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1) {
#pragma unroll
    for (int j = 0; j < w; ++j) m[j] = fmaxf(m[j], m[j + w]);
  }
  if (m[0] > val[KP - 1]) {   // rare once the list has warmed up
#pragma unroll
    for (int j = 0; j < 32; ++j)   // fully unrolled: s[] must stay in registers
      if (s[j] > val[KP - 1]) topk_insert<KP>(val, id, s[j], col0 + j);
  }
}

}  // namespace tok
