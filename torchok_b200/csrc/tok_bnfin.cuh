// tok_bnfin.cuh — device side of ApplyFin (tok_conv.cuh): BatchNorm finalize inside the apply kernels.
#pragma once
#include <cuda_runtime.h>

#include "tok_conv.cuh"

namespace tok {

// scale / shift of channels c0 .. c0+7 from the completed batch sums (same arithmetic as bn_finalize_train_kernel)
__device__ __forceinline__ void applyfin_coefs(const ApplyFin& f, int c0, float (&sc)[8], float (&sf)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const float mean = __ldcg(f.sum + c) / f.count;
    float var = __ldcg(f.sqsum + c) / f.count - mean * mean;
    var = fmaxf(var, 0.f);
    const float invstd = rsqrtf(var + f.eps);
    const bool real = c < f.Cv;
    const float g = real ? (f.gamma ? f.gamma[c] : 1.f) : 0.f;
    const float b = real ? (f.beta ? f.beta[c] : 0.f) : 0.f;
    sc[j] = g * invstd;
    sf[j] = b - mean * g * invstd;
  }
}

// Called by EVERY thread of EVERY CTA after its own reads of the sums (applyfin_coefs): CTA `first` publishes the
// per-channel results and the running statistics, the last CTA through the ticket zeroes the sums.  s_flag: one int of
// shared memory.
__device__ __forceinline__ void applyfin_publish(const ApplyFin& f, bool first, unsigned total_ctas, int* s_flag) {
  if (first) {
    for (int c = threadIdx.x; c < f.C; c += blockDim.x) {
      const float mean = __ldcg(f.sum + c) / f.count;
      float var = __ldcg(f.sqsum + c) / f.count - mean * mean;
      var = fmaxf(var, 0.f);
      const float invstd = rsqrtf(var + f.eps);
      const bool real = c < f.Cv;
      const float g = real ? (f.gamma ? f.gamma[c] : 1.f) : 0.f;
      const float b = real ? (f.beta ? f.beta[c] : 0.f) : 0.f;
      f.scale[c] = g * invstd;
      f.shift[c] = b - mean * g * invstd;
      f.save_mean[c] = mean;
      f.save_invstd[c] = invstd;
      if (f.running_mean && real) {
        const float unbiased = f.count > 1.f ? var * f.count / (f.count - 1.f) : var;
        f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
        f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * unbiased;
      }
    }
    // chain mode: the accumulators the previous fused apply of this stream left behind (that kernel has finished)
    for (int i = threadIdx.x; i < f.zero_n; i += blockDim.x) f.zero_ptr[i] = 0.f;
  }
  if (f.counter == nullptr) return;   // chain mode: no ticket, the sums are zeroed by the next fused apply
  __syncthreads();   // every thread of this CTA has read the sums
  if (threadIdx.x == 0) {
    __threadfence();
    *s_flag = atomicAdd(f.counter, 1u) == total_ctas - 1 ? 1 : 0;
  }
  __syncthreads();
  if (*s_flag) {
    __threadfence();
    for (int c = threadIdx.x; c < f.C; c += blockDim.x) {
      f.sum[c] = 0.f;
      f.sqsum[c] = 0.f;
    }
    if (threadIdx.x == 0) *f.counter = 0u;
  }
}

}  // namespace tok
