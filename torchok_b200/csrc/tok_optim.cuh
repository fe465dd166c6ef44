// tok_optim.cuh — per-element optimizer arithmetic shared by the arena step kernels (tok_elem.cu) and the fused
// peer-memory gradient-exchange + optimizer kernel (tok_comm.cu).  torch.optim.SGD / Adam / AdamW semantics
// (registered by torchok/optim/optimizers/__init__.py:9-19).
#pragma once
#include <cuda_runtime.h>

namespace tok {

struct SgdArgs {
  float lr, mu, wd, damp, gscale;
  int nesterov, first, zero_grad;
};
__device__ __forceinline__ float sgd_one(float w, float g, float* buf, const SgdArgs& a) {
  float d = g * a.gscale + a.wd * w;
  if (a.mu != 0.f) {
    const float b = a.first ? d : a.mu * (*buf) + (1.f - a.damp) * d;
    *buf = b;
    d = a.nesterov ? d + a.mu * b : b;
  }
  return w - a.lr * d;
}
// paramwise_cfg (torchok/constructor/constructor.py:162-251): per-parameter lr / weight-decay multipliers, looked up by
// the element offset in a sorted segment table (one segment per parameter of the arena); n == 0 means "all ones".
struct ParamSegs {
  const int* begin;
  const float* lr_mult;
  const float* wd_mult;
  int n;
  // optional (Adam): per-parameter step counts.  torch.optim.Adam keeps state['step'] per parameter and skips parameters
  // without a gradient, so a parameter thawed by FreezeUnfreeze starts its bias correction at 1; the counters advance
  // only while the parameter's lr multiplier is non-zero (seg_steps_advance_kernel).  nullptr: the global step is used.
  const int* steps;
};
__device__ __forceinline__ int seg_lookup(const ParamSegs& s, long long elem, float& lr_mult, float& wd_mult) {
  lr_mult = wd_mult = 1.f;
  if (s.n == 0) return 0;
  int lo = 0, hi = s.n - 1;   // largest k with begin[k] <= elem
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(s.begin + mid) <= elem) lo = mid;
    else hi = mid - 1;
  }
  lr_mult = __ldg(s.lr_mult + lo);
  wd_mult = __ldg(s.wd_mult + lo);
  return lo;
}
struct AdamArgs {
  float lr, b1, b2, eps, wd, gscale, step, rbc2;
  int decoupled;
};
__device__ __forceinline__ float adam_one(float w, float g, float* m, float* v, const AdamArgs& a) {
  float d = g * a.gscale;
  if (a.decoupled)
    w *= 1.f - a.lr * a.wd;
  else
    d += a.wd * w;
  const float mi = a.b1 * (*m) + (1.f - a.b1) * d;
  const float vi = a.b2 * (*v) + (1.f - a.b2) * d * d;
  *m = mi;
  *v = vi;
  return w - a.step * mi / (sqrtf(vi) * a.rbc2 + a.eps);
}

}  // namespace tok
