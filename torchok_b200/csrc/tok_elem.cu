// tok_elem.cu — HBM-bound passes around the tensor-core kernels: BatchNorm finalize/apply/backward with fused
// ReLU and residual add, pooling, softmax cross-entropy, layout packing, optimizer steps.
// Layout everywhere: NHWC bf16, i.e. a [rows = N*H*W][C] matrix with C contiguous; 8 channels (16 bytes) per access.
//
// Reference call sites: torch.nn.BatchNorm2d + ReLU in ConvBnAct (torchok/models/modules/bricks/convbnact.py:44-53),
// timm block tails `x += shortcut; act(x)` built by torchok/models/backbones/resnet.py:363-405, the stem maxpool
// (resnet.py:510), SelectAdaptivePool2d (torchok/models/poolings/classification/pooling.py:8-12),
// torch.nn.CrossEntropyLoss (torchok/losses/__init__.py:26).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/tokb200.h"
#include "tok_internal.h"
#include "tok_ptx.cuh"
#include "tok_optim.cuh"
#include "tok_bnfin.cuh"

namespace tok {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x);
  f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z);
  f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

static inline int elem_grid(long long work_items, int block) {
  long long b = (work_items + block - 1) / block;
  const long long cap = 148LL * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------------------------------------ dgrad helper
// dst[n, p*sh, q*sw, :] = src[n, p, q, :]  (dst pre-zeroed): zero-dilation of an output gradient.
__global__ void dilate_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long total, int P,
                                   int Q, int cvec, int H, int W, int sh, int sw) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int q = (int)(pix % Q);
    pix /= Q;
    const int p = (int)(pix % P);
    const long long n = pix / P;
    dst[((n * H + (long long)p * sh) * W + (long long)q * sw) * cvec + cv] = src[i];
  }
}
void launch_dilate_rows(const __nv_bfloat16* src, __nv_bfloat16* dst, int n, int p, int q, int c, int H, int W,
                        int sh, int sw, cudaStream_t st) {
  const long long total = (long long)n * p * q * (c / 8);
  (void)launch_pdl(dilate_rows_kernel, dim3(elem_grid(total, 256)), dim3(256), 0, st, reinterpret_cast<const uint4*>(src),
                                                            reinterpret_cast<uint4*>(dst), total, p, q, c / 8, H, W,
                                                            sh, sw);
}

// ------------------------------------------------------------------------------------------------ BatchNorm forward
// Training-mode statistics -> per-channel affine (scale, shift); running-stat update as torch.nn.BatchNorm2d:
// biased variance normalises, unbiased variance feeds running_var.
__global__ void bn_finalize_train_kernel(float* __restrict__ sum, float* __restrict__ sqsum, float count,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                         float momentum, float* running_mean, float* running_var,
                                         float* __restrict__ scale, float* __restrict__ shift,
                                         float* __restrict__ save_mean, float* __restrict__ save_invstd, int C,
                                         int Cv) {
  pdl_wait();
  pdl_launch();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = sum[c] / count;
  float var = sqsum[c] / count - mean * mean;
  var = fmaxf(var, 0.f);
  sum[c] = 0.f;  // consumed: the accumulators are handed back zeroed for the next step (no memset launches)
  sqsum[c] = 0.f;
  const float invstd = rsqrtf(var + eps);
  // c >= Cv: pad lane of a channel count that is not a multiple of 8 (gamma / beta / running buffers hold Cv entries);
  // gamma = beta = 0 keeps the lane exactly zero
  const bool real = c < Cv;
  const float g = real ? (gamma ? gamma[c] : 1.f) : 0.f;
  const float b = real ? (beta ? beta[c] : 0.f) : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - mean * g * invstd;
  save_mean[c] = mean;
  save_invstd[c] = invstd;
  if (running_mean && real) {
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}
__global__ void bn_finalize_eval_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                        float* __restrict__ scale, float* __restrict__ shift, int C, int Cv) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c >= Cv) {   // pad lane
    scale[c] = 0.f;
    shift[c] = 0.f;
    return;
  }
  const float invstd = rsqrtf(running_var[c] + eps);
  const float g = gamma ? gamma[c] : 1.f;
  const float b = beta ? beta[c] : 0.f;
  scale[c] = g * invstd;
  shift[c] = b - running_mean[c] * g * invstd;
}

// out = act(y * scale[c] + shift[c] (+ residual)).  One 16-byte vector (8 channels) per thread per iteration.
template <bool HAS_RES, bool RELU>
__global__ void __launch_bounds__(256) bn_apply_kernel(const uint4* __restrict__ y, const uint4* __restrict__ res,
                                                       uint4* __restrict__ out, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, long long total, int cvec,
                                                       const ApplyFin fin) {
  __shared__ int s_flag;
  pdl_wait();
  pdl_launch();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool invariant = (stride % cvec) == 0;   // the host guarantees it when fin.counter != nullptr
  float sc[8], sf[8];
  int cv = (int)(i % cvec);
  if (fin.fused) {
    applyfin_coefs(fin, cv * 8, sc, sf);
    applyfin_publish(fin, blockIdx.x == 0, gridDim.x, &s_flag);
  } else if (invariant) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(scale + cv * 8 + j);
      sf[j] = __ldg(shift + cv * 8 + j);
    }
  }
  if (invariant) {
    // four independent 16-byte loads in flight per thread (one load -> compute -> store chain per iteration left the
    // kernel latency-bound at 0.67-0.77 of the HBM rate; the unrolled bn2 passes reach 0.87-0.94)
    constexpr int U = 4;
    for (; i < total; i += stride * U) {
      uint4 vy[U], vr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long k = i + u * stride;
        if (k < total) {
          vy[u] = ldg_stream(y + k);
          if (HAS_RES) vr[u] = ldg_stream(res + k);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long k = i + u * stride;
        if (k < total) {
          float f[8];
          unpack8(vy[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sf[j]);
          if (HAS_RES) {
            float r[8];
            unpack8(vr[u], r);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += r[j];
          }
          if (RELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          out[k] = pack8(f);
        }
      }
    }
    return;
  }
  for (; i < total; i += stride) {
    cv = (int)(i % cvec);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(scale + cv * 8 + j);
      sf[j] = __ldg(shift + cv * 8 + j);
    }
    float f[8];
    unpack8(ldg_stream(y + i), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sf[j]);
    if (HAS_RES) {
      float r[8];
      unpack8(ldg_stream(res + i), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    if (RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    out[i] = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------ BatchNorm backward
// Pass 1: g = (dout (+dout2)) * [out > 0];  sum_g[c] += sum g ; sum_gy[c] += sum g*y   (fp32 atomics).
// A 256-thread block is (rows_per_block x cvec_b) with the channel vector fixed per thread.
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ dout2, const uint4* __restrict__ out,
                     const uint4* __restrict__ y, float* __restrict__ sum_g, float* __restrict__ sum_gy,
                     long long rows, int cvec, int cvec_b, int rows_per_cta) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ float red[];  // [cvec_b*8][2]
  const int rlanes = 256 / cvec_b;
  const int cl = threadIdx.x % cvec_b;
  const int rl = threadIdx.x / cvec_b;
  const int cv = blockIdx.y * cvec_b + cl;
  for (int i = threadIdx.x; i < cvec_b * 16; i += 256) red[i] = 0.f;
  __syncthreads();
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  if (rl < rlanes && cv < cvec) {
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    long long r1 = r0 + rows_per_cta;
    if (r1 > rows) r1 = rows;
    for (long long r = r0 + rl; r < r1; r += rlanes) {
      const long long idx = r * cvec + cv;
      float g[8], yy[8];
      unpack8(ldg_stream(dout + idx), g);
      if (dout2) {
        float g2[8];
        unpack8(ldg_stream(dout2 + idx), g2);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] += g2[j];
      }
      if (out) {
        float o[8];
        unpack8(ldg_stream(out + idx), o);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
      }
      unpack8(ldg_stream(y + idx), yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] += g[j];
        a2[j] = fmaf(g[j], yy[j], a2[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[(cl * 8 + j) * 2], a1[j]);
      atomicAdd(&red[(cl * 8 + j) * 2 + 1], a2[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cvec_b * 8; i += 256) {
    const int c = blockIdx.y * cvec_b * 8 + i;
    if (c < cvec * 8) {
      atomicAdd(sum_g + c, red[i * 2]);
      atomicAdd(sum_gy + c, red[i * 2 + 1]);
    }
  }
}

// Coefficients of the data gradient  dy = a*g + c1*y + c0  and the parameter gradients.
//   xhat = (y-mean)*invstd ; sum_gx = invstd*(sum_gy - mean*sum_g)
//   dy = gamma*invstd*(g - sum_g/M - xhat*sum_gx/M)
__global__ void bn_bwd_finalize_kernel(float* __restrict__ sum_g, float* __restrict__ sum_gy,
                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                       const float* __restrict__ gamma, float count, float* __restrict__ coef_a,
                                       float* __restrict__ coef_c1, float* __restrict__ coef_c0, float* dgamma,
                                       float* dbeta, int accumulate, int C, int Cv) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sg = sum_g[c];
  const float mu = mean[c];
  const float is = invstd[c];
  const float sgx = is * (sum_gy[c] - mu * sg);
  sum_g[c] = 0.f;  // consumed (see bn_finalize_train_kernel)
  sum_gy[c] = 0.f;
  const bool real = c < Cv;
  const float g = real ? (gamma ? gamma[c] : 1.f) : 0.f;
  const float a = g * is;
  const float k1 = sg / count;
  const float k2 = sgx / count;
  coef_a[c] = a;
  coef_c1[c] = -a * k2 * is;
  coef_c0[c] = -a * k1 + a * k2 * is * mu;
  if (dgamma && real) dgamma[c] = accumulate ? dgamma[c] + sgx : sgx;
  if (dbeta && real) dbeta[c] = accumulate ? dbeta[c] + sg : sg;
}

// Pass 2: dy = a*g + c1*y + c0 ; optionally also stores g (the gradient that flows to the residual branch).
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ dout2, const uint4* __restrict__ out,
                    const uint4* __restrict__ y, const float* __restrict__ coef_a, const float* __restrict__ coef_c1,
                    const float* __restrict__ coef_c0, uint4* __restrict__ dy, uint4* __restrict__ dres,
                    long long total, int cvec) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool invariant = (stride % cvec) == 0;
  float ca[8], c1[8], c0[8];
  int cv = (int)(i % cvec);
  if (invariant) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ca[j] = __ldg(coef_a + cv * 8 + j);
      c1[j] = __ldg(coef_c1 + cv * 8 + j);
      c0[j] = __ldg(coef_c0 + cv * 8 + j);
    }
  }
  for (; i < total; i += stride) {
    if (!invariant) {
      cv = (int)(i % cvec);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ca[j] = __ldg(coef_a + cv * 8 + j);
        c1[j] = __ldg(coef_c1 + cv * 8 + j);
        c0[j] = __ldg(coef_c0 + cv * 8 + j);
      }
    }
    float g[8], yy[8];
    unpack8(ldg_stream(dout + i), g);
    if (dout2) {
      float g2[8];
      unpack8(ldg_stream(dout2 + i), g2);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += g2[j];
    }
    if (out) {
      float o[8];
      unpack8(ldg_stream(out + i), o);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
    }
    unpack8(ldg_stream(y + i), yy);
    if (dres) dres[i] = pack8(g);
#pragma unroll
    for (int j = 0; j < 8; ++j) yy[j] = fmaf(ca[j], g[j], fmaf(c1[j], yy[j], c0[j]));
    dy[i] = pack8(yy);
  }
}

// ------------------------------------------------------------------------------------------------ pooling
// Max pool, NHWC, window k x k, stride s, padding pd (implicit -inf). First maximum in (r, s) scan order wins, as in
// ATen's max_pool2d; the winner's window slot (r*k + s) is stored for the backward pass.
__global__ void maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, uint2* __restrict__ arg,
                                   long long total, int H, int W, int P, int Q, int cvec, int k, int s, int pd) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long t = i / cvec;
    const int q = (int)(t % Q);
    t /= Q;
    const int p = (int)(t % P);
    const long long n = t / P;
    float best[8];
    uint32_t bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -INFINITY;
      bi[j] = 0;
    }
    for (int r = 0; r < k; ++r) {
      const int h = p * s - pd + r;
      if (h < 0 || h >= H) continue;
      for (int c = 0; c < k; ++c) {
        const int w = q * s - pd + c;
        if (w < 0 || w >= W) continue;
        float f[8];
        unpack8(__ldg(x + ((n * H + h) * W + w) * cvec + cv), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (f[j] > best[j] || f[j] != f[j]) {
            best[j] = f[j];
            bi[j] = r * k + c;
          }
        }
      }
    }
    out[i] = pack8(best);
    arg[i] = make_uint2(bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24),
                        bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24));
  }
}

// dx[n,h,w,c] = sum over windows (p,q) that contain (h,w) of dout[n,p,q,c] * [arg[n,p,q,c] == slot of (h,w)].
__global__ void maxpool_bwd_kernel(const uint4* __restrict__ dout, const uint2* __restrict__ arg,
                                   uint4* __restrict__ dx, long long total, int H, int W, int P, int Q, int cvec,
                                   int k, int s, int pd) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long t = i / cvec;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const long long n = t / H;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // windows p with p*s - pd <= h <= p*s - pd + k - 1
    int p_lo = (h + pd - (k - 1) + s - 1) / s;
    if (h + pd - (k - 1) < 0) p_lo = 0;
    int p_hi = (h + pd) / s;
    if (p_hi > P - 1) p_hi = P - 1;
    int q_lo = (w + pd - (k - 1) + s - 1) / s;
    if (w + pd - (k - 1) < 0) q_lo = 0;
    int q_hi = (w + pd) / s;
    if (q_hi > Q - 1) q_hi = Q - 1;
    for (int p = p_lo; p <= p_hi; ++p) {
      const int r = h - (p * s - pd);
      for (int q = q_lo; q <= q_hi; ++q) {
        const int c = w - (q * s - pd);
        const uint32_t slot = r * k + c;
        const long long o = ((n * P + p) * Q + q) * cvec + cv;
        const uint2 a = __ldg(arg + o);
        float g[8];
        unpack8(__ldg(dout + o), g);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (((a.x >> (8 * j)) & 0xFF) == slot) acc[j] += g[j];
          if (((a.y >> (8 * j)) & 0xFF) == slot) acc[4 + j] += g[4 + j];
        }
      }
    }
    dx[i] = pack8(acc);
  }
}

// Global average (and/or max) pool over HW: x [N][HW][C] -> out [N][C].
// mode 0 avg, 1 max, 2 avgmax = 0.5*(avg+max).  (timm SelectAdaptivePool2d, pooling.py:8-12)
__global__ void gap_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int N, int HW, int cvec,
                               int mode) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)N * cvec) return;
  const int cv = (int)(i % cvec);
  const long long n = i / cvec;
  float s[8], m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j] = 0.f;
    m[j] = -INFINITY;
  }
  for (int r = 0; r < HW; ++r) {
    float f[8];
    unpack8(__ldg(x + (n * HW + r) * cvec + cv), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += f[j];
      m[j] = fmaxf(m[j], f[j]);
    }
  }
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float avg = s[j] / HW;
    o[j] = mode == 0 ? avg : (mode == 1 ? m[j] : 0.5f * (avg + m[j]));
  }
  out[i] = pack8(o);
}
// Backward of the average pool: dx[n, r, c] = dout[n, c] / HW.
__global__ void gap_bwd_kernel(const uint4* __restrict__ dout, uint4* __restrict__ dx, long long total, int HW,
                               int cvec) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const float inv = 1.f / HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const long long n = i / ((long long)cvec * HW);
    float f[8];
    unpack8(__ldg(dout + n * cvec + cv), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= inv;
    dx[i] = pack8(f);
  }
}

// Backward of the max / avgmax pool (F.adaptive_max_pool2d semantics: the FIRST maximal position in scan order takes
// the whole max-path gradient).  One thread per (n, 8-channel vector): pass 1 finds the arg-max rows, pass 2 writes dx.
__global__ void gap_bwd_max_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ x, uint4* __restrict__ dx,
                                   int N, int HW, int cvec, int mode) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)N * cvec) return;
  const int cv = (int)(i % cvec);
  const long long n = i / cvec;
  float m[8];
  int am[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m[j] = -INFINITY;
    am[j] = 0;
  }
  for (int r = 0; r < HW; ++r) {
    float f[8];
    unpack8(__ldg(x + (n * HW + r) * cvec + cv), f);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (f[j] > m[j] || (r == 0)) {   // NaN-free inputs; strict '>' keeps the first maximum
        m[j] = f[j];
        am[j] = r;
      }
  }
  float g[8];
  unpack8(__ldg(dout + n * cvec + cv), g);
  const float wmax = mode == 1 ? 1.f : 0.5f;
  const float wavg = mode == 1 ? 0.f : 0.5f / HW;
  for (int r = 0; r < HW; ++r) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = g[j] * (wavg + (am[j] == r ? wmax : 0.f));
    dx[(n * HW + r) * cvec + cv] = pack8(o);
  }
}

// ------------------------------------------------------------------------------------------------ cross entropy
// One block per row. loss_sum += -log softmax(logits[row])[target] / norm ; dlogits = (softmax - onehot) * gscale.
// Rows whose target == ignore_index contribute nothing (torch.nn.CrossEntropyLoss semantics).
__global__ void __launch_bounds__(256)
softmax_xent_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ target,
                    float* __restrict__ loss_sum, __nv_bfloat16* __restrict__ dlogits, int C, long long ld,
                    float inv_norm, float gscale, const float* __restrict__ gscale_dev, long long ignore_index,
                    int* __restrict__ correct) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float sred[32];
  if (gscale_dev) gscale *= __ldg(gscale_dev);
  __shared__ int sidx[32];
  const long long row = blockIdx.x;
  const __nv_bfloat16* lp = logits + row * ld;
  const long long tgt = target[row];
  float mx = -INFINITY;
  int amax = 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = __bfloat162float(lp[c]);
    if (v > mx) {
      mx = v;
      amax = c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, amax, o);
    if (ov > mx || (ov == mx && oi < amax)) {
      mx = ov;
      amax = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    sred[threadIdx.x >> 5] = mx;
    sidx[threadIdx.x >> 5] = amax;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sred[threadIdx.x] : -INFINITY;
    int vi = threadIdx.x < (blockDim.x >> 5) ? sidx[threadIdx.x] : 0x7fffffff;
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
      if (ov > v || (ov == v && oi < vi)) {
        v = ov;
        vi = oi;
      }
    }
    if (threadIdx.x == 0) {
      sred[0] = v;
      sidx[0] = vi;
    }
  }
  __syncthreads();
  mx = sred[0];
  amax = sidx[0];
  __syncthreads();
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += __expf(__bfloat162float(lp[c]) - mx);
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = se;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? sred[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) sred[0] = v;
  }
  __syncthreads();
  se = sred[0];
  const bool ignored = (tgt == ignore_index);
  if (threadIdx.x == 0 && !ignored) {
    const float lt = __bfloat162float(lp[tgt]);
    if (loss_sum) atomicAdd(loss_sum, (logf(se) + mx - lt) * inv_norm);
    if (correct && amax == (int)tgt) atomicAdd(correct, 1);
  }
  if (dlogits) {
    __nv_bfloat16* dp = dlogits + row * ld;
    const float inv = 1.f / se;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float g = 0.f;
      if (!ignored) {
        g = __expf(__bfloat162float(lp[c]) - mx) * inv;
        if (c == tgt) g -= 1.f;
        g *= gscale;
      }
      dp[c] = __float2bfloat16(g);
    }
  }
}

// ------------------------------------------------------------------------------------------------ layout packing
// NCHW (fp32 or bf16) -> NHWC bf16 with the channel count padded to Cp (zeros), via a 32x32 shared-memory transpose.
template <typename T>
__global__ void nchw_to_nhwc_kernel(const T* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C, int HW,
                                    int Cp) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, hw = hw0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && hw < HW) ? (float)src[(n * C + c) * HW + hw] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int hw = hw0 + j, c = c0 + threadIdx.x;
    if (hw < HW && c < Cp) dst[(n * HW + hw) * Cp + c] = __float2bfloat16(tile[threadIdx.x][j]);
  }
}
// Few channels (Cp <= 32: the RGB image, the 19-class logit gradient): one thread per pixel reads its C planes (coalesced
// across the warp) and writes Cp / 8 16-byte vectors.  The 32x32 transpose tile above uses 3 of its 32 channel rows for
// an image: 573 us for the 100 MB input of the HRNet step against 25 us of traffic.
template <typename T, int CV>
__global__ void __launch_bounds__(256) nchw_to_nhwc_small_kernel(const T* __restrict__ src, uint4* __restrict__ dst,
                                                                 int C, long long HW, long long total) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long n = i / HW, hw = i - n * HW;
    const T* s = src + n * C * HW + hw;
    uint32_t w[CV * 4];
#pragma unroll
    for (int j = 0; j < CV * 4; ++j) {
      const float lo = 2 * j < C ? (float)s[(2 * j) * HW] : 0.f;
      const float hi = 2 * j + 1 < C ? (float)s[(2 * j + 1) * HW] : 0.f;
      w[j] = pack_bf16x2(lo, hi);
    }
#pragma unroll
    for (int v = 0; v < CV; ++v) dst[i * CV + v] = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
  }
}
// NHWC bf16 (pitch Cp) -> NCHW (fp32 or bf16), first C channels.
template <typename T>
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, T* __restrict__ dst, int C, int HW,
                                    int Cp) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float tile[32][33];
  const long long n = blockIdx.z;
  const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int hw = hw0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (hw < HW && c < C) ? __bfloat162float(src[(n * HW + hw) * Cp + c]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, hw = hw0 + threadIdx.x;
    if (c < C && hw < HW) dst[(n * C + c) * HW + hw] = (T)tile[threadIdx.x][j];
  }
}

// Stem input packing: NCHW image (C<=4 channels used) -> zero-padded 2x2 space-to-depth NHWC16 bf16.
//   dst[n, h2, w2, (dh*2+dw)*4 + c] = src[n, c, 2*h2+dh-pad, 2*w2+dw-pad]   (0 outside the image / for c >= C)
// A 7x7 stride-2 pad-3 convolution over src equals a 4x4 stride-1 pad-0 convolution over dst (16 channels); because
// the 4 horizontal taps are 4 adjacent NHWC16 pixels (64 contiguous elements) the conv kernel fetches them as ONE
// 64-"channel" im2col row, so the stem runs through the same tcgen05 pipeline as every other conv.
template <typename T>
__global__ void stem_s2d_pack_kernel(const T* __restrict__ src, uint4* __restrict__ dst, int N, int C, int H, int W,
                                     int H2, int W2, int pad) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const long long total = (long long)N * H2 * W2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w2 = (int)(i % W2);
    long long t = i / W2;
    const int h2 = (int)(t % H2);
    const long long n = t / H2;
    float f[16];
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const int h = 2 * h2 + dh - pad, w = 2 * w2 + dw - pad;
        const bool in = h >= 0 && h < H && w >= 0 && w < W;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          f[(dh * 2 + dw) * 4 + c] = (in && c < C) ? (float)src[((n * C + c) * H + h) * W + w] : 0.f;
      }
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = f[j];
      b[j] = f[8 + j];
    }
    dst[i * 2] = pack8(a);
    dst[i * 2 + 1] = pack8(b);
  }
}

// Stem weights: fp32 [K][7][7][C] (channels_last OIHW) -> bf16 [K][4][4][16] matching stem_s2d_pack_kernel.
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int K, int R,
                                        int S, int C) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int total = K * 256;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / 256, rem = i % 256;
    const int r2 = rem / 64, s2 = (rem / 16) % 4, e = rem % 16;
    const int dh = e / 8, dw = (e / 4) % 2, c = e % 4;
    const int r = 2 * r2 + dh, s = 2 * s2 + dw;
    float v = 0.f;
    if (r < R && s < S && c < C) v = w[((k * R + r) * S + s) * C + c];
    wp[i] = __float2bfloat16(v);
  }
}
// Inverse gather for the gradient: dw[K][7][7][C] (+)= dwp[K][4][4][16].
__global__ void stem_unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int K, int R, int S,
                                         int C, int accumulate) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int total = K * R * S * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C;
    int t = i / C;
    const int s = t % S;
    t /= S;
    const int r = t % R;
    const int k = t / R;
    const float v = dwp[k * 256 + (r / 2) * 64 + (s / 2) * 16 + ((r % 2) * 2 + (s % 2)) * 4 + c];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

// ------------------------------------------------------------------------------------------------ optimizer
// SGD with momentum (torch.optim.SGD semantics: d = g + wd*p ; buf = mu*buf + (1-damp)*d ; p -= lr*(nesterov ? d+mu*buf : buf))
// fused with the fp32 -> bf16 shadow-weight cast used by the next forward.  `first` = momentum buffer is uninitialised.
__global__ void sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                __nv_bfloat16* __restrict__ shadow, long long n, float lr, float mu, float wd,
                                float damp, int nesterov, float gscale, int first) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float w = p[i];
    float d = g[i] * gscale + wd * w;
    if (mu != 0.f) {
      const float b = first ? d : mu * buf[i] + (1.f - damp) * d;
      buf[i] = b;
      d = nesterov ? d + mu * b : b;
    }
    w -= lr * d;
    p[i] = w;
    if (shadow) shadow[i] = __float2bfloat16(w);
  }
}
// Adam / AdamW (torch.optim.Adam semantics, amsgrad off). bc1 = 1-beta1^t, bc2 = 1-beta2^t computed by the host.
__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, long long n, float lr,
                                 float b1, float b2, float eps, float wd, int decoupled, float bc1, float bc2,
                                 float gscale) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const float step = lr / bc1;
  const float rbc2 = rsqrtf(bc2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float w = p[i];
    float d = g[i] * gscale;
    if (decoupled)
      w *= 1.f - lr * wd;
    else
      d += wd * w;
    const float mi = b1 * m[i] + (1.f - b1) * d;
    const float vi = b2 * v[i] + (1.f - b2) * d * d;
    m[i] = mi;
    v[i] = vi;
    w -= step * mi / (sqrtf(vi) * rbc2 + eps);
    p[i] = w;
    if (shadow) shadow[i] = __float2bfloat16(w);
  }
}
// Graph-replayable variants: lr and the step counter are read from device memory (the counter is advanced by
// step_advance_kernel just before), and the consumed gradient is zeroed for the next step's atomic accumulation.
__global__ void step_advance_kernel(int* step) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch(); *step += 1; }
// per-parameter Adam step counts: advance where the parameter is being optimised (lr multiplier != 0)
__global__ void seg_steps_advance_kernel(int* steps, const float* lr_mult, const float* wd_mult, int n) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(lr_mult[i] == 0.f && wd_mult[i] == 0.f)) steps[i] += 1;
}

// The *_dev kernels process 4 parameters per thread with 16-byte accesses (the arenas are 256-byte aligned and padded
// to 64 elements); a scalar tail covers n % 4.
__global__ void __launch_bounds__(256)
sgd_step_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf,
                    __nv_bfloat16* __restrict__ shadow, long long n, const float* __restrict__ lr_dev,
                    const int* __restrict__ step_dev, float mu, float wd, float damp, int nesterov, float gscale,
                    int zero_grad, const ParamSegs segs) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  SgdArgs a;
  a.lr = __ldg(lr_dev);
  const float lr0 = a.lr;
  a.first = __ldg(step_dev) <= 1;
  a.mu = mu; a.wd = wd; a.damp = damp; a.gscale = gscale; a.nesterov = nesterov; a.zero_grad = zero_grad;
  const bool aligned = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)buf | ((uintptr_t)shadow << 1)) & 15) == 0;
  const long long n4 = aligned ? (n >> 2) : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    if (segs.n) {   // parameters start on 64-element boundaries, so a float4 never straddles two segments
      float lm, wm;
      seg_lookup(segs, i << 2, lm, wm);
      a.lr = lr0 * lm;
      a.wd = wd * wm;
      if (lm == 0.f && wm == 0.f) {   // frozen parameter: torch.optim skips it (momentum buffer untouched)
        if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
    }
    float4 w = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mu != 0.f && !a.first) b = reinterpret_cast<float4*>(buf)[i];
    w.x = sgd_one(w.x, gg.x, &b.x, a);
    w.y = sgd_one(w.y, gg.y, &b.y, a);
    w.z = sgd_one(w.z, gg.z, &b.z, a);
    w.w = sgd_one(w.w, gg.w, &b.w, a);
    if (mu != 0.f) reinterpret_cast<float4*>(buf)[i] = b;
    reinterpret_cast<float4*>(p)[i] = w;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16x2(w.x, w.y), pack_bf16x2(w.z, w.w));
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    if (segs.n) {
      float lm, wm;
      seg_lookup(segs, i, lm, wm);
      a.lr = lr0 * lm;
      a.wd = wd * wm;
      if (lm == 0.f && wm == 0.f) {
        if (zero_grad) g[i] = 0.f;
        continue;
      }
    }
    float b = (mu != 0.f && !a.first) ? buf[i] : 0.f;
    const float w = sgd_one(p[i], g[i], &b, a);
    if (mu != 0.f) buf[i] = b;
    p[i] = w;
    if (zero_grad) g[i] = 0.f;
    if (shadow) shadow[i] = __float2bfloat16(w);
  }
}
__global__ void __launch_bounds__(256)
adam_step_dev_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                     __nv_bfloat16* __restrict__ shadow, long long n, const float* __restrict__ lr_dev,
                     const int* __restrict__ step_dev, float b1, float b2, float eps, float wd, int decoupled,
                     float gscale, int zero_grad, const ParamSegs segs) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  AdamArgs a;
  a.lr = __ldg(lr_dev);
  const float lr0 = a.lr;
  const float t = (float)__ldg(step_dev);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  a.step = a.lr / bc1;
  a.rbc2 = rsqrtf(bc2);
  a.b1 = b1; a.b2 = b2; a.eps = eps; a.wd = wd; a.gscale = gscale; a.decoupled = decoupled;
  const bool aligned = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | ((uintptr_t)shadow << 1)) & 15) == 0;
  const long long n4 = aligned ? (n >> 2) : 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    if (segs.n) {
      float lm, wm;
      const int sg = seg_lookup(segs, i << 2, lm, wm);
      if (lm == 0.f && wm == 0.f) {   // frozen parameter: torch.optim.Adam skips it (exp_avg / exp_avg_sq / step untouched)
        if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      a.lr = lr0 * lm;
      a.wd = wd * wm;
      float c1 = bc1;
      if (segs.steps) {
        const float ts = (float)__ldg(segs.steps + sg);
        c1 = 1.f - powf(b1, ts);
        a.rbc2 = rsqrtf(1.f - powf(b2, ts));
      }
      a.step = a.lr / c1;
    }
    float4 w = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    w.x = adam_one(w.x, gg.x, &mm.x, &vv.x, a);
    w.y = adam_one(w.y, gg.y, &mm.y, &vv.y, a);
    w.z = adam_one(w.z, gg.z, &mm.z, &vv.z, a);
    w.w = adam_one(w.w, gg.w, &mm.w, &vv.w, a);
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(p)[i] = w;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) reinterpret_cast<uint2*>(shadow)[i] = make_uint2(pack_bf16x2(w.x, w.y), pack_bf16x2(w.z, w.w));
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    if (segs.n) {
      float lm, wm;
      const int sg = seg_lookup(segs, i, lm, wm);
      if (lm == 0.f && wm == 0.f) {
        if (zero_grad) g[i] = 0.f;
        continue;
      }
      a.lr = lr0 * lm;
      a.wd = wd * wm;
      float c1 = bc1;
      if (segs.steps) {
        const float ts = (float)__ldg(segs.steps + sg);
        c1 = 1.f - powf(b1, ts);
        a.rbc2 = rsqrtf(1.f - powf(b2, ts));
      }
      a.step = a.lr / c1;
    }
    float mi = m[i], vi = v[i];
    const float w = adam_one(p[i], g[i], &mi, &vi, a);
    m[i] = mi;
    v[i] = vi;
    p[i] = w;
    if (zero_grad) g[i] = 0.f;
    if (shadow) shadow[i] = __float2bfloat16(w);
  }
}
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

// Conv weights of a layer whose channel counts are not multiples of 8 (HRNet 18 / 36 / 72 ...), for the kernels that need
// the padded form: dst[k][t][in_pos(c)] = bf16(src[k][t][c]), zero elsewhere.  src: fp32 [K][T][C] (KRSC memory), dst:
// bf16 [Kp][T][Cp]; in_map (nullable): padded position of every input channel (concatenated, individually padded
// segments).  One launch instead of torch's zeros + cast + index_put.
__global__ void pad_weight_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int K, int T, int C,
                                  int Kp, int Cp, const int* __restrict__ inv_map, long long total) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cp = (int)(i % Cp);
    long long r = i / Cp;
    const int t = (int)(r % T);
    const int k = (int)(r / T);
    const int c = inv_map ? inv_map[cp] : (cp < C ? cp : -1);
    float v = 0.f;
    if (k < K && c >= 0) v = src[((long long)k * T + t) * C + c];
    dst[i] = __float2bfloat16(v);
  }
}
// The reverse for the weight gradient: dst[k][t][c] += src[k][t][in_pos(c)]  (fp32), one launch instead of slice +
// permute + add_.
__global__ void unpad_wgrad_add_kernel(const float* __restrict__ src, float* __restrict__ dst, int K, int T, int C,
                                       int Cp, const int* __restrict__ map, long long total) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;   // k * T + t
    const int cp = map ? map[c] : c;
    dst[i] += src[r * Cp + cp];
  }
}

}  // namespace tok

using namespace tok;

#define TOK_VEC_CHECK(c)                                                                         \
  if ((c) <= 0 || ((c) % 8) != 0) return set_error(TOK_ERR_INVALID, "channel count must be a positive multiple of 8 (got %d)", (int)(c))

extern "C" {

int tok_bn_finalize_train_cv(int C, int c_valid, double count, float* sum, float* sqsum, const float* gamma,
                             const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                             float* scale, float* shift, float* save_mean, float* save_invstd, void* stream) {
  if (C <= 0 || count <= 0 || c_valid <= 0 || c_valid > C) return set_error(TOK_ERR_INVALID, "bn_finalize: bad size");
  cudaError_t le = launch_pdl(bn_finalize_train_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, sum, sqsum,
                              (float)count, gamma, beta, eps, momentum, running_mean, running_var, scale, shift, save_mean,
                              save_invstd, C, c_valid);
  if (le != cudaSuccess) return set_error(TOK_ERR_CUDA, "bn_finalize_train: %s", cudaGetErrorString(le));
  return TOK_OK;
}
int tok_bn_finalize_train(int C, double count, float* sum, float* sqsum, const float* gamma,
                          const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                          float* scale, float* shift, float* save_mean, float* save_invstd, void* stream) {
  return tok_bn_finalize_train_cv(C, C, count, sum, sqsum, gamma, beta, eps, momentum, running_mean, running_var, scale,
                                  shift, save_mean, save_invstd, stream);
}

int tok_bn_finalize_eval_cv(int C, int c_valid, const float* running_mean, const float* running_var,
                            const float* gamma, const float* beta, float eps, float* scale, float* shift,
                            void* stream) {
  if (C <= 0 || c_valid <= 0 || c_valid > C) return set_error(TOK_ERR_INVALID, "bn_finalize: bad size");
  (void)launch_pdl(bn_finalize_eval_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, running_mean, running_var, gamma, beta,
                                                                               eps, scale, shift, C, c_valid);
  TOK_CHECK_LAUNCH("bn_finalize_eval");
  return TOK_OK;
}
int tok_bn_finalize_eval(int C, const float* running_mean, const float* running_var, const float* gamma,
                         const float* beta, float eps, float* scale, float* shift, void* stream) {
  return tok_bn_finalize_eval_cv(C, C, running_mean, running_var, gamma, beta, eps, scale, shift, stream);
}

static int launch_bn_apply(long long rows, int C, const void* y, const float* scale, const float* shift,
                           const void* residual, int relu, void* out, const ApplyFin& fin, void* stream) {
  TOK_VEC_CHECK(C);
  if (rows <= 0) return set_error(TOK_ERR_INVALID, "bn_apply: no rows");
  const int cvec = C / 8;
  const long long total = rows * cvec;
  // >= 4 vectors per thread (the unrolled loop), and a grid stride that is a multiple of the channel-vector count so
  // every thread keeps ONE set of scale / shift registers (HRNet's 3 / 5 / 9-vector rows included)
  static const int per_thread = getenv("TOK_BN_APPLY_VEC") ? atoi(getenv("TOK_BN_APPLY_VEC")) : 4;
  int grid = elem_grid(total, 256 * per_thread);
  {
    int a = cvec, b = 256;
    while (b) { const int t = a % b; a = b; b = t; }
    const int m = cvec / a;
    if (grid >= m) grid = grid / m * m;
  }
  if (fin.fused && ((long long)grid * 256) % cvec != 0)
    return set_error(TOK_ERR_INVALID, "bn_apply_train: C / 8 = %d does not divide the grid stride", cvec);
  cudaStream_t st = (cudaStream_t)stream;
  const uint4* yp = (const uint4*)y;
  const uint4* rp = (const uint4*)residual;
  uint4* op = (uint4*)out;
  cudaError_t le;
  if (residual && relu)
    le = launch_pdl(bn_apply_kernel<true, true>, dim3(grid), dim3(256), 0, st, yp, rp, op, scale, shift, total, cvec, fin);
  else if (residual)
    le = launch_pdl(bn_apply_kernel<true, false>, dim3(grid), dim3(256), 0, st, yp, rp, op, scale, shift, total, cvec, fin);
  else if (relu)
    le = launch_pdl(bn_apply_kernel<false, true>, dim3(grid), dim3(256), 0, st, yp, rp, op, scale, shift, total, cvec, fin);
  else
    le = launch_pdl(bn_apply_kernel<false, false>, dim3(grid), dim3(256), 0, st, yp, rp, op, scale, shift, total, cvec, fin);
  if (le != cudaSuccess) return set_error(TOK_ERR_CUDA, "bn_apply: %s", cudaGetErrorString(le));
  return TOK_OK;
}

int tok_bn_apply(long long rows, int C, const void* y, const float* scale, const float* shift, const void* residual,
                 int relu, void* out, void* stream) {
  ApplyFin fin;
  memset(&fin, 0, sizeof(fin));
  return launch_bn_apply(rows, C, y, scale, shift, residual, relu, out, fin, stream);
}

int tok_bn_apply_train_supported(long long rows, int C) {
  if (C <= 0 || (C % 8) || rows <= 0) return 0;
  const int cvec = C / 8;
  int grid = elem_grid(rows * cvec, 256 * 4);
  int a = cvec, b = 256;
  while (b) { const int t = a % b; a = b; b = t; }
  const int m = cvec / a;
  if (grid >= m) grid = grid / m * m;
  return ((long long)grid * 256) % cvec == 0 ? 1 : 0;
}

int tok_bn_apply_train(long long rows, int C, const void* y, float* sum, float* sqsum, const float* gamma,
                       const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                       float* scale, float* shift, float* save_mean, float* save_invstd, unsigned* counter,
                       const void* residual, int relu, void* out, void* stream) {
  if (!sum || !sqsum || !scale || !shift || !save_mean || !save_invstd || !counter)
    return set_error(TOK_ERR_INVALID, "bn_apply_train: accumulators, outputs and the ticket counter are required");
  ApplyFin fin;
  memset(&fin, 0, sizeof(fin));
  fin.fused = 1;
  fin.Cv = C;
  fin.counter = counter;
  fin.sum = sum;
  fin.sqsum = sqsum;
  fin.count = (float)rows;
  fin.eps = eps;
  fin.momentum = momentum;
  fin.gamma = gamma;
  fin.beta = beta;
  fin.running_mean = running_mean;
  fin.running_var = running_var;
  fin.scale = scale;
  fin.shift = shift;
  fin.save_mean = save_mean;
  fin.save_invstd = save_invstd;
  fin.C = C;
  return launch_bn_apply(rows, C, y, scale, shift, residual, relu, out, fin, stream);
}

int tok_bn_apply_chain(long long rows, int C, int c_valid, const void* y, float* sum, float* sqsum, const float* gamma,
                       const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                       float* scale, float* shift, float* save_mean, float* save_invstd, float* zero_ptr, int zero_n,
                       const void* residual, int relu, void* out, void* stream) {
  if (!sum || !sqsum || !scale || !shift || !save_mean || !save_invstd || c_valid <= 0 || c_valid > C || zero_n < 0 ||
      (zero_n > 0 && !zero_ptr))
    return set_error(TOK_ERR_INVALID, "bn_apply_chain: accumulators and outputs are required");
  ApplyFin fin;
  memset(&fin, 0, sizeof(fin));
  fin.fused = 1;
  fin.Cv = c_valid;
  fin.zero_ptr = zero_ptr;
  fin.zero_n = zero_n;
  fin.sum = sum;
  fin.sqsum = sqsum;
  fin.count = (float)rows;
  fin.eps = eps;
  fin.momentum = momentum;
  fin.gamma = gamma;
  fin.beta = beta;
  fin.running_mean = running_mean;
  fin.running_var = running_var;
  fin.scale = scale;
  fin.shift = shift;
  fin.save_mean = save_mean;
  fin.save_invstd = save_invstd;
  fin.C = C;
  return launch_bn_apply(rows, C, y, scale, shift, residual, relu, out, fin, stream);
}

int tok_bn_bwd_reduce(long long rows, int C, const void* dout, const void* dout2, const void* out, const void* y,
                      float* sum_g, float* sum_gy, void* stream) {
  TOK_VEC_CHECK(C);
  if (rows <= 0) return set_error(TOK_ERR_INVALID, "bn_bwd_reduce: no rows");
  const int cvec = C / 8;
  const int cvec_b = cvec < 256 ? cvec : 256;
  const int gy = (cvec + cvec_b - 1) / cvec_b;
  long long ctas = 148LL * 8 / gy;
  if (ctas < 1) ctas = 1;
  const int rlanes = 256 / cvec_b;
  long long min_rows = (long long)rlanes * 4;
  long long rows_per_cta = (rows + ctas - 1) / ctas;
  if (rows_per_cta < min_rows) rows_per_cta = min_rows;
  ctas = (rows + rows_per_cta - 1) / rows_per_cta;
  dim3 grid((unsigned)ctas, gy);
  (void)launch_pdl(bn_bwd_reduce_kernel, dim3(grid), dim3(256), cvec_b * 16 * sizeof(float), (cudaStream_t)stream, 
      (const uint4*)dout, (const uint4*)dout2, (const uint4*)out, (const uint4*)y, sum_g, sum_gy, rows, cvec, cvec_b,
      (int)rows_per_cta);
  TOK_CHECK_LAUNCH("bn_bwd_reduce");
  return TOK_OK;
}

int tok_bn_bwd_finalize_cv(int C, int c_valid, double count, float* sum_g, float* sum_gy, const float* save_mean,
                           const float* save_invstd, const float* gamma, float* coef_a, float* coef_c1,
                           float* coef_c0, float* dgamma, float* dbeta, int accumulate, void* stream) {
  if (C <= 0 || count <= 0 || c_valid <= 0 || c_valid > C) return set_error(TOK_ERR_INVALID, "bn_bwd_finalize: bad size");
  (void)launch_pdl(bn_bwd_finalize_kernel, dim3((C + 127) / 128), dim3(128), 0, (cudaStream_t)stream, 
      sum_g, sum_gy, save_mean, save_invstd, gamma, (float)count, coef_a, coef_c1, coef_c0, dgamma, dbeta, accumulate,
      C, c_valid);
  TOK_CHECK_LAUNCH("bn_bwd_finalize");
  return TOK_OK;
}
int tok_bn_bwd_finalize(int C, double count, float* sum_g, float* sum_gy, const float* save_mean,
                        const float* save_invstd, const float* gamma, float* coef_a, float* coef_c1, float* coef_c0,
                        float* dgamma, float* dbeta, int accumulate, void* stream) {
  return tok_bn_bwd_finalize_cv(C, C, count, sum_g, sum_gy, save_mean, save_invstd, gamma, coef_a, coef_c1, coef_c0,
                                dgamma, dbeta, accumulate, stream);
}

int tok_bn_bwd_apply(long long rows, int C, const void* dout, const void* dout2, const void* out, const void* y,
                     const float* coef_a, const float* coef_c1, const float* coef_c0, void* dy, void* dres,
                     void* stream) {
  TOK_VEC_CHECK(C);
  if (rows <= 0) return set_error(TOK_ERR_INVALID, "bn_bwd_apply: no rows");
  const int cvec = C / 8;
  const long long total = rows * cvec;
  (void)launch_pdl(bn_bwd_apply_kernel, dim3(elem_grid(total, 256 * 2)), dim3(256), 0, (cudaStream_t)stream, 
      (const uint4*)dout, (const uint4*)dout2, (const uint4*)out, (const uint4*)y, coef_a, coef_c1, coef_c0,
      (uint4*)dy, (uint4*)dres, total, cvec);
  TOK_CHECK_LAUNCH("bn_bwd_apply");
  return TOK_OK;
}

int tok_maxpool_fwd(int n, int h, int w, int c, int k, int s, int pad, const void* x, void* out, void* argmax,
                    void* stream) {
  TOK_VEC_CHECK(c);
  if (k <= 0 || k > 15 || s <= 0) return set_error(TOK_ERR_INVALID, "maxpool: bad window");
  const int P = (h + 2 * pad - k) / s + 1, Q = (w + 2 * pad - k) / s + 1;
  const long long total = (long long)n * P * Q * (c / 8);
  (void)launch_pdl(maxpool_fwd_kernel, dim3(elem_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const uint4*)x, (uint4*)out,
                                                                             (uint2*)argmax, total, h, w, P, Q, c / 8,
                                                                             k, s, pad);
  TOK_CHECK_LAUNCH("maxpool_fwd");
  return TOK_OK;
}

int tok_maxpool_bwd(int n, int h, int w, int c, int k, int s, int pad, const void* dout, const void* argmax, void* dx,
                    void* stream) {
  TOK_VEC_CHECK(c);
  if (k <= 0 || k > 15 || s <= 0) return set_error(TOK_ERR_INVALID, "maxpool: bad window");
  const int P = (h + 2 * pad - k) / s + 1, Q = (w + 2 * pad - k) / s + 1;
  const long long total = (long long)n * h * w * (c / 8);
  (void)launch_pdl(maxpool_bwd_kernel, dim3(elem_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const uint4*)dout, (const uint2*)argmax,
                                                                             (uint4*)dx, total, h, w, P, Q, c / 8, k,
                                                                             s, pad);
  TOK_CHECK_LAUNCH("maxpool_bwd");
  return TOK_OK;
}

int tok_gap_fwd(int n, int hw, int c, int mode, const void* x, void* out, void* stream) {
  TOK_VEC_CHECK(c);
  if (mode < 0 || mode > 2) return set_error(TOK_ERR_INVALID, "gap: mode must be 0 (avg), 1 (max) or 2 (avgmax)");
  const long long total = (long long)n * (c / 8);
  (void)launch_pdl(gap_fwd_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, (const uint4*)x, (uint4*)out, n,
                                                                                   hw, c / 8, mode);
  TOK_CHECK_LAUNCH("gap_fwd");
  return TOK_OK;
}

int tok_gap_bwd(int n, int hw, int c, const void* dout, void* dx, void* stream) {
  TOK_VEC_CHECK(c);
  const long long total = (long long)n * hw * (c / 8);
  (void)launch_pdl(gap_bwd_kernel, dim3(elem_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const uint4*)dout, (uint4*)dx, total, hw,
                                                                         c / 8);
  TOK_CHECK_LAUNCH("gap_bwd");
  return TOK_OK;
}

int tok_gap_bwd_max(int n, int hw, int c, int mode, const void* dout, const void* x, void* dx, void* stream) {
  TOK_VEC_CHECK(c);
  if (mode != 1 && mode != 2) return set_error(TOK_ERR_INVALID, "gap_bwd_max: mode must be 1 (max) or 2 (avgmax)");
  const long long total = (long long)n * (c / 8);
  (void)launch_pdl(gap_bwd_max_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, 
      (const uint4*)dout, (const uint4*)x, (uint4*)dx, n, hw, c / 8, mode);
  TOK_CHECK_LAUNCH("gap_bwd_max");
  return TOK_OK;
}

int tok_softmax_xent(int rows, int C, long long ld, const void* logits, const long long* target, float* loss_sum,
                     void* dlogits, float inv_norm, float gscale, const float* gscale_dev, long long ignore_index,
                     int* correct, void* stream) {
  if (rows <= 0 || C <= 0) return set_error(TOK_ERR_INVALID, "softmax_xent: bad size");
  (void)launch_pdl(softmax_xent_kernel, dim3(rows), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)logits, target, loss_sum,
                                                             (__nv_bfloat16*)dlogits, C, ld, inv_norm, gscale,
                                                             gscale_dev, ignore_index, correct);
  TOK_CHECK_LAUNCH("softmax_xent");
  return TOK_OK;
}

int tok_nchw_to_nhwc(int n, int c, int hw, int cp, int src_is_bf16, const void* src, void* dst, void* stream) {
  if (n <= 0 || c <= 0 || hw <= 0 || cp < c) return set_error(TOK_ERR_INVALID, "nchw_to_nhwc: bad size");
  if (cp <= 32 && cp % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const long long total = (long long)n * hw;
    const int g = elem_grid(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define TOK_SMALL(T, CV) (void)launch_pdl(nchw_to_nhwc_small_kernel<T, CV>, dim3(g), dim3(256), 0, st, (const T*)src, (uint4*)dst, c, hw, total)
    if (src_is_bf16) {
      if (cp == 8) TOK_SMALL(__nv_bfloat16, 1); else if (cp == 16) TOK_SMALL(__nv_bfloat16, 2);
      else if (cp == 24) TOK_SMALL(__nv_bfloat16, 3); else TOK_SMALL(__nv_bfloat16, 4);
    } else {
      if (cp == 8) TOK_SMALL(float, 1); else if (cp == 16) TOK_SMALL(float, 2);
      else if (cp == 24) TOK_SMALL(float, 3); else TOK_SMALL(float, 4);
    }
#undef TOK_SMALL
    TOK_CHECK_LAUNCH("nchw_to_nhwc");
    return TOK_OK;
  }
  dim3 grid((hw + 31) / 32, (cp + 31) / 32, n), block(32, 8);
  if (src_is_bf16)
    (void)launch_pdl(nchw_to_nhwc_kernel<__nv_bfloat16>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, (const __nv_bfloat16*)src,
                                                                                 (__nv_bfloat16*)dst, c, hw, cp);
  else
    (void)launch_pdl(nchw_to_nhwc_kernel<float>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, (const float*)src, (__nv_bfloat16*)dst, c, hw,
                                                                         cp);
  TOK_CHECK_LAUNCH("nchw_to_nhwc");
  return TOK_OK;
}

int tok_nhwc_to_nchw(int n, int c, int hw, int cp, int dst_is_bf16, const void* src, void* dst, void* stream) {
  if (n <= 0 || c <= 0 || hw <= 0 || cp < c) return set_error(TOK_ERR_INVALID, "nhwc_to_nchw: bad size");
  dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
  if (dst_is_bf16)
    (void)launch_pdl(nhwc_to_nchw_kernel<__nv_bfloat16>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, (const __nv_bfloat16*)src,
                                                                                 (__nv_bfloat16*)dst, c, hw, cp);
  else
    (void)launch_pdl(nhwc_to_nchw_kernel<float>, dim3(grid), dim3(block), 0, (cudaStream_t)stream, (const __nv_bfloat16*)src, (float*)dst, c, hw,
                                                                         cp);
  TOK_CHECK_LAUNCH("nhwc_to_nchw");
  return TOK_OK;
}

int tok_stem_pack_input(int n, int c, int h, int w, int src_is_bf16, const void* src, void* dst, void* stream) {
  if (n <= 0 || c <= 0 || c > 4 || h <= 0 || w <= 0) return set_error(TOK_ERR_INVALID, "stem_pack_input: needs 1..4 channels");
  const int P = (h - 1) / 2 + 1, Q = (w - 1) / 2 + 1, H2 = P + 3, W2 = Q + 3;
  const long long total = (long long)n * H2 * W2;
  if (src_is_bf16)
    (void)launch_pdl(stem_s2d_pack_kernel<__nv_bfloat16>, dim3(elem_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)src, (uint4*)dst, n, c, h, w, H2, W2, 3);
  else
    (void)launch_pdl(stem_s2d_pack_kernel<float>, dim3(elem_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const float*)src, (uint4*)dst,
                                                                                        n, c, h, w, H2, W2, 3);
  TOK_CHECK_LAUNCH("stem_pack_input");
  return TOK_OK;
}

int tok_stem_pack_weight(int k, int c, const float* w, void* wp, void* stream) {
  if (k <= 0 || c <= 0 || c > 4) return set_error(TOK_ERR_INVALID, "stem_pack_weight: needs 1..4 channels");
  (void)launch_pdl(stem_pack_weight_kernel, dim3((k * 256 + 255) / 256), dim3(256), 0, (cudaStream_t)stream, w, (__nv_bfloat16*)wp, k, 7, 7, c);
  TOK_CHECK_LAUNCH("stem_pack_weight");
  return TOK_OK;
}

int tok_stem_unpack_wgrad(int k, int c, const float* dwp, float* dw, int accumulate, void* stream) {
  if (k <= 0 || c <= 0 || c > 4) return set_error(TOK_ERR_INVALID, "stem_unpack_wgrad: needs 1..4 channels");
  (void)launch_pdl(stem_unpack_wgrad_kernel, dim3((k * 49 * c + 255) / 256), dim3(256), 0, (cudaStream_t)stream, dwp, dw, k, 7, 7, c,
                                                                                      accumulate);
  TOK_CHECK_LAUNCH("stem_unpack_wgrad");
  return TOK_OK;
}

int tok_sgd_step(long long n, float* param, const float* grad, float* momentum_buf, void* shadow_bf16, float lr,
                 float momentum, float weight_decay, float dampening, int nesterov, float grad_scale, int first_step,
                 void* stream) {
  if (n <= 0) return TOK_OK;
  (void)launch_pdl(sgd_step_kernel, dim3(elem_grid(n, 256 * 4)), dim3(256), 0, (cudaStream_t)stream, param, grad, momentum_buf,
                                                                          (__nv_bfloat16*)shadow_bf16, n, lr, momentum,
                                                                          weight_decay, dampening, nesterov,
                                                                          grad_scale, first_step);
  TOK_CHECK_LAUNCH("sgd_step");
  return TOK_OK;
}

int tok_adam_step(long long n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled, int step,
                  float grad_scale, void* stream) {
  if (n <= 0) return TOK_OK;
  if (step <= 0) return set_error(TOK_ERR_INVALID, "adam: step must be >= 1");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  (void)launch_pdl(adam_step_kernel, dim3(elem_grid(n, 256 * 4)), dim3(256), 0, (cudaStream_t)stream, 
      param, grad, exp_avg, exp_avg_sq, (__nv_bfloat16*)shadow_bf16, n, lr, beta1, beta2, eps, weight_decay, decoupled,
      bc1, bc2, grad_scale);
  TOK_CHECK_LAUNCH("adam_step");
  return TOK_OK;
}

static int check_segs(ParamSegs* s, const int* begin, const float* lr_mult, const float* wd_mult, int n, const char* who) {
  s->steps = nullptr;
  s->begin = begin;
  s->lr_mult = lr_mult;
  s->wd_mult = wd_mult;
  s->n = n;
  if (n < 0 || (n > 0 && (!begin || !lr_mult || !wd_mult)))
    return set_error(TOK_ERR_INVALID, "%s: a segment table needs begin / lr_mult / wd_mult arrays", who);
  return TOK_OK;
}

int tok_sgd_step_dev_groups(long long n, float* param, float* grad, float* momentum_buf, void* shadow_bf16,
                            const float* lr_dev, int* step_dev, float momentum, float weight_decay, float dampening,
                            int nesterov, float grad_scale, int zero_grad, const int* seg_begin,
                            const float* seg_lr_mult, const float* seg_wd_mult, int n_segs, void* stream) {
  if (n <= 0) return TOK_OK;
  if (!lr_dev || !step_dev) return set_error(TOK_ERR_INVALID, "sgd_step_dev: lr_dev and step_dev are required");
  ParamSegs segs;
  int rc = check_segs(&segs, seg_begin, seg_lr_mult, seg_wd_mult, n_segs, "sgd_step_dev_groups");
  if (rc) return rc;
  (void)launch_pdl(step_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, step_dev);
  (void)launch_pdl(sgd_step_dev_kernel, dim3(elem_grid(n, 256 * 4)), dim3(256), 0, (cudaStream_t)stream, 
      param, grad, momentum_buf, (__nv_bfloat16*)shadow_bf16, n, lr_dev, step_dev, momentum, weight_decay, dampening,
      nesterov, grad_scale, zero_grad, segs);
  TOK_CHECK_LAUNCH("sgd_step_dev");
  return TOK_OK;
}

int tok_sgd_step_dev(long long n, float* param, float* grad, float* momentum_buf, void* shadow_bf16,
                     const float* lr_dev, int* step_dev, float momentum, float weight_decay, float dampening,
                     int nesterov, float grad_scale, int zero_grad, void* stream) {
  return tok_sgd_step_dev_groups(n, param, grad, momentum_buf, shadow_bf16, lr_dev, step_dev, momentum, weight_decay,
                                 dampening, nesterov, grad_scale, zero_grad, nullptr, nullptr, nullptr, 0, stream);
}

int tok_adam_step_dev_groups(long long n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                             void* shadow_bf16, const float* lr_dev, int* step_dev, float beta1, float beta2, float eps,
                             float weight_decay, int decoupled, float grad_scale, int zero_grad, const int* seg_begin,
                             const float* seg_lr_mult, const float* seg_wd_mult, int* seg_steps, int n_segs,
                             void* stream) {
  if (n <= 0) return TOK_OK;
  if (!lr_dev || !step_dev) return set_error(TOK_ERR_INVALID, "adam_step_dev: lr_dev and step_dev are required");
  ParamSegs segs;
  int rc = check_segs(&segs, seg_begin, seg_lr_mult, seg_wd_mult, n_segs, "adam_step_dev_groups");
  if (rc) return rc;
  (void)launch_pdl(step_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, step_dev);
  if (seg_steps && n_segs > 0) {
    (void)launch_pdl(seg_steps_advance_kernel, dim3((n_segs + 255) / 256), dim3(256), 0, (cudaStream_t)stream, seg_steps, seg_lr_mult, seg_wd_mult,
                                                                                   n_segs);
    segs.steps = seg_steps;
  }
  (void)launch_pdl(adam_step_dev_kernel, dim3(elem_grid(n, 256 * 4)), dim3(256), 0, (cudaStream_t)stream, 
      param, grad, exp_avg, exp_avg_sq, (__nv_bfloat16*)shadow_bf16, n, lr_dev, step_dev, beta1, beta2, eps,
      weight_decay, decoupled, grad_scale, zero_grad, segs);
  TOK_CHECK_LAUNCH("adam_step_dev");
  return TOK_OK;
}

int tok_adam_step_dev(long long n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                      const float* lr_dev, int* step_dev, float beta1, float beta2, float eps, float weight_decay,
                      int decoupled, float grad_scale, int zero_grad, void* stream) {
  return tok_adam_step_dev_groups(n, param, grad, exp_avg, exp_avg_sq, shadow_bf16, lr_dev, step_dev, beta1, beta2, eps,
                                  weight_decay, decoupled, grad_scale, zero_grad, nullptr, nullptr, nullptr, nullptr, 0, stream);
}

int tok_pad_weight(int K, int T, int C, int Kp, int Cp, const float* src, const int* inv_map, void* dst, void* stream) {
  if (K <= 0 || T <= 0 || C <= 0 || Kp < K || Cp < C) return set_error(TOK_ERR_INVALID, "pad_weight: bad shape");
  const long long total = (long long)Kp * T * Cp;
  (void)launch_pdl(pad_weight_kernel, dim3(elem_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst, K, T, C, Kp, Cp,
                                                                             inv_map, total);
  TOK_CHECK_LAUNCH("pad_weight");
  return TOK_OK;
}

int tok_unpad_wgrad_add(int K, int T, int C, int Cp, const float* src, const int* map, float* dst, void* stream) {
  if (K <= 0 || T <= 0 || C <= 0 || Cp < C) return set_error(TOK_ERR_INVALID, "unpad_wgrad_add: bad shape");
  const long long total = (long long)K * T * C;
  (void)launch_pdl(unpad_wgrad_add_kernel, dim3(elem_grid(total, 256)), dim3(256), 0, (cudaStream_t)stream, src, dst, K, T, C, Cp, map, total);
  TOK_CHECK_LAUNCH("unpad_wgrad_add");
  return TOK_OK;
}

int tok_cast_f32_bf16(long long n, const float* src, void* dst, void* stream) {
  if (n <= 0) return TOK_OK;
  (void)launch_pdl(cast_f32_bf16_kernel, dim3(elem_grid(n, 256 * 4)), dim3(256), 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst, n);
  TOK_CHECK_LAUNCH("cast_f32_bf16");
  return TOK_OK;
}

}  // extern "C"
