// tok_heads.cu — warp-reduced kernels around the embedding heads and the pairwise loss.
//
// Reference call sites:
//   F.normalize in LinearHead (torchok/models/heads/representation/linear_head.py:33-35) and ArcFaceHead
//     (torchok/models/heads/classification/arcface_head.py:125-126)                       -> rownorm fwd / bwd
//   ArcFaceHead.__add_margin (arcface_head.py:95-108): phi = cos*cos m - sin*sin m on the target column, x scale
//                                                                                         -> arcface margin fwd / bwd
//   ContrastiveLoss.calc_loss (torchok/losses/representation/pairwise.py:126-136): S = cdist(emb1, emb2),
//     L_i = sum_j (1-R)relu(mu-S)^2 + R S^2                                               -> contrastive fwd / bwd
// The cosine GEMM itself (x_hat * w_hat^T, 256 x 512 x 11318 for SOP) runs on the tcgen05 linear kernels; what is
// here is the O(rows * D) work around it, one warp per row, fp32 arithmetic.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/tokb200.h"
#include "tok_internal.h"
#include "tok_ptx.cuh"

namespace tok {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ld_any(const void* p, long long i, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                 : reinterpret_cast<const float*>(p)[i];
}

// xhat[r, :] = scale * x[r, :] / max(|x[r]|, eps)  (bf16, row pitch ldo, pad columns zero); inv_norm[r] = 1/max(|x|,eps)
__global__ void __launch_bounds__(128)
rownorm_fwd_kernel(int rows, int d, const void* __restrict__ x, int x_is_bf16, float scale,
                   __nv_bfloat16* __restrict__ xhat, int ldo, float* __restrict__ inv_norm) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float ss = 0.f;
  for (int e = lane; e < d; e += 32) {
    const float v = ld_any(x, (long long)r * d + e, x_is_bf16);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int e = lane; e < ldo; e += 32) {
    const float v = e < d ? ld_any(x, (long long)r * d + e, x_is_bf16) * inv * scale : 0.f;
    xhat[(long long)r * ldo + e] = __float2bfloat16(v);
  }
  if (lane == 0 && inv_norm) inv_norm[r] = inv;
}

// Backward of y = scale * x/|x|:  dx = scale * inv_norm * (g - u (u.g)),  u = x/|x| recomputed from x.
// g: bf16 or fp32 with pitch ldg; dx: bf16 (stored) or fp32 (accumulated when `accumulate`).
__global__ void __launch_bounds__(128)
rownorm_bwd_kernel(int rows, int d, const void* __restrict__ x, int x_is_bf16, const float* __restrict__ inv_norm,
                   float scale, const void* __restrict__ g, int g_is_bf16, int ldg, void* __restrict__ dx,
                   int dx_is_bf16, int accumulate) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float inv = inv_norm[r];
  float ug = 0.f;
  for (int e = lane; e < d; e += 32)
    ug = fmaf(ld_any(x, (long long)r * d + e, x_is_bf16) * inv, ld_any(g, (long long)r * ldg + e, g_is_bf16), ug);
  ug = warp_sum(ug);
  for (int e = lane; e < d; e += 32) {
    const float u = ld_any(x, (long long)r * d + e, x_is_bf16) * inv;
    const float v = scale * inv * (ld_any(g, (long long)r * ldg + e, g_is_bf16) - u * ug);
    const long long o = (long long)r * d + e;
    if (dx_is_bf16) {
      reinterpret_cast<__nv_bfloat16*>(dx)[o] = __float2bfloat16(v);
    } else {
      float* p = reinterpret_cast<float*>(dx) + o;
      *p = accumulate ? *p + v : v;
    }
  }
}

struct Margin {
  float scale, cos_m, sin_m, th, mm;
  int easy;
};
__device__ __forceinline__ float margin_phi(float c, const Margin& m, float* dphi) {
  const float s2 = fminf(fmaxf(1.f - c * c, 0.f), 1.f);
  const float sine = sqrtf(s2);
  const float phi = c * m.cos_m - sine * m.sin_m;
  const bool take = m.easy ? (c > 0.f) : (c > m.th);
  if (dphi) *dphi = take ? (m.cos_m + m.sin_m * c / fmaxf(sine, 1e-6f)) : 1.f;
  return take ? phi : (m.easy ? c : c - m.mm);
}

// logits = scale * cosine arrives from the GEMM; the target column of every row is replaced by scale * phi(cos_t), with
// cos_t recomputed in fp32 from the normalised operands (xs = scale * x_hat, wh = w_hat), and saved for the backward.
__global__ void __launch_bounds__(128)
arcface_margin_fwd_kernel(int rows, int d, int ldx, const __nv_bfloat16* __restrict__ xs,
                          const __nv_bfloat16* __restrict__ wh, const long long* __restrict__ target, int num_classes,
                          __nv_bfloat16* __restrict__ logits, long long ldl, Margin m, float* __restrict__ cos_t) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const long long t = target[r];
  if (t < 0 || t >= num_classes) return;
  float acc = 0.f;
  for (int e = lane; e < d; e += 32)
    acc = fmaf(__bfloat162float(xs[(long long)r * ldx + e]), __bfloat162float(wh[t * ldx + e]), acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    const float c = fminf(fmaxf(acc / m.scale, -1.f), 1.f);
    logits[(long long)r * ldl + t] = __float2bfloat16(m.scale * margin_phi(c, m, nullptr));
    cos_t[r] = c;
  }
}
// dlogits[r, target] *= dphi/dcos
__global__ void arcface_margin_bwd_kernel(int rows, const long long* __restrict__ target, int num_classes,
                                          const float* __restrict__ cos_t, __nv_bfloat16* __restrict__ dlogits,
                                          long long ldl, Margin m) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const long long t = target[r];
  if (t < 0 || t >= num_classes) return;
  float dphi;
  margin_phi(cos_t[r], m, &dphi);
  const long long o = (long long)r * ldl + t;
  dlogits[o] = __float2bfloat16(__bfloat162float(dlogits[o]) * dphi);
}

// ---- contrastive loss --------------------------------------------------------------------------------------------
// CTA per row i of emb1; warp w takes j = w, w+8, ...; S_ij in fp32 by direct difference (no Gram cancellation).
__global__ void __launch_bounds__(256)
contrastive_fwd_kernel(int B, int M, int d, const float* __restrict__ e1, const float* __restrict__ e2,
                       const float* __restrict__ R, float margin, float* __restrict__ S, float* __restrict__ Lrow) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ float a[];  // emb1 row
  __shared__ float part[8];
  const int i = blockIdx.x;
  for (int e = threadIdx.x; e < d; e += 256) a[e] = e1[(long long)i * d + e];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int j = warp; j < M; j += 8) {
    float ss = 0.f;
    for (int e = lane; e < d; e += 32) {
      const float df = a[e] - e2[(long long)j * d + e];
      ss = fmaf(df, df, ss);
    }
    ss = warp_sum(ss);
    const float s = sqrtf(ss);
    if (lane == 0) {
      S[(long long)i * M + j] = s;
      const float r = R[(long long)i * M + j];
      const float h = fmaxf(margin - s, 0.f);
      acc += (1.f - r) * h * h + r * ss;
    }
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w];
    Lrow[i] = t;
  }
}
// coefficient of (a_i - b_j) in dL/da_i:  w_ij = gL_i * (2 R - 2 (1-R) relu(mu - S)/S)
__device__ __forceinline__ float pair_coef(float s, float r, float margin, float gl) {
  const float neg = s > 1e-12f ? fmaxf(margin - s, 0.f) / s : 0.f;
  return gl * (2.f * r - 2.f * (1.f - r) * neg);
}
// which == 0: CTA per row i, d_e1[i,:] = sum_j w_ij (a_i - b_j);  which == 1: CTA per row j, d_e2[j,:] = -sum_i w_ij (a_i - b_j)
__global__ void __launch_bounds__(256)
contrastive_bwd_kernel(int B, int M, int d, const float* __restrict__ e1, const float* __restrict__ e2,
                       const float* __restrict__ R, const float* __restrict__ S, const float* __restrict__ gL,
                       float margin, int which, float* __restrict__ out) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ float coef[];  // M (which 0) or B (which 1) coefficients
  const int row = blockIdx.x;
  const int n_other = which == 0 ? M : B;
  float csum_local = 0.f;
  for (int o = threadIdx.x; o < n_other; o += 256) {
    const int i = which == 0 ? row : o, j = which == 0 ? o : row;
    const float c = pair_coef(S[(long long)i * M + j], R[(long long)i * M + j], margin, gL[i]);
    coef[o] = c;
    csum_local += c;
  }
  __shared__ float red[256];
  red[threadIdx.x] = csum_local;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const float csum = red[0];
  const float* self = which == 0 ? e1 : e2;
  const float* other = which == 0 ? e2 : e1;
  const float sign = which == 0 ? 1.f : -1.f;
  for (int e = threadIdx.x; e < d; e += 256) {
    float acc = 0.f;
    for (int o = 0; o < n_other; ++o) acc = fmaf(coef[o], other[(long long)o * d + e], acc);
    // which 0: sum_j w (a_i - b_j) = a_i csum - acc ; which 1: -sum_i w (a_i - b_j) = -(acc - b_j csum)
    const float v = which == 0 ? self[(long long)row * d + e] * csum - acc : -(acc - self[(long long)row * d + e] * csum);
    (void)sign;
    out[(long long)row * d + e] = v;
  }
}

}  // namespace
}  // namespace tok

using namespace tok;

extern "C" {

int tok_rownorm_fwd(int rows, int d, const void* x, int x_is_bf16, float scale, void* xhat_bf16, int ld_out,
                    float* inv_norm, void* stream) {
  if (rows <= 0 || d <= 0 || ld_out < d) return set_error(TOK_ERR_INVALID, "rownorm_fwd: bad shape");
  const long long threads = (long long)rows * 32;
  (void)launch_pdl(rownorm_fwd_kernel, dim3((unsigned)((threads + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, 
      rows, d, x, x_is_bf16, scale, (__nv_bfloat16*)xhat_bf16, ld_out, inv_norm);
  TOK_CHECK_LAUNCH("rownorm_fwd");
  return TOK_OK;
}

int tok_rownorm_bwd(int rows, int d, const void* x, int x_is_bf16, const float* inv_norm, float scale, const void* g,
                    int g_is_bf16, int ld_g, void* dx, int dx_is_bf16, int accumulate, void* stream) {
  if (rows <= 0 || d <= 0 || ld_g < d) return set_error(TOK_ERR_INVALID, "rownorm_bwd: bad shape");
  const long long threads = (long long)rows * 32;
  (void)launch_pdl(rownorm_bwd_kernel, dim3((unsigned)((threads + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, 
      rows, d, x, x_is_bf16, inv_norm, scale, g, g_is_bf16, ld_g, dx, dx_is_bf16, accumulate);
  TOK_CHECK_LAUNCH("rownorm_bwd");
  return TOK_OK;
}

static Margin make_margin(float scale, float margin, int easy) {
  Margin m;
  m.scale = scale;
  m.cos_m = cosf(margin);
  m.sin_m = sinf(margin);
  m.th = cosf((float)M_PI - margin);
  m.mm = sinf((float)M_PI - margin) * margin;
  m.easy = easy;
  return m;
}

int tok_arcface_margin_fwd(int rows, int d, int ld_x, const void* xs_bf16, const void* wh_bf16,
                           const long long* target, int num_classes, void* logits_bf16, long long ld_logits,
                           float scale, float margin, int easy_margin, float* cos_t, void* stream) {
  if (rows <= 0 || d <= 0 || num_classes <= 0 || !cos_t) return set_error(TOK_ERR_INVALID, "arcface_margin_fwd: bad arguments");
  const long long threads = (long long)rows * 32;
  (void)launch_pdl(arcface_margin_fwd_kernel, dim3((unsigned)((threads + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, 
      rows, d, ld_x, (const __nv_bfloat16*)xs_bf16, (const __nv_bfloat16*)wh_bf16, target, num_classes,
      (__nv_bfloat16*)logits_bf16, ld_logits, make_margin(scale, margin, easy_margin), cos_t);
  TOK_CHECK_LAUNCH("arcface_margin_fwd");
  return TOK_OK;
}

int tok_arcface_margin_bwd(int rows, const long long* target, int num_classes, const float* cos_t,
                           void* dlogits_bf16, long long ld_logits, float scale, float margin, int easy_margin,
                           void* stream) {
  if (rows <= 0 || num_classes <= 0) return set_error(TOK_ERR_INVALID, "arcface_margin_bwd: bad arguments");
  (void)launch_pdl(arcface_margin_bwd_kernel, dim3((rows + 127) / 128), dim3(128), 0, (cudaStream_t)stream, 
      rows, target, num_classes, cos_t, (__nv_bfloat16*)dlogits_bf16, ld_logits, make_margin(scale, margin, easy_margin));
  TOK_CHECK_LAUNCH("arcface_margin_bwd");
  return TOK_OK;
}

int tok_contrastive_fwd(int B, int M, int d, const float* emb1, const float* emb2, const float* R, float margin,
                        float* S, float* loss_rows, void* stream) {
  if (B <= 0 || M <= 0 || d <= 0 || d > 8192) return set_error(TOK_ERR_INVALID, "contrastive_fwd: bad shape");
  (void)launch_pdl(contrastive_fwd_kernel, dim3(B), dim3(256), d * sizeof(float), (cudaStream_t)stream, B, M, d, emb1, emb2, R, margin, S,
                                                                             loss_rows);
  TOK_CHECK_LAUNCH("contrastive_fwd");
  return TOK_OK;
}

int tok_contrastive_bwd(int B, int M, int d, const float* emb1, const float* emb2, const float* R, const float* S,
                        const float* grad_rows, float margin, float* d_emb1, float* d_emb2, void* stream) {
  if (B <= 0 || M <= 0 || d <= 0 || B > 8192 || M > 8192) return set_error(TOK_ERR_INVALID, "contrastive_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (d_emb1)
    (void)launch_pdl(contrastive_bwd_kernel, dim3(B), dim3(256), M * sizeof(float), st, B, M, d, emb1, emb2, R, S, grad_rows, margin, 0, d_emb1);
  if (d_emb2)
    (void)launch_pdl(contrastive_bwd_kernel, dim3(M), dim3(256), B * sizeof(float), st, B, M, d, emb1, emb2, R, S, grad_rows, margin, 1, d_emb2);
  TOK_CHECK_LAUNCH("contrastive_bwd");
  return TOK_OK;
}

}  // extern "C"
