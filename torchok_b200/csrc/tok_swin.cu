// tok_swin.cu — Swin-V2 passes around the tcgen05 linear layers: LayerNorm (with the res-post-norm residual add), GELU,
// and shifted-window cosine attention.
//
// Reference call sites (timm 0.6.13 swin_transformer_v2, used through torchok/models/backbones/swin.py:71-81,127,156-171;
// semantics restated in SURVEY Appendix A.3):
//   SwinTransformerBlock.forward: x = x + drop_path(norm1(attn(x))); x = x + drop_path(norm2(mlp(x)))   -> layernorm_*
//   Mlp: fc1 -> GELU(erf) -> fc2                                                                          -> gelu_*
//   WindowAttention.forward: softmax(normalize(q) normalize(k)^T * exp(min(logit_scale, ln 100)) + 16 sigmoid(cpb) +
//     mask) v, on windows cut from the (cyclically shifted) token grid                                   -> window_attn_*
// The roll / window_partition / window_reverse copies of the reference are index arithmetic here: a window's tokens are
// gathered from and scattered to their home positions in the (B, H, W, 3C) qkv tensor.
//
// Forward attention runs on tcgen05 (window_attn_fwd_tc_kernel: two windows stacked into one 128-row UMMA tile, S and O
// in TMEM); the backward and the TOK_ATTN_CUDA_CORES=1 forward are CUDA-core kernels (one thread per row, window <= 8x8).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

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

// ---- LayerNorm -----------------------------------------------------------------------------------------------------
// y = LN(x) * gamma + beta ; out = (res ? res + rowscale[b] * y : y).  One warp per row, C <= 1024 (x kept in registers).
constexpr int kLnMaxPerLane = 32;  // C <= 1024; the kernels are instantiated for 4 / 8 / 16 / 32 elements per lane

template <int PER_LANE>
__global__ void __launch_bounds__(128)
layernorm_fwd_kernel(long long rows, int C, const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, const __nv_bfloat16* __restrict__ res,
                     const float* __restrict__ rowscale, int rows_per_sample, __nv_bfloat16* __restrict__ out,
                     float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float v[PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int c = lane + i * 32;
    v[i] = c < C ? __bfloat162float(x[r * C + c]) : 0.f;
    s += v[i];
  }
  const float mu = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int c = lane + i * 32;
    const float d = c < C ? v[i] - mu : 0.f;
    q = fmaf(d, d, q);
  }
  const float rs = rsqrtf(warp_sum(q) / C + eps);
  const float sc = rowscale ? rowscale[r / rows_per_sample] : 1.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < C) {
      float y = (v[i] - mu) * rs * gamma[c] + beta[c];
      if (res) y = __bfloat162float(res[r * C + c]) + sc * y;
      out[r * C + c] = __float2bfloat16(y);
    }
  }
  if (lane == 0) {
    mean[r] = mu;
    rstd[r] = rs;
  }
}

// dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)) with g = rowscale * dout (dres = dout passes through
// unchanged on the Python side); dgamma += sum g*xhat, dbeta += sum g (per-CTA partials in shared memory, then atomics).
template <int PER_LANE>
__global__ void __launch_bounds__(128)
layernorm_bwd_kernel(long long rows, int C, const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const __nv_bfloat16* __restrict__ dout, const float* __restrict__ rowscale, int rows_per_sample,
                     __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     int rows_per_cta) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ float sh[];  // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float ag[PER_LANE], ab[PER_LANE];
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) ag[i] = ab[i] = 0.f;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > rows) r1 = rows;
  for (long long r = r0 + warp; r < r1; r += 4) {
    const float mu = mean[r], rs = rstd[r];
    const float sc = rowscale ? rowscale[r / rows_per_sample] : 1.f;
    float xh[PER_LANE], gg[PER_LANE];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
      const int c = lane + i * 32;
      if (c < C) {
        xh[i] = (__bfloat162float(x[r * C + c]) - mu) * rs;
        const float g = sc * __bfloat162float(dout[r * C + c]);
        ag[i] = fmaf(g, xh[i], ag[i]);
        ab[i] += g;
        gg[i] = g * gamma[c];
        s1 += gg[i];
        s2 = fmaf(gg[i], xh[i], s2);
      } else {
        xh[i] = gg[i] = 0.f;
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
      const int c = lane + i * 32;
      if (c < C) dx[r * C + c] = __float2bfloat16(rs * (gg[i] - s1 - xh[i] * s2));
    }
  }
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int c = lane + i * 32;
    if (c < C) {
      atomicAdd(&sh[c], ag[i]);
      atomicAdd(&sh[C + c], ab[i]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + c, sh[c]);
    if (dbeta) atomicAdd(dbeta + c, sh[C + c]);
  }
}

// Vector path for C = 8 * CH * G (every Swin width: 96 * 2^k = 8 * 3 * G): a row is owned by G lanes, each holding CH
// 16-byte chunks (chunk index = lane_in_group + i * G, so a group reads consecutive 16-byte pieces); a warp works on
// 32 / G rows at once and walks the matrix with a grid stride.  Same arithmetic as the scalar kernels above.
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

template <int G, int CH>
__global__ void __launch_bounds__(256)
layernorm_fwd_vec_kernel(long long rows, const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, const __nv_bfloat16* __restrict__ res,
                         const float* __restrict__ rowscale, int rows_per_sample, __nv_bfloat16* __restrict__ out,
                         float* __restrict__ mean, float* __restrict__ rstd) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  constexpr int C = 8 * CH * G;
  constexpr int RPW = 32 / G;  // rows per warp pass
  const int lane = threadIdx.x & 31, l = lane % G, sub = lane / G;
  const long long warp_id = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  // the trip count is warp-uniform (the group shuffles need every lane); rows past the end are computed on row
  // `rows - 1` and not stored
  // software pipeline: the rows of the NEXT pass (x and the residual) are in flight while this pass reduces and stores
  // (r3: one load -> reduce -> load residual -> store chain per pass left the kernel at 0.58 of the HBM rate)
  uint4 nx[CH], nr[CH];
  auto fetch = [&](long long b) {
    const long long rr = b + sub < rows ? b + sub : rows - 1;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      nx[i] = __ldg(reinterpret_cast<const uint4*>(x + rr * C) + l + i * G);
      if (res) nr[i] = __ldg(reinterpret_cast<const uint4*>(res + rr * C) + l + i * G);
    }
  };
  if (warp_id * RPW < rows) fetch(warp_id * RPW);
  for (long long base = warp_id * RPW; base < rows; base += warps * RPW) {
    const bool live = base + sub < rows;
    const long long r = live ? base + sub : rows - 1;
    float v[CH][8];
    uint4 rraw[CH];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      unpack8(nx[i], v[i]);
      if (res) rraw[i] = nr[i];
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
    if (base + warps * RPW < rows) fetch(base + warps * RPW);
    const float mu = group_sum<G>(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mu;
        q = fmaf(d, d, q);
      }
    const float rs = rsqrtf(group_sum<G>(q) * (1.f / C) + eps);
    const float sc = rowscale ? rowscale[r / rows_per_sample] : 1.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c0 = (l + i * G) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float y[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) y[e] = fmaf((v[i][e] - mu) * rs, gm[e], bt[e]);
      if (res) {
        float rr[8];
        unpack8(rraw[i], rr);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = fmaf(sc, y[e], rr[e]);
      }
      if (live) reinterpret_cast<uint4*>(out + r * C)[l + i * G] = pack8(y);
    }
    if (l == 0 && live) {
      mean[r] = mu;
      rstd[r] = rs;
    }
  }
}

// Backward, vector path.  Per-thread column partials (dgamma, dbeta and, when asked, the column sums of dx = the bias
// gradient of the linear layer that produced x) stay in registers over the whole grid-stride walk.
template <int G, int CH>
__global__ void __launch_bounds__(128, CH >= 3 ? 3 : 4)
layernorm_bwd_vec_kernel(long long rows, const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                         const float* __restrict__ mean, const float* __restrict__ rstd,
                         const __nv_bfloat16* __restrict__ dout, const float* __restrict__ rowscale, int rows_per_sample,
                         __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                         float* __restrict__ dxsum) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  constexpr int C = 8 * CH * G;
  constexpr int RPW = 32 / G;
  extern __shared__ __align__(16) float sh[];  // [warps][3][C]: column-sum slices, written once at the end
  const int lane = threadIdx.x & 31, l = lane % G, sub = lane / G;
  const long long warp_id = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  // the row stays in registers as the packed bf16 it arrived in (x, dout: 8 words per chunk); gamma comes from L1
  float ag[CH][8], ab[CH][8], ax[CH][8];
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) ag[i][e] = ab[i][e] = ax[i][e] = 0.f;
  for (long long base = warp_id * RPW; base < rows; base += warps * RPW) {   // warp-uniform trip count (shuffles)
    const bool live = base + sub < rows;
    const long long r = live ? base + sub : rows - 1;
    uint4 xr[CH], gr[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      xr[i] = __ldg(reinterpret_cast<const uint4*>(x + r * C) + l + i * G);
      gr[i] = live ? __ldg(reinterpret_cast<const uint4*>(dout + r * C) + l + i * G) : make_uint4(0, 0, 0, 0);
    }
    const float mu = mean[r], rs = rstd[r];
    const float sc = rowscale ? rowscale[r / rows_per_sample] : 1.f;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c0 = (l + i * G) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float xv[8], gv[8];
      unpack8(xr[i], xv);
      unpack8(gr[i], gv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (xv[e] - mu) * rs;
        const float g = sc * gv[e];
        ag[i][e] = fmaf(g, xh, ag[i][e]);
        ab[i][e] += g;
        const float gg = g * gm[e];
        s1 += gg;
        s2 = fmaf(gg, xh, s2);
      }
    }
    s1 = group_sum<G>(s1) * (1.f / C);
    s2 = group_sum<G>(s2) * (1.f / C);
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c0 = (l + i * G) * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      float xv[8], gv[8], d[8];
      unpack8(xr[i], xv);
      unpack8(gr[i], gv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (xv[e] - mu) * rs;
        d[e] = rs * (sc * gv[e] * gm[e] - s1 - xh * s2);
        ax[i][e] += d[e];
      }
      if (live) reinterpret_cast<uint4*>(dx + r * C)[l + i * G] = pack8(d);
    }
  }
  // Column sums of the CTA.  The lanes of a warp that share a column group (same l, RPW of them) add up with shuffles,
  // each warp owns a [3][C] slice of shared memory (plain stores), the slices are summed and leave with 16-byte vector
  // reds.  r5: the float atomicAdd on shared memory this replaces is a load / add / compare-and-swap spin loop — 72 of them
  // per thread with up to 8-way contention at the end of every CTA, which dominated the stage 3 / 4 launches (2-4 row
  // iterations per warp).
  const int wslice = (threadIdx.x >> 5) * 3 * C;
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float a = ag[i][e], b = ab[i][e], x2 = ax[i][e];
#pragma unroll
      for (int s2 = G; s2 < 32; s2 <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, s2);
        b += __shfl_xor_sync(0xffffffffu, b, s2);
        if (dxsum) x2 += __shfl_xor_sync(0xffffffffu, x2, s2);
      }
      if (sub == 0) {
        const int c = (l + i * G) * 8 + e;
        sh[wslice + c] = a;
        sh[wslice + C + c] = b;
        sh[wslice + 2 * C + c] = x2;
      }
    }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  for (int c4 = threadIdx.x; c4 < 3 * C / 4; c4 += blockDim.x) {   // C % 8 == 0: a float4 never straddles two of the arrays
    float4 t = *reinterpret_cast<const float4*>(sh + c4 * 4);
    for (int w = 1; w < nw; ++w) {
      const float4 u = *reinterpret_cast<const float4*>(sh + w * 3 * C + c4 * 4);
      t.x += u.x;
      t.y += u.y;
      t.z += u.z;
      t.w += u.w;
    }
    const int which = (c4 * 4) / C, c = c4 * 4 - which * C;
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dxsum);
    if (dst == nullptr) continue;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w)
                   : "memory");
    } else {   // a gradient slice that does not start on a 16-byte boundary
      atomicAdd(dst + c, t.x);
      atomicAdd(dst + c + 1, t.y);
      atomicAdd(dst + c + 2, t.z);
      atomicAdd(dst + c + 3, t.w);
    }
  }
}

// (G, CH) with C == 8 * CH * G, G in {4, 8, 16, 32}, CH in {3, 4, 2, 1}; false if the width has no vector path
inline bool ln_vec_shape(int C, int* G, int* CH) {
  if (C % 8) return false;
  const int chunks = C / 8;
  const int chs[4] = {3, 4, 2, 1};
  for (int i = 0; i < 4; ++i) {
    if (chunks % chs[i]) continue;
    const int g = chunks / chs[i];
    if (g == 4 || g == 8 || g == 16 || g == 32) {
      *G = g;
      *CH = chs[i];
      return true;
    }
  }
  return false;
}

// ---- GELU (exact, erf): gelu_parts / gelu_value / gelu_grad live in tok_ptx.cuh (shared with the linear-dgrad epilogue)
__global__ void gelu_fwd_kernel(const __nv_bfloat162* __restrict__ x, __nv_bfloat162* __restrict__ y, long long n2) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const float2 v = __bfloat1622float2(x[i]);
    y[i] = __floats2bfloat162_rn(gelu_value(v.x), gelu_value(v.y));
  }
}
__global__ void gelu_bwd_kernel(const __nv_bfloat162* __restrict__ x, const __nv_bfloat162* __restrict__ dy,
                                __nv_bfloat162* __restrict__ dx, long long n2) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
    const float2 v = __bfloat1622float2(x[i]);
    const float2 g = __bfloat1622float2(dy[i]);
    dx[i] = __floats2bfloat162_rn(g.x * gelu_grad(v.x), g.y * gelu_grad(v.y));
  }
}
// 16-byte version (n % 8 == 0)
__global__ void __launch_bounds__(256)
gelu_fwd_vec_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long n8) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n8; i += 4 * stride) {   // four independent 16-byte loads in flight per thread
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = __ldg(x + i + k * stride);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8];
      unpack8(u[k], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = gelu_value(v[e]);
      y[i + k * stride] = pack8(v);
    }
  }
  for (; i < n8; i += stride) {
    float v[8];
    unpack8(__ldg(x + i), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = gelu_value(v[e]);
    y[i] = pack8(v);
  }
}
// Backward over a (rows, C) matrix, C % 128 == 0: CTA = 16 chunk-columns x 16 row lanes; a thread keeps its 8 columns
// for every row it visits, so the column sums of dx (the bias gradient of the linear layer that produced x) are
// register partials, merged through shared memory once per CTA.
__global__ void __launch_bounds__(256)
gelu_bwd_vec_kernel(long long rows, int C, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                    __nv_bfloat16* __restrict__ dx, float* __restrict__ dbias) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float sh[16][129];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int c0 = blockIdx.x * 128 + tx * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const long long rstep = gridDim.y * 16LL;
  long long r = blockIdx.y * 16LL + ty;
  for (; r + rstep < rows; r += 2 * rstep) {   // two rows (four 16-byte loads) in flight per thread
    const uint4 xa = __ldg(reinterpret_cast<const uint4*>(x + r * C + c0));
    const uint4 ga = __ldg(reinterpret_cast<const uint4*>(dy + r * C + c0));
    const uint4 xb = __ldg(reinterpret_cast<const uint4*>(x + (r + rstep) * C + c0));
    const uint4 gb = __ldg(reinterpret_cast<const uint4*>(dy + (r + rstep) * C + c0));
    float v[8], g[8];
    unpack8(xa, v);
    unpack8(ga, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = g[e] * gelu_grad(v[e]);
      acc[e] += v[e];
    }
    *reinterpret_cast<uint4*>(dx + r * C + c0) = pack8(v);
    unpack8(xb, v);
    unpack8(gb, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = g[e] * gelu_grad(v[e]);
      acc[e] += v[e];
    }
    *reinterpret_cast<uint4*>(dx + (r + rstep) * C + c0) = pack8(v);
  }
  for (; r < rows; r += rstep) {
    float v[8], g[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * C + c0)), v);
    unpack8(__ldg(reinterpret_cast<const uint4*>(dy + r * C + c0)), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = g[e] * gelu_grad(v[e]);
      acc[e] += v[e];
    }
    *reinterpret_cast<uint4*>(dx + r * C + c0) = pack8(v);
  }
  if (dbias) {
#pragma unroll
    for (int e = 0; e < 8; ++e) sh[ty][tx * 8 + e] = acc[e];
    __syncthreads();
    if (threadIdx.x < 128) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) t += sh[j][threadIdx.x];
      atomicAdd(dbias + blockIdx.x * 128 + threadIdx.x, t);
    }
  }
}

// ---- patch embedding ----------------------------------------------------------------------------------------------
// timm PatchEmbed (torchok/models/backbones/swin.py:156-171 builds it): Conv2d(3, E, kernel 4, stride 4) on the NCHW fp32
// image, flatten(2).transpose(1, 2).  K = 3 * 4 * 4 = 48 is too thin for the tensor-core conv (it ran at 578 us plus an
// 870 us NCHW->NHWC pass for 0.07 GFLOP/img); here one CTA takes one row of patches: the 12 image rows it needs are read
// coalesced straight from NCHW into a [tokens][48] shared-memory patch matrix, thread (o, half) keeps row o of the
// weight matrix in registers and walks half of the tokens with broadcast 16-byte shared loads.  fp32 math, bf16 out.
constexpr int kPeK = 48;       // in_chans * patch * patch
constexpr int kPePitch = 52;   // floats per patch row in shared memory (16-byte aligned, conflict-free 16-byte stores)

__device__ __forceinline__ void pe_load_patches(const float* __restrict__ img, int b, int i, int H, int W, int TW,
                                                float* __restrict__ sp) {
  // image rows (c, 4i + dy), c < 3, dy < 4: W floats each, one float4 = the 4 dx of one (c, dy, token)
  const int per_row = W / 4;
  for (int idx = threadIdx.x; idx < 12 * per_row; idx += blockDim.x) {
    const int rowi = idx / per_row, j = idx - rowi * per_row;
    const int c = rowi >> 2, dy = rowi & 3;
    const float4 v = __ldg(reinterpret_cast<const float4*>(img + (((long long)b * 3 + c) * H + 4 * i + dy) * W) + j);
    *reinterpret_cast<float4*>(sp + j * kPePitch + c * 16 + dy * 4) = v;
  }
}

__global__ void __launch_bounds__(512)
patch_embed_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                       int B, int H, int W, int E, int sO, int sC, int sH, int sW, __nv_bfloat16* __restrict__ out) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ __align__(16) float sp[];
  const int TH = H / 4, TW = W / 4;
  const int o = threadIdx.x % E, half = threadIdx.x / E;
  // persistent CTA: the weight row is fetched once (it was 48 uncoalesced loads per thread and patch row otherwise)
  float wr[kPeK];
#pragma unroll
  for (int k = 0; k < kPeK; ++k) wr[k] = __ldg(w + (long long)o * sO + (k >> 4) * sC + ((k >> 2) & 3) * sH + (k & 3) * sW);
  const float bo = bias ? __ldg(bias + o) : 0.f;
  const int t0 = half * ((TW + 1) / 2), t1 = min(TW, t0 + (TW + 1) / 2);
  for (int rowi = blockIdx.x; rowi < B * TH; rowi += gridDim.x) {
    const int b = rowi / TH, i = rowi - b * TH;
    __syncthreads();   // the previous row's patches are consumed
    pe_load_patches(img, b, i, H, W, TW, sp);
    __syncthreads();
    for (int tk = t0; tk < t1; ++tk) {
      float acc = bo;
#pragma unroll
      for (int k4 = 0; k4 < kPeK / 4; ++k4) {
        const float4 pv = *reinterpret_cast<const float4*>(sp + tk * kPePitch + k4 * 4);
        acc = fmaf(wr[k4 * 4], pv.x, acc);
        acc = fmaf(wr[k4 * 4 + 1], pv.y, acc);
        acc = fmaf(wr[k4 * 4 + 2], pv.z, acc);
        acc = fmaf(wr[k4 * 4 + 3], pv.w, acc);
      }
      out[((long long)rowi * TW + tk) * E + o] = __float2bfloat16(acc);
    }
  }
}

// dW[o][k] += sum_tokens dy[token][o] * patch[token][k], dbias[o] += sum dy: thread (o, half) accumulates its 48 + 1
// values in registers over every patch row its CTA visits (the row of dy is staged in shared memory next to the
// patches, so no thread waits on a global load inside the token loop); halves merge in shared memory, CTAs with atomics.
__global__ void __launch_bounds__(512)
patch_embed_bwd_kernel(const float* __restrict__ img, const __nv_bfloat16* __restrict__ dy, int B, int H, int W, int E,
                       int sO, int sC, int sH, int sW, float* __restrict__ dw, float* __restrict__ dbias) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ __align__(16) float sp[];
  const int TH = H / 4, TW = W / 4;
  __nv_bfloat16* sg = reinterpret_cast<__nv_bfloat16*>(sp + TW * kPePitch);   // [TW][E]
  const int o = threadIdx.x % E, half = threadIdx.x / E;
  float acc[kPeK];
#pragma unroll
  for (int k = 0; k < kPeK; ++k) acc[k] = 0.f;
  float accb = 0.f;
  const int t0 = half * ((TW + 1) / 2), t1 = min(TW, t0 + (TW + 1) / 2);
  for (int rowi = blockIdx.x; rowi < B * TH; rowi += gridDim.x) {
    const int b = rowi / TH, i = rowi - b * TH;
    __syncthreads();   // the previous row's patches are consumed
    pe_load_patches(img, b, i, H, W, TW, sp);
    {
      const uint4* src = reinterpret_cast<const uint4*>(dy + (long long)rowi * TW * E);
      for (int idx = threadIdx.x; idx < TW * E / 8; idx += blockDim.x) reinterpret_cast<uint4*>(sg)[idx] = __ldg(src + idx);
    }
    __syncthreads();
    for (int tk = t0; tk < t1; ++tk) {
      const float g = __bfloat162float(sg[tk * E + o]);
      accb += g;
#pragma unroll
      for (int k4 = 0; k4 < kPeK / 4; ++k4) {
        const float4 pv = *reinterpret_cast<const float4*>(sp + tk * kPePitch + k4 * 4);
        acc[k4 * 4] = fmaf(g, pv.x, acc[k4 * 4]);
        acc[k4 * 4 + 1] = fmaf(g, pv.y, acc[k4 * 4 + 1]);
        acc[k4 * 4 + 2] = fmaf(g, pv.z, acc[k4 * 4 + 2]);
        acc[k4 * 4 + 3] = fmaf(g, pv.w, acc[k4 * 4 + 3]);
      }
    }
  }
  // merge the two halves through shared memory (the buffer is sized for max(patch rows + dy row, 49 * E floats))
  __syncthreads();
  if (half == 1) {
#pragma unroll
    for (int k = 0; k < kPeK; ++k) sp[k * E + o] = acc[k];
    sp[kPeK * E + o] = accb;
  }
  __syncthreads();
  if (half == 0) {
#pragma unroll
    for (int k = 0; k < kPeK; ++k)
      atomicAdd(dw + (long long)o * sO + (k >> 4) * sC + ((k >> 2) & 3) * sH + (k & 3) * sW, acc[k] + sp[k * E + o]);
    if (dbias) atomicAdd(dbias + o, accb + sp[kPeK * E + o]);
  }
}

// ---- patch embedding on the tensor cores (r3): im2col of the non-overlapping P x P patches of an NCHW image into a bf16
// [B*H/P*W/P][P*P*C] matrix in the weight's own memory order (dy, dx, c) — the [E][P][P][C] conv weight IS the [E][K]
// matrix of a linear layer then, so timm's PatchEmbed.proj (torchok/models/backbones/swin.py:71-81) runs as tok_linear_fwd /
// tok_linear_wgrad.  One thread per token row of P pixels x C channels; 503 + 357 us of CUDA-core convolution per Swin-T
// step become ~50 us of layout pass plus two epilogue-bound GEMMs.
template <typename T, int P, int C>
__global__ void __launch_bounds__(256)
patchify_kernel(const T* __restrict__ img, __nv_bfloat16* __restrict__ dst, int B, int H, int W, long long total) {
  pdl_wait();
  pdl_launch();
  const int Wp = W / P, Hp = H / P;
  constexpr int K = P * P * C;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    // i = ((b * Hp + ph) * Wp + pw) * P + dy : one image row segment of a patch per thread, P pixels x C channels
    const int dy = (int)(i % P);
    long long t = i / P;
    const int pw = (int)(t % Wp);
    t /= Wp;
    const int ph = (int)(t % Hp);
    const long long b = t / Hp;
    float v[C][P];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const T* s = img + ((b * C + c) * H + (ph * P + dy)) * (long long)W + pw * P;
#pragma unroll
      for (int dx = 0; dx < P; ++dx) v[c][dx] = (float)s[dx];
    }
    __nv_bfloat16* d = dst + ((b * Hp + ph) * (long long)Wp + pw) * K + dy * (P * C);
    float f[P * C];
#pragma unroll
    for (int dx = 0; dx < P; ++dx)
#pragma unroll
      for (int c = 0; c < C; ++c) f[dx * C + c] = v[c][dx];
    static_assert((P * C) % 4 == 0, "a row segment is stored as 8-byte vectors");
#pragma unroll
    for (int j = 0; j < P * C / 4; ++j)   // byte offset dy * 2 * P * C: 8-byte aligned
      reinterpret_cast<uint2*>(d)[j] = make_uint2(pack_bf16x2(f[4 * j], f[4 * j + 1]), pack_bf16x2(f[4 * j + 2], f[4 * j + 3]));
  }
}

// ---- PatchMerging gather -----------------------------------------------------------------------------------------
// timm PatchMerging (used by torchok/models/backbones/swin.py:71-81 through BasicLayer.downsample):
//   cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)  on a (B, H, W, C) tensor.
// A pure permutation in 16-byte pieces: out[b, i, j, q*C + c] = x[b, 2i + (q & 1), 2j + (q >> 1), c]; the backward is
// the same walk with source and destination swapped (the reference pays 8 strided slices, a cat and, in backward,
// 8 zero-fills + 8 strided copies + 3 accumulations for it).
__global__ void __launch_bounds__(256)
patch_merge_kernel(int B, int H, int W, int C8, const uint4* __restrict__ src, uint4* __restrict__ dst, int inverse) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int H2 = H / 2, W2 = W / 2;
  const long long total = (long long)B * H2 * W2 * 4 * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long t = i / C8;
    const int q = (int)(t & 3);
    t >>= 2;
    const int j = (int)(t % W2);
    t /= W2;
    const int ii = (int)(t % H2);
    const long long b = t / H2;
    const long long xi = ((b * H + 2 * ii + (q & 1)) * W + 2 * j + (q >> 1)) * C8 + c;
    if (inverse) dst[xi] = __ldg(src + i);
    else dst[i] = __ldg(src + xi);
  }
}

// ---- shifted-window cosine attention ---------------------------------------------------------------------------------
constexpr int kHd = 32;      // head dimension (always 32 in Swin-V2: embed_dim 96 / 3 heads, doubling together)
constexpr int kMaxN = 64;    // tokens per window (window <= 8x8)
constexpr int kPitch = kHd + 4;  // shared-memory row pitch in floats: 16-byte aligned rows for LDS.128 broadcasts

__device__ __forceinline__ float dot32(const float (&a)[kHd], const float* __restrict__ row) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < kHd; e += 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + e);
    s = fmaf(a[e], v.x, s);
    s = fmaf(a[e + 1], v.y, s);
    s = fmaf(a[e + 2], v.z, s);
    s = fmaf(a[e + 3], v.w, s);
  }
  return s;
}
__device__ __forceinline__ void axpy32(float (&acc)[kHd], float a, const float* __restrict__ row) {
#pragma unroll
  for (int e = 0; e < kHd; e += 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + e);
    acc[e] = fmaf(a, v.x, acc[e]);
    acc[e + 1] = fmaf(a, v.y, acc[e + 1]);
    acc[e + 2] = fmaf(a, v.z, acc[e + 2]);
    acc[e + 3] = fmaf(a, v.w, acc[e + 3]);
  }
}

struct AttnGeom {
  int B, H, W, C, heads, ws, shift;
  int nwy, nwx;  // windows per axis
};

// home position (row in the (B*H*W, 3C) qkv matrix) of local token t of window (b, wy, wx)
__device__ __forceinline__ long long token_row(const AttnGeom& g, int b, int wy, int wx, int t, int& region) {
  const int ty = t / g.ws, tx = t - ty * g.ws;
  const int ys = wy * g.ws + ty, xs = wx * g.ws + tx;  // coordinates in the shifted grid
  int ry = 0, rx = 0;
  if (g.shift > 0) {
    ry = ys < g.H - g.ws ? 0 : (ys < g.H - g.shift ? 1 : 2);
    rx = xs < g.W - g.ws ? 0 : (xs < g.W - g.shift ? 1 : 2);
  }
  region = ry * 3 + rx;
  int y = ys + g.shift, x = xs + g.shift;
  if (y >= g.H) y -= g.H;
  if (x >= g.W) x -= g.W;
  return ((long long)b * g.H + y) * g.W + x;
}

// One CTA (64 threads) per (window, head); thread i owns query row i.  Saves nothing: the backward recomputes.
__global__ void __launch_bounds__(64)
window_attn_fwd_kernel(AttnGeom g, const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ logit_scale,
                       const float* __restrict__ bias /* [heads][N][N] */, __nv_bfloat16* __restrict__ out) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ __align__(16) float sk[kMaxN][kPitch], sv[kMaxN][kPitch];
  __shared__ int sreg[kMaxN];
  const int N = g.ws * g.ws;
  const int head = blockIdx.x % g.heads;
  int w = blockIdx.x / g.heads;
  const int wx = w % g.nwx;
  w /= g.nwx;
  const int wy = w % g.nwy;
  const int b = w / g.nwy;
  const int i = threadIdx.x;
  float q[kHd];
  long long my_row = 0;
  if (i < N) {
    int region;
    my_row = token_row(g, b, wy, wx, i, region);
    sreg[i] = region;
    const __nv_bfloat16* base = qkv + my_row * 3 * g.C + head * kHd;
    float qq = 0.f, kk = 0.f;
#pragma unroll
    for (int e = 0; e < kHd; ++e) {
      q[e] = __bfloat162float(base[e]);
      qq = fmaf(q[e], q[e], qq);
      const float kv = __bfloat162float(base[g.C + e]);
      sk[i][e] = kv;
      kk = fmaf(kv, kv, kk);
      sv[i][e] = __bfloat162float(base[2 * g.C + e]);
    }
    const float qi = 1.f / fmaxf(sqrtf(qq), 1e-12f), ki = 1.f / fmaxf(sqrtf(kk), 1e-12f);
#pragma unroll
    for (int e = 0; e < kHd; ++e) {
      q[e] *= qi;
      sk[i][e] *= ki;
    }
  }
  __syncthreads();
  if (i >= N) return;
  const float scale = __expf(fminf(logit_scale[head], 4.6051702f));  // ln(100)
  const float* brow = bias + ((long long)head * N + i) * N;
  const int my_reg = sreg[i];
  float p[kMaxN];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kMaxN; ++j) {
    if (j < N) {
      float s = dot32(q, sk[j]);
      s = s * scale + brow[j] + (sreg[j] != my_reg ? -100.f : 0.f);
      p[j] = s;
      mx = fmaxf(mx, s);
    }
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxN; ++j) {
    if (j < N) {
      p[j] = __expf(p[j] - mx);
      l += p[j];
    }
  }
  const float inv = 1.f / l;
  float o[kHd];
#pragma unroll
  for (int e = 0; e < kHd; ++e) o[e] = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxN; ++j) {
    if (j < N) {
      axpy32(o, p[j] * inv, sv[j]);
    }
  }
  __nv_bfloat16* op = out + my_row * g.C + head * kHd;
#pragma unroll
  for (int e = 0; e < kHd; ++e) op[e] = __float2bfloat16(o[e]);
}

// ------------------------------------------------------------------------------------------------------------------
// Windows LARGER than 8x8 (N = 144 / 256 / 576 tokens: swinv2_*_window12_192, *_window16_256, *_window12to16/24_*,
// torchok/models/backbones/swin.py:293-402; the reference's own backbone test uses swinv2_tiny_window16_256).  One CTA of
// 256 threads per (window, head); K-hat and V (forward) of the whole window live in shared memory as fp32, a thread owns a
// query row and folds the keys with an online softmax, so nothing of size N x N is ever stored.  The backward makes three
// sweeps (row statistics + output, query gradients, key / value gradients) with the resident pair of operands swapped
// between the second and the third.  CUDA cores: this path exists for coverage of the registered large-window variants;
// the tcgen05 kernels above serve windows up to 8x8 (all of Swin-T/S/B at 224 / 256 with window 7 / 8).
constexpr int kBigMaxN = 576;
constexpr int kBigThreads = 256;
constexpr int kBigSmem = (2 * kBigMaxN * kPitch + 4 * kBigMaxN) * 4 + kBigMaxN * 4 + 32 * 4 + 64;

__device__ __forceinline__ void load_row32(const __nv_bfloat16* p, float (&v)[kHd]) {
#pragma unroll
  for (int e = 0; e < kHd; e += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(p + e);
    v[e] = bf16_lo(u.x); v[e + 1] = bf16_hi(u.x);
    v[e + 2] = bf16_lo(u.y); v[e + 3] = bf16_hi(u.y);
    v[e + 4] = bf16_lo(u.z); v[e + 5] = bf16_hi(u.z);
    v[e + 6] = bf16_lo(u.w); v[e + 7] = bf16_hi(u.w);
  }
}
__device__ __forceinline__ void store_row32(__nv_bfloat16* p, const float (&v)[kHd]) {
#pragma unroll
  for (int e = 0; e < kHd; e += 8)
    *reinterpret_cast<uint4*>(p + e) = make_uint4(pack_bf16x2(v[e], v[e + 1]), pack_bf16x2(v[e + 2], v[e + 3]),
                                                  pack_bf16x2(v[e + 4], v[e + 5]), pack_bf16x2(v[e + 6], v[e + 7]));
}
__device__ __forceinline__ float normalize32(float (&v)[kHd]) {   // v <- v / max(|v|, eps); returns 1 / max(|v|, eps)
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < kHd; ++e) ss = fmaf(v[e], v[e], ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
  for (int e = 0; e < kHd; ++e) v[e] *= inv;
  return inv;
}

__global__ void __launch_bounds__(kBigThreads)
window_attn_big_fwd_kernel(AttnGeom g, const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ logit_scale,
                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ __align__(16) float dyn[];
  const int N = g.ws * g.ws;
  float (*sk)[kPitch] = reinterpret_cast<float (*)[kPitch]>(dyn);
  float (*sv)[kPitch] = sk + N;
  int* sreg = reinterpret_cast<int*>(sv + N);
  const int head = blockIdx.x % g.heads;
  int w = blockIdx.x / g.heads;
  const int wx = w % g.nwx;
  w /= g.nwx;
  const int wy = w % g.nwy;
  const int b = w / g.nwy;
  for (int t = threadIdx.x; t < N; t += kBigThreads) {
    int region;
    const long long row = token_row(g, b, wy, wx, t, region);
    sreg[t] = region;
    const __nv_bfloat16* base = qkv + row * 3 * g.C + head * kHd;
    float kv[kHd], vv[kHd];
    load_row32(base + g.C, kv);
    load_row32(base + 2 * g.C, vv);
    normalize32(kv);
#pragma unroll
    for (int e = 0; e < kHd; ++e) {
      sk[t][e] = kv[e];
      sv[t][e] = vv[e];
    }
  }
  __syncthreads();
  const float scale = __expf(fminf(logit_scale[head], 4.6051702f));  // ln(100)
  for (int i = threadIdx.x; i < N; i += kBigThreads) {
    int region;
    const long long row = token_row(g, b, wy, wx, i, region);
    float q[kHd];
    load_row32(qkv + row * 3 * g.C + head * kHd, q);
    normalize32(q);
    const float* brow = bias + ((long long)head * N + i) * N;
    float mx = -INFINITY, l = 0.f, o[kHd];
#pragma unroll
    for (int e = 0; e < kHd; ++e) o[e] = 0.f;
    for (int j = 0; j < N; ++j) {
      const float sc = dot32(q, sk[j]) * scale + __ldg(brow + j) + (sreg[j] != region ? -100.f : 0.f);
      if (sc > mx) {   // rescale the running sums to the new maximum
        const float c = __expf(mx - sc);
        l *= c;
#pragma unroll
        for (int e = 0; e < kHd; ++e) o[e] *= c;
        mx = sc;
      }
      const float pr = __expf(sc - mx);
      l += pr;
      axpy32(o, pr, sv[j]);
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int e = 0; e < kHd; ++e) o[e] *= inv;
    store_row32(out + row * g.C + head * kHd, o);
  }
}

__global__ void __launch_bounds__(kBigThreads)
window_attn_big_bwd_kernel(AttnGeom g, const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ logit_scale,
                           const float* __restrict__ bias, const __nv_bfloat16* __restrict__ dout,
                           __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dbias,
                           float* __restrict__ dlogit_scale) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ __align__(16) float dyn[];
  const int N = g.ws * g.ws;
  float (*sa)[kPitch] = reinterpret_cast<float (*)[kPitch]>(dyn);   // sweep 1-2: K-hat     sweep 3: Q-hat
  float (*sb)[kPitch] = sa + N;                                      // sweep 1-2: V         sweep 3: dO
  float* smx = reinterpret_cast<float*>(sb + N);
  float* sl = smx + N;
  float* sdelta = sl + N;
  float* sqn = sdelta + N;         // unused tail kept for alignment of the ints below
  int* sreg = reinterpret_cast<int*>(sqn + N);
  float* s_red = reinterpret_cast<float*>(sreg + N);
  const int head = blockIdx.x % g.heads;
  int w = blockIdx.x / g.heads;
  const int wx = w % g.nwx;
  w /= g.nwx;
  const int wy = w % g.nwy;
  const int b = w / g.nwy;
  const float raw_ls = logit_scale[head];
  const float scale = __expf(fminf(raw_ls, 4.6051702f));
  for (int t = threadIdx.x; t < N; t += kBigThreads) {
    int region;
    const long long row = token_row(g, b, wy, wx, t, region);
    sreg[t] = region;
    const __nv_bfloat16* base = qkv + row * 3 * g.C + head * kHd;
    float kv[kHd], vv[kHd];
    load_row32(base + g.C, kv);
    load_row32(base + 2 * g.C, vv);
    normalize32(kv);
#pragma unroll
    for (int e = 0; e < kHd; ++e) {
      sa[t][e] = kv[e];
      sb[t][e] = vv[e];
    }
  }
  __syncthreads();
  float dls = 0.f;   // sum over this thread's rows of ds * cos
  // ---- sweeps 1 + 2: thread = query row
  for (int i = threadIdx.x; i < N; i += kBigThreads) {
    int region;
    const long long row = token_row(g, b, wy, wx, i, region);
    float q[kHd], go[kHd];
    load_row32(qkv + row * 3 * g.C + head * kHd, q);
    const float qinv = normalize32(q);
    load_row32(dout + row * g.C + head * kHd, go);
    const float* brow = bias + ((long long)head * N + i) * N;
    float mx = -INFINITY, l = 0.f, o[kHd];
#pragma unroll
    for (int e = 0; e < kHd; ++e) o[e] = 0.f;
    for (int j = 0; j < N; ++j) {
      const float sc = dot32(q, sa[j]) * scale + __ldg(brow + j) + (sreg[j] != region ? -100.f : 0.f);
      if (sc > mx) {
        const float c = __expf(mx - sc);
        l *= c;
#pragma unroll
        for (int e = 0; e < kHd; ++e) o[e] *= c;
        mx = sc;
      }
      const float pr = __expf(sc - mx);
      l += pr;
      axpy32(o, pr, sb[j]);
    }
    const float linv = 1.f / l;
    float delta = 0.f;   // sum_j p_ij dp_ij = dO_i . O_i
#pragma unroll
    for (int e = 0; e < kHd; ++e) delta = fmaf(go[e], o[e] * linv, delta);
    smx[i] = mx;
    sl[i] = linv;
    sdelta[i] = delta;
    float dq[kHd];
#pragma unroll
    for (int e = 0; e < kHd; ++e) dq[e] = 0.f;
    float* db = dbias + ((long long)head * N + i) * N;
    for (int j = 0; j < N; ++j) {
      const float cs = dot32(q, sa[j]);
      const float sc = cs * scale + __ldg(brow + j) + (sreg[j] != region ? -100.f : 0.f);
      const float pr = __expf(sc - mx) * linv;
      const float ds = pr * (dot32(go, sb[j]) - delta);
      atomicAdd(db + j, ds);
      dls = fmaf(ds, cs, dls);
      axpy32(dq, ds * scale, sa[j]);
    }
    // through q-hat = q / |q|
    float qd = 0.f;
#pragma unroll
    for (int e = 0; e < kHd; ++e) qd = fmaf(q[e], dq[e], qd);
#pragma unroll
    for (int e = 0; e < kHd; ++e) dq[e] = (dq[e] - q[e] * qd) * qinv;
    store_row32(dqkv + row * 3 * g.C + head * kHd, dq);
  }
  // logit_scale gradient of this (window, head)
  dls = warp_sum(dls);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dls;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < kBigThreads / 32; ++k) t += s_red[k];
    if (raw_ls < 4.6051702f) atomicAdd(dlogit_scale + head, t * scale);
  }
  // ---- sweep 3: thread = key row; resident operands become Q-hat and dO
  for (int t = threadIdx.x; t < N; t += kBigThreads) {
    int region;
    const long long row = token_row(g, b, wy, wx, t, region);
    float q[kHd], go[kHd];
    load_row32(qkv + row * 3 * g.C + head * kHd, q);
    normalize32(q);
    load_row32(dout + row * g.C + head * kHd, go);
#pragma unroll
    for (int e = 0; e < kHd; ++e) {
      sa[t][e] = q[e];
      sb[t][e] = go[e];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < N; j += kBigThreads) {
    int region;
    const long long row = token_row(g, b, wy, wx, j, region);
    const __nv_bfloat16* base = qkv + row * 3 * g.C + head * kHd;
    float k[kHd], v[kHd];
    load_row32(base + g.C, k);
    const float kinv = normalize32(k);
    load_row32(base + 2 * g.C, v);
    float dk[kHd], dv[kHd];
#pragma unroll
    for (int e = 0; e < kHd; ++e) dk[e] = dv[e] = 0.f;
    const float* bcol = bias + (long long)head * N * N + j;
    for (int i = 0; i < N; ++i) {
      const float sc = dot32(k, sa[i]) * scale + __ldg(bcol + (long long)i * N) + (sreg[i] != region ? -100.f : 0.f);
      const float pr = __expf(sc - smx[i]) * sl[i];
      const float ds = pr * (dot32(v, sb[i]) - sdelta[i]);
      axpy32(dk, ds * scale, sa[i]);
      axpy32(dv, pr, sb[i]);
    }
    float kd = 0.f;
#pragma unroll
    for (int e = 0; e < kHd; ++e) kd = fmaf(k[e], dk[e], kd);
#pragma unroll
    for (int e = 0; e < kHd; ++e) dk[e] = (dk[e] - k[e] * kd) * kinv;
    store_row32(dqkv + row * 3 * g.C + g.C + head * kHd, dk);
    store_row32(dqkv + row * 3 * g.C + 2 * g.C + head * kHd, dv);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// tcgen05 forward: QK^T -> scale + bias + mask -> softmax -> PV in ONE kernel, accumulators in TMEM.
// A CTA (128 threads, thread = row) owns one head and walks pairs of windows.  The two windows of a pair are stacked
// along M (rows 0-63 / 64-127, 49 or 64 valid tokens each):
//   S[128 x 128] = Qhat[128 x 32] . Khat[128 x 32]^T      one UMMA chain (2 x K16); only the two diagonal 64 x 64 blocks
//                                                          are meaningful (query and key of the same window)
//   softmax per row in registers straight from TMEM (tcgen05.ld), P written to shared memory as bf16
//   O[128 x 64]  = [P_A; 0] . V_A + [0; P_B] . V_B         two K64 blocks: block-diagonal P against the per-window V
// Operand tiles are written by the threads themselves in the 128B-swizzled K-major / MN-major layouts the UMMA
// descriptors expect (the same layouts TMA produces for the conv kernels); rows are 128 bytes of which 64 carry data.
constexpr int kTcThreads = 128;
constexpr int kTcSmem = 16384 * 2 + 8192 * 2 + 16384 * 2 + kMaxN * kMaxN * 4 + 128 * 4 + 2 * 8 + 16 + 1024;

__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(kTcThreads)
window_attn_fwd_tc_kernel(AttnGeom g, int groups, const __nv_bfloat16* __restrict__ qkv,
                          const float* __restrict__ logit_scale, const float* __restrict__ bias,
                          __nv_bfloat16* __restrict__ out) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // [128][128 B]  K-major A of S
  uint8_t* sK = sQ + 16384;           // [128][128 B]  K-major B of S
  uint8_t* sV = sK + 16384;           // 2 x [64 keys][128 B]  MN-major B of O (cols 32..63 stay zero)
  uint8_t* sP = sV + 16384;           // 2 x [128][128 B]  K-major A of O (block-diagonal halves)
  float* sbias = reinterpret_cast<float*>(sP + 32768);  // [N][N] of this head
  int* sreg = reinterpret_cast<int*>(sbias + kMaxN * kMaxN);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sreg + 128);  // [0] S ready, [1] O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);

  const int N = g.ws * g.ws;
  const int head = blockIdx.x % g.heads;
  const int grp = blockIdx.x / g.heads;
  const int r = threadIdx.x;
  const int warp = r >> 5;
  const int prob = r >> 6;        // which window of the pair
  const int t = r & 63;           // token inside the window
  const bool tok_ok = t < N;
  const float scale = __expf(fminf(logit_scale[head], 4.6051702f));
  const int total_windows = g.B * g.nwy * g.nwx;
  const int total_pairs = (total_windows + 1) / 2;

  // one-time: zero every operand tile (pad rows / pad columns are never written again), barriers, TMEM, bias table
  for (int i = r; i < (16384 * 2 + 16384 + 32768) / 16; i += kTcThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = r; i < N * N; i += kTcThreads) sbias[i] = bias[(long long)head * N * N + i];
  if (r == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot;        // S: columns [0, 128)
  const uint32_t tmem_o = tmem_s + 128;      // O: columns [128, 192)
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);
  const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);

  // rows of a pair: the next pair's q / k / v travel while the current one is in the tensor cores and the softmax
  uint4 qraw[4], kraw[4], vraw[4];
  long long nrow = 0;
  int nregion = 0;
  bool nvalid = false;
  auto fetch = [&](int pair) {
    const int w = pair * 2 + prob;
    nvalid = tok_ok && pair < total_pairs && w < total_windows;
    nrow = 0;
    nregion = 0;
    if (nvalid) {
      const int wx = w % g.nwx, wy = (w / g.nwx) % g.nwy, b = w / (g.nwx * g.nwy);
      nrow = token_row(g, b, wy, wx, t, nregion);
      const uint4* base = reinterpret_cast<const uint4*>(qkv + nrow * 3 * g.C + head * kHd);
      const int cstep = g.C / 8;  // uint4 per C bf16
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        qraw[c] = __ldg(base + c);
        kraw[c] = __ldg(base + cstep + c);
        vraw[c] = __ldg(base + 2 * cstep + c);
      }
    }
  };
  fetch(grp);

  uint32_t phase = 0;
  for (int pair = grp; pair < total_pairs; pair += groups, phase ^= 1) {
    // ---- this row's token: normalise q / k, write the operand rows
    const bool valid = nvalid;
    const long long my_row = nrow;
    const int region = nregion;
    if (valid) {
      float qq = 0.f, kk = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t qw[4] = {qraw[c].x, qraw[c].y, qraw[c].z, qraw[c].w};
        const uint32_t kw[4] = {kraw[c].x, kraw[c].y, kraw[c].z, kraw[c].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          qq = fmaf(bf16_lo(qw[e]), bf16_lo(qw[e]), fmaf(bf16_hi(qw[e]), bf16_hi(qw[e]), qq));
          kk = fmaf(bf16_lo(kw[e]), bf16_lo(kw[e]), fmaf(bf16_hi(kw[e]), bf16_hi(kw[e]), kk));
        }
      }
      const float qi = 1.f / fmaxf(sqrtf(qq), 1e-12f), ki = 1.f / fmaxf(sqrtf(kk), 1e-12f);
      // Split-bf16 cosine: q_hat = q_hi + q_lo, k_hat = k_hi + k_lo (bf16 each); S = q_hi.k_hi + q_lo.k_hi + q_hi.k_lo
      // keeps the logits at ~2^-16 relative error although the logit scale multiplies them by up to 100.  The low
      // halves ride in the 64 padding bytes of the operand rows: sQ = [q_hi | q_lo], sK = [k_hi | k_hi],
      // sV = [v | k_lo] (P.V then also produces 32 unused output columns).
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t qw[4] = {qraw[c].x, qraw[c].y, qraw[c].z, qraw[c].w};
        const uint32_t kw[4] = {kraw[c].x, kraw[c].y, kraw[c].z, kraw[c].w};
        uint32_t qo[4], ko[4], ql[4], kl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float q0 = bf16_lo(qw[e]) * qi, q1 = bf16_hi(qw[e]) * qi;
          const float k0 = bf16_lo(kw[e]) * ki, k1 = bf16_hi(kw[e]) * ki;
          qo[e] = pack_bf16x2(q0, q1);
          ko[e] = pack_bf16x2(k0, k1);
          ql[e] = pack_bf16x2(q0 - bf16_lo(qo[e]), q1 - bf16_hi(qo[e]));
          kl[e] = pack_bf16x2(k0 - bf16_lo(ko[e]), k1 - bf16_hi(ko[e]));
        }
        sts128(aQ + sw128_off(r, c), make_uint4(qo[0], qo[1], qo[2], qo[3]));
        sts128(aQ + sw128_off(r, c + 4), make_uint4(ql[0], ql[1], ql[2], ql[3]));
        sts128(aK + sw128_off(r, c), make_uint4(ko[0], ko[1], ko[2], ko[3]));
        sts128(aK + sw128_off(r, c + 4), make_uint4(ko[0], ko[1], ko[2], ko[3]));
        sts128(aV + sw128_off(r, c), vraw[c]);
        sts128(aV + sw128_off(r, c + 4), make_uint4(kl[0], kl[1], kl[2], kl[3]));
      }
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        sts128(aQ + sw128_off(r, c), make_uint4(0, 0, 0, 0));
        sts128(aK + sw128_off(r, c), make_uint4(0, 0, 0, 0));
        sts128(aV + sw128_off(r, c), make_uint4(0, 0, 0, 0));
      }
    }
    sreg[r] = region;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- S = Qhat . Khat^T
    if (r == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)   // [q_hi | q_lo] . [k_hi | k_hi]
        umma_bf16(tmem_s, make_smem_desc_sw128(aQ + k * 32, 16, 1024), make_smem_desc_sw128(aK + k * 32, 16, 1024),
                  idesc_s, k);
#pragma unroll
      for (int k = 0; k < 2; ++k)   // q_hi . k_lo (k_lo lives in the second half of the V rows)
        umma_bf16(tmem_s, make_smem_desc_sw128(aQ + k * 32, 16, 1024), make_smem_desc_sw128(aV + (k + 2) * 32, 16, 1024),
                  idesc_s, 1u);
      umma_commit(&bar[0]);
    }
    fetch(pair + groups);
    mbar_wait(&bar[0], phase);
    tc_fence_after();
    // ---- softmax of this row over its own window's keys
    float p[kMaxN];
    {
      uint32_t raw[32];
      float mx = -INFINITY;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tmem_ld_32x32b_x32(tmem_s + lane_addr + prob * 64 + half * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int key = half * 32 + j;
          float sc = -INFINITY;
          if (key < N && valid)
            sc = __uint_as_float(raw[j]) * scale + sbias[t * N + key] + (sreg[prob * 64 + key] != region ? -100.f : 0.f);
          p[key] = sc;
          mx = fmaxf(mx, sc);
        }
      }
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxN; ++j) {
        p[j] = (j < N && valid) ? __expf(p[j] - mx) : 0.f;
        l += p[j];
      }
      const float inv = valid ? 1.f / l : 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t o4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o4[e] = pack_bf16x2(p[c * 8 + 2 * e] * inv, p[c * 8 + 2 * e + 1] * inv);
        sts128(aP + prob * 16384 + sw128_off(r, c), make_uint4(o4[0], o4[1], o4[2], o4[3]));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    // ---- O = [P_A; 0] . V_A + [0; P_B] . V_B
    if (r == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_o, make_smem_desc_sw128(aP + kb * 16384 + k * 32, 16, 1024),
                    make_smem_desc_sw128(aV + kb * 8192 + k * 2048, 8192, 1024), idesc_o, (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], phase);
    tc_fence_after();
    {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(tmem_o + lane_addr, raw);
      tmem_ld_wait();
      if (valid) {
        uint4* op = reinterpret_cast<uint4*>(out + my_row * g.C + head * kHd);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          op[c] = make_uint4(pack_bf16x2(__uint_as_float(raw[c * 8]), __uint_as_float(raw[c * 8 + 1])),
                             pack_bf16x2(__uint_as_float(raw[c * 8 + 2]), __uint_as_float(raw[c * 8 + 3])),
                             pack_bf16x2(__uint_as_float(raw[c * 8 + 4]), __uint_as_float(raw[c * 8 + 5])),
                             pack_bf16x2(__uint_as_float(raw[c * 8 + 6]), __uint_as_float(raw[c * 8 + 7])));
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM and the operand tiles are free for the next pair
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem_s, 256);
}

// Walks the windows w0, w0 + step, w0 + 2 step, ... of one row slot t of the stacked pair without per-iteration divisions
// (r5: the five runtime divisions of token_row cost 800-1700 clocks per pair on the critical path of the tcgen05
// kernels): (b, wy, wx) advance by the pre-split step with carries, (ty, tx) are fixed per thread.
struct WindowWalker {
  int b, wy, wx, sb, swy, swx, ty, tx;
  __device__ __forceinline__ void init(const AttnGeom& g, int w0, int step, int t) {
    wx = w0 % g.nwx;
    wy = (w0 / g.nwx) % g.nwy;
    b = w0 / (g.nwx * g.nwy);
    swx = step % g.nwx;
    swy = (step / g.nwx) % g.nwy;
    sb = step / (g.nwx * g.nwy);
    ty = t / g.ws;
    tx = t - ty * g.ws;
  }
  __device__ __forceinline__ bool in_range(const AttnGeom& g) const { return b < g.B; }
  __device__ __forceinline__ long long row(const AttnGeom& g, int& region) const {
    const int ys = wy * g.ws + ty, xs = wx * g.ws + tx;
    int ry = 0, rx = 0;
    if (g.shift > 0) {
      ry = ys < g.H - g.ws ? 0 : (ys < g.H - g.shift ? 1 : 2);
      rx = xs < g.W - g.ws ? 0 : (xs < g.W - g.shift ? 1 : 2);
    }
    region = ry * 3 + rx;
    int y = ys + g.shift, x = xs + g.shift;
    if (y >= g.H) y -= g.H;
    if (x >= g.W) x -= g.W;
    return ((long long)b * g.H + y) * g.W + x;
  }
  __device__ __forceinline__ void advance(const AttnGeom& g) {
    wx += swx;
    if (wx >= g.nwx) { wx -= g.nwx; ++wy; }
    wy += swy;
    if (wy >= g.nwy) { wy -= g.nwy; ++b; }
    b += sb;
  }
};

// Softmax inputs of the two kernels below.  The bias table lives in shared memory at a pitch of 68 floats (16-byte rows,
// conflict-free LDS.128 for consecutive rows), pre-multiplied by log2(e) so the exponentials are bare ex2; the columns
// of keys that do not exist (N .. 63) hold -inf, the rows of tokens that do not exist hold 0, so the per-key loop needs
// neither a bounds test nor a branch (r5: the branchy scalar version compiled to ~47 instructions and two dependent
// shared-memory round trips per key — half of the forward kernel's time per window pair, scripts/attn_one.py).
constexpr int kBiasPitch = 68;
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fill_bias_table(uint32_t aBias, const float* __restrict__ bias_head, int N, int tid,
                                                int nthreads) {
  for (int i = tid; i < kMaxN * kBiasPitch; i += nthreads) {
    const int tr = i / kBiasPitch, tc = i - tr * kBiasPitch;
    float v = 0.f;
    if (tc >= N) v = -INFINITY;
    else if (tr < N) v = bias_head[tr * N + tc] * kLog2e;
    sts_f32(aBias + i * 4, v);
  }
}
// logits of keys [key0, key0 + 4 NV) of one row: e[j] = s[j] * scale2 + bias2[key0 + j] (+ shift mask), returns their max
template <int NV>
__device__ __forceinline__ float logits_row(const uint32_t* sraw, float* e, uint32_t brow, uint32_t rrow, float scale2,
                                            int region, bool shifted) {
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const uint4 b4 = lds128(brow + c * 16);
    const float bb[4] = {__uint_as_float(b4.x), __uint_as_float(b4.y), __uint_as_float(b4.z), __uint_as_float(b4.w)};
    float mk[4] = {0.f, 0.f, 0.f, 0.f};
    if (shifted) {   // uniform: the -100 of timm's shifted-window mask, in log2 units
      const uint4 r4 = lds128(rrow + c * 16);
      const int rg[4] = {(int)r4.x, (int)r4.y, (int)r4.z, (int)r4.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) mk[q] = rg[q] != region ? -100.f * kLog2e : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float sc = fmaf(__uint_as_float(sraw[c * 4 + q]), scale2, bb[q]) + mk[q];
      e[c * 4 + q] = sc;
      mx = fmaxf(mx, sc);
    }
  }
  return fmaxf(mx, -1e30f);   // a quarter made of absent keys only: keeps exp2(-inf - mx) = 0 instead of NaN
}

// Forward, second layout: the same UMMA chains, but TWO threads per row (256 threads, warps 0-3 own key columns 0-31 of
// the row's window, warps 4-7 columns 32-63) and two CTAs per SM, so sixteen warps hide each other's latencies where
// the one-thread-per-row kernel above was issue-bound (2300 instructions per row and iteration on 8 warps per SM).
// The two halves of a row merge their (max, sum) once through shared memory (online-softmax merge).  Gather roles:
// half 0 stages q (hi | lo) and v, half 1 stages k (hi | hi, lo into the v rows); O columns are split 16 / 16.
constexpr int kTc2Threads = 256;
constexpr int kTc2Smem = 3 * 16384 + 32768 + kMaxN * kBiasPitch * 4 + 128 * 2 * 8 + 128 * 4 + 64 + 1024;

__device__ __forceinline__ uint4 ldg_na(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
template <bool PROF>
__global__ void __launch_bounds__(kTc2Threads + 32, 2)
window_attn_fwd_tc2_kernel(AttnGeom g, int groups, const __nv_bfloat16* __restrict__ qkv,
                           const float* __restrict__ logit_scale, const float* __restrict__ bias,
                           __nv_bfloat16* __restrict__ out, long long* __restrict__ prof, int dbg) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t aQ = smem_u32(smem), aK = aQ + 16384, aV = aK + 16384, aP = aV + 16384;
  const uint32_t aBias = aP + 32768;               // float [64][kBiasPitch]: fill_bias_table
  const uint32_t aX = aBias + kMaxN * kBiasPitch * 4;   // float2 [128 rows][2 halves]: (max, sum)
  const uint32_t aReg = aX + 128 * 2 * 8;          // int [128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (aReg - aQ) + 128 * 4);   // [0] S ready, [1] O ready (UMMA commits),
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 4);                  // [2] operands staged, [3] P staged (workers)

  const int N = g.ws * g.ws;
  const int head = blockIdx.x % g.heads;
  const int grp = blockIdx.x / g.heads;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int h = warp >> 2;                      // column half / gather role
  const int r = (warp & 3) * 32 + (tid & 31);   // stacked row = TMEM lane
  const int prob = r >> 6;
  const int t = r & 63;
  const float scale2 = __expf(fminf(logit_scale[head], 4.6051702f)) * kLog2e;
  const int total_windows = g.B * g.nwy * g.nwx;
  const int total_pairs = (total_windows + 1) / 2;

  for (int i = tid; i < (3 * 16384 + 32768) / 16; i += kTc2Threads + 32) sts128(aQ + i * 16, make_uint4(0, 0, 0, 0));
  fill_bias_table(aBias, bias + (long long)head * N * N, N, tid, kTc2Threads + 32);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], kTc2Threads / 32);   // one arrival per worker warp
    mbar_init(&bar[3], kTc2Threads / 32);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot;      // S: columns [0, 128)
  const uint32_t tmem_o = tmem_s + 128;    // O: columns [128, 192)
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 32, false, true);   // N = the 32 head dims: the k_lo half of the V rows is not fetched

  // ---- warp 8: the UMMA issuer.  r5: issued from a worker warp, the two UMMA chains blocked that warp for their whole
  // execution (~1300 clocks per pair: the shared-memory operand fetch of the 14 instructions) and every other warp then
  // waited for it at the CTA barriers; here the workers only arrive on mbarriers and go on.
  if (warp == kTc2Threads / 32) {
    if (elect_one()) {
      uint32_t ph = 0;
      for (int pair = grp; pair < total_pairs; pair += groups, ph ^= 1) {
        mbar_wait(&bar[2], ph);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_s, make_smem_desc_sw128(aQ + k * 32, 16, 1024), make_smem_desc_sw128(aK + k * 32, 16, 1024),
                    idesc_s, k);
#pragma unroll
        for (int k = 0; k < 2; ++k)
          umma_bf16(tmem_s, make_smem_desc_sw128(aQ + k * 32, 16, 1024),
                    make_smem_desc_sw128(aV + (k + 2) * 32, 16, 1024), idesc_s, 1u);
        umma_commit(&bar[0]);
        mbar_wait(&bar[3], ph);
        tc_fence_after();
        // O = [P_A; 0] . V_A + [0; P_B] . V_B
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_o, make_smem_desc_sw128(aP + kb * 16384 + k * 32, 16, 1024),
                      make_smem_desc_sw128(aV + kb * 8192 + k * 2048, 8192, 1024), idesc_o, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&bar[1]);
      }
      // the last commit has completed before any worker leaves its loop; TMEM is released by warp 0 below
    }
    return;
  }

  // The next pair's rows are fetched while P . V runs.  Each row's home position is worked out by the thread that owns the
  // row (softmax, output store); the LOADS are dealt the other way round: lane l moves 16-byte chunk (l & 3) of the four
  // rows R0 + rmap(l >> 2) + 8 i of its warp's 32-row block, so one warp instruction covers 8 rows x 64 contiguous bytes
  // instead of 32 rows x 16 bytes (r5: the row-per-lane loads kept the LSU busy for ~20 clocks per instruction, 1000-2000
  // clocks per pair and CTA).  rmap interleaves rows {0, 4, 1, 5, ...} so the eight lanes of a 16-byte shared-memory
  // store phase hit eight different swizzled columns.
  const int lane = tid & 31;
  const int lc = lane & 3;
  const int lrow8 = ((lane >> 2) & 1) * 4 + (lane >> 3);
  const int crow0 = (warp & 3) * 32 + lrow8;     // stacked row of this lane's chunk i: crow0 + 8 i
  uint4 xr[4], vr[4];
  long long nrow = 0;
  int nregion = 0;
  bool nvalid = false;
  WindowWalker ww;
  ww.init(g, grp * 2 + prob, groups * 2, t);
  auto fetch = [&]() {
    nvalid = t < N && ww.in_range(g);
    nrow = 0;
    nregion = 0;
    if (nvalid) nrow = ww.row(g, nregion);
    ww.advance(g);
    const int mine = nvalid ? static_cast<int>(nrow) : -1;
    const int cstep = g.C / 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gr = __shfl_sync(0xffffffffu, mine, lrow8 + 8 * i);
      xr[i] = make_uint4(0, 0, 0, 0);
      vr[i] = make_uint4(0, 0, 0, 0);
      if (gr >= 0) {
        const uint4* base = reinterpret_cast<const uint4*>(qkv + (long long)gr * 3 * g.C + head * kHd) + lc;
        if (PROF && dbg == 2) continue;   // timing experiment: no loads at all (results are garbage)
        if (PROF && dbg == 1) {           // timing experiment: loads that do not allocate in L1
          xr[i] = ldg_na(base + h * cstep);
          if (h == 0) vr[i] = ldg_na(base + 2 * cstep);
          continue;
        }
        xr[i] = __ldg(base + h * cstep);            // q (half 0) or k (half 1)
        if (h == 0) vr[i] = __ldg(base + 2 * cstep);
      }
    }
  };
  fetch();

  // TOK_ATTN_PROFILE=1 (bring-up aid): clock64 phase sums of CTA 0, threads 0 (UMMA issuer) and 128, read back by
  // tok_debug_attn_profile()
  const bool profiling = PROF && prof != nullptr && blockIdx.x == 0 && (tid == 0 || tid == 128);
  long long pt[PROF ? 14 : 1] = {0};
  long long tp = profiling ? clock64() : 0;
#define TOK_APROF(i)                   \
  if (PROF && profiling) {                     \
    const long long now = clock64();   \
    pt[i] += now - tp;                 \
    tp = now;                          \
  }
  uint32_t phase = 0;
  for (int pair = grp; pair < total_pairs; pair += groups, phase ^= 1) {
    const bool valid = nvalid;
    const long long my_row = nrow;
    const int region = nregion;
    {
      // cosine normalisation: the four lanes holding a row's chunks add their partial sums of squares (absent rows were
      // fetched as zeros and stage zeros)
      float inv_norm[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t wv[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w};
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) ss = fmaf(bf16_lo(wv[e]), bf16_lo(wv[e]), fmaf(bf16_hi(wv[e]), bf16_hi(wv[e]), ss));
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        inv_norm[i] = rsqrtf(fmaxf(ss, 1e-24f));   // 1 / max(|x|, 1e-12) (F.normalize), one MUFU
      }
      if (PROF && profiling && inv_norm[0] == 123.f) pt[0] = 1;   // keeps the stamp below after the first use of the loads
      TOK_APROF(13)
      // split-bf16 cosine operands (see window_attn_fwd_tc_kernel): sQ = [q_hi | q_lo], sK = [k_hi | k_hi], sV = [v | k_lo]
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cr = crow0 + 8 * i;
        const uint32_t wv[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w};
        uint32_t o[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x0 = bf16_lo(wv[e]) * inv_norm[i], x1 = bf16_hi(wv[e]) * inv_norm[i];
          o[e] = pack_bf16x2(x0, x1);
          lo[e] = pack_bf16x2(x0 - bf16_lo(o[e]), x1 - bf16_hi(o[e]));
        }
        if (h == 0) {
          sts128(aQ + sw128_off(cr, lc), make_uint4(o[0], o[1], o[2], o[3]));
          sts128(aQ + sw128_off(cr, lc + 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
          sts128(aV + sw128_off(cr, lc), vr[i]);
        } else {
          sts128(aK + sw128_off(cr, lc), make_uint4(o[0], o[1], o[2], o[3]));
          sts128(aK + sw128_off(cr, lc + 4), make_uint4(o[0], o[1], o[2], o[3]));
          sts128(aV + sw128_off(cr, lc + 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
        }
      }
    }
    if (h == 0) sts_f32(aReg + r * 4, __int_as_float(region));
    TOK_APROF(0)
    fence_proxy_async_smem();
    tc_fence_before();   // this thread's TMEM reads of the previous pair are ordered before the issuer's next UMMAs
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar[2]);
    TOK_APROF(1)
    TOK_APROF(10)
    mbar_wait(&bar[0], phase);
    tc_fence_after();
    TOK_APROF(2)
    // ---- softmax of this half-row, merged with the other half
    float e[32];
    float mx;
    {
      uint32_t sraw[32];
      tmem_ld_32x32b_x32(tmem_s + lane_addr + prob * 64 + h * 32, sraw);
      tmem_ld_wait();
      mx = logits_row<8>(sraw, e, aBias + (t * kBiasPitch + h * 32) * 4, aReg + (prob * 64 + h * 32) * 4, scale2, region,
                         g.shift != 0);
    }
    float lsum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      e[j] = ex2_approx(e[j] - mx);
      lsum += e[j];
    }
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(aX + (r * 2 + h) * 8), "f"(mx), "f"(lsum) : "memory");
    TOK_APROF(3)
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the two halves of every row have published (max, sum)
    TOK_APROF(4)

    float factor = 0.f;
    {
      float m0, l0, m1, l1;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(m0), "=f"(l0), "=f"(m1), "=f"(l1) : "r"(aX + r * 16));
      const float m = fmaxf(m0, m1);
      const float f0 = ex2_approx(m0 - m), f1 = ex2_approx(m1 - m);
      const float l = l0 * f0 + l1 * f1;
      if (valid && l > 0.f) factor = (h == 0 ? f0 : f1) / l;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t o4[4];
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) o4[q2] = pack_bf16x2(e[c * 8 + 2 * q2] * factor, e[c * 8 + 2 * q2 + 1] * factor);
      sts128(aP + prob * 16384 + sw128_off(r, 4 * h + c), make_uint4(o4[0], o4[1], o4[2], o4[3]));
    }
    TOK_APROF(5)
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar[3]);
    TOK_APROF(6)
    TOK_APROF(11)
    fetch();
    TOK_APROF(12)
    mbar_wait(&bar[1], phase);
    tc_fence_after();
    TOK_APROF(7)
    {
      uint32_t acc[16];
      tmem_ld_32x32b_x16(tmem_o + lane_addr + h * 16, acc);
      tmem_ld_wait();
      if (valid) {
        uint4* op = reinterpret_cast<uint4*>(out + my_row * g.C + head * kHd + h * 16);
#pragma unroll
        for (int c = 0; c < 2; ++c)
          op[c] = make_uint4(pack_bf16x2(__uint_as_float(acc[c * 8]), __uint_as_float(acc[c * 8 + 1])),
                             pack_bf16x2(__uint_as_float(acc[c * 8 + 2]), __uint_as_float(acc[c * 8 + 3])),
                             pack_bf16x2(__uint_as_float(acc[c * 8 + 4]), __uint_as_float(acc[c * 8 + 5])),
                             pack_bf16x2(__uint_as_float(acc[c * 8 + 6]), __uint_as_float(acc[c * 8 + 7])));
      }
    }
    TOK_APROF(8)
    // no CTA barrier here: this thread has seen O ready, so both UMMA chains of the pair are complete and the operand
    // tiles may be restaged; the S / O accumulators are rewritten only after every worker warp has arrived again
  }
#undef TOK_APROF
  if (PROF && profiling) {
#pragma unroll
    for (int i = 0; i < (PROF ? 14 : 1); ++i) prof[(tid >> 7) * 14 + i] = pt[i];
  }
  tc_fence_before();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (warp == 0) tmem_dealloc(tmem_s, 256);
}

// Backward.  One CTA per (head, group); the CTA loops over the windows of its group so that the bias / logit-scale
// gradients accumulate on chip and are flushed once.  Pass 1 (thread = query row i): softmax statistics, O_i,
// delta_i = dO_i . O_i, dq_i.  Pass 2 (thread = key row j): dk_j, dv_j, dbias[:, j].  Row vectors live in registers,
// the other operand is read from shared memory with 16-byte broadcasts.
constexpr int kAttnBwdSmem = (4 * kMaxN * kPitch + kMaxN * kMaxN + 5 * kMaxN) * 4 + kMaxN * 4 + kMaxN * 8 + 64 * 4 + 64;

__global__ void __launch_bounds__(64)
window_attn_bwd_kernel(AttnGeom g, int groups, const __nv_bfloat16* __restrict__ qkv,
                       const float* __restrict__ logit_scale, const float* __restrict__ bias,
                       const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqkv,
                       float* __restrict__ dbias, float* __restrict__ dlogit_scale) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ __align__(16) float dyn[];
  float (*sq)[kPitch] = reinterpret_cast<float (*)[kPitch]>(dyn);
  float (*sk)[kPitch] = sq + kMaxN;
  float (*sv)[kPitch] = sk + kMaxN;
  float (*sdo)[kPitch] = sv + kMaxN;
  float (*sdb)[kMaxN] = reinterpret_cast<float (*)[kMaxN]>(sdo + kMaxN);  // dbias[i][j] of this CTA, owner = thread j
  float* sm = reinterpret_cast<float*>(sdb + kMaxN);
  float* sl = sm + kMaxN;
  float* sdelta = sl + kMaxN;
  float* sqn = sdelta + kMaxN;
  float* skn = sqn + kMaxN;
  float* s_red = skn + kMaxN;
  int* sreg = reinterpret_cast<int*>(s_red + 64);
  long long* srow = reinterpret_cast<long long*>(sreg + kMaxN);
  const int N = g.ws * g.ws;
  const int head = blockIdx.x % g.heads;
  const int grp = blockIdx.x / g.heads;
  const int t = threadIdx.x;
  const float ls = logit_scale[head];
  const float scale = __expf(fminf(ls, 4.6051702f));
  const int total_windows = g.B * g.nwy * g.nwx;
  for (int i = 0; i < kMaxN; ++i) sdb[i][t] = 0.f;
  float dls = 0.f;
  for (int w = grp; w < total_windows; w += groups) {
    const int wx = w % g.nwx;
    const int wy = (w / g.nwx) % g.nwy;
    const int b = w / (g.nwx * g.nwy);
    __syncthreads();  // previous window fully consumed
    if (t < N) {
      int region;
      const long long row = token_row(g, b, wy, wx, t, region);
      srow[t] = row;
      sreg[t] = region;
      const __nv_bfloat16* base = qkv + row * 3 * g.C + head * kHd;
      const __nv_bfloat16* dob = dout + row * g.C + head * kHd;
      float qq = 0.f, kk = 0.f;
#pragma unroll
      for (int e = 0; e < kHd; ++e) {
        const float qv = __bfloat162float(base[e]), kv = __bfloat162float(base[g.C + e]);
        sq[t][e] = qv;
        sk[t][e] = kv;
        qq = fmaf(qv, qv, qq);
        kk = fmaf(kv, kv, kk);
        sv[t][e] = __bfloat162float(base[2 * g.C + e]);
        sdo[t][e] = __bfloat162float(dob[e]);
      }
      const float qi = 1.f / fmaxf(sqrtf(qq), 1e-12f), ki = 1.f / fmaxf(sqrtf(kk), 1e-12f);
      sqn[t] = qi;
      skn[t] = ki;
#pragma unroll
      for (int e = 0; e < kHd; ++e) {
        sq[t][e] *= qi;  // q_hat
        sk[t][e] *= ki;  // k_hat
      }
    }
    __syncthreads();
    // ---- pass 1: row i = t
    if (t < N) {
      float qr[kHd], dor[kHd];
#pragma unroll
      for (int e = 0; e < kHd; ++e) {
        qr[e] = sq[t][e];
        dor[e] = sdo[t][e];
      }
      const float* brow = bias + ((long long)head * N + t) * N;
      const int my_reg = sreg[t];
      float p[kMaxN];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < kMaxN; ++j) {
        if (j < N) {
          const float s = dot32(qr, sk[j]) * scale + brow[j] + (sreg[j] != my_reg ? -100.f : 0.f);
          p[j] = s;
          mx = fmaxf(mx, s);
        }
      }
      float l = 0.f;
      float o[kHd];
#pragma unroll
      for (int e = 0; e < kHd; ++e) o[e] = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxN; ++j) {
        if (j < N) {
          p[j] = __expf(p[j] - mx);
          l += p[j];
          axpy32(o, p[j], sv[j]);
        }
      }
      const float inv = 1.f / l;
      float delta = 0.f;  // sum_j P_ij dP_ij = dO_i . O_i
#pragma unroll
      for (int e = 0; e < kHd; ++e) delta = fmaf(dor[e], o[e] * inv, delta);
      float dq[kHd];
#pragma unroll
      for (int e = 0; e < kHd; ++e) dq[e] = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxN; ++j) {
        if (j < N) {
          const float dp = dot32(dor, sv[j]);
          axpy32(dq, p[j] * inv * (dp - delta) * scale, sk[j]);
        }
      }
      sm[t] = mx;
      sl[t] = inv;
      sdelta[t] = delta;
      // through the normalisation: dq = (dq_hat - q_hat (q_hat . dq_hat)) / |q|
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < kHd; ++e) dot = fmaf(qr[e], dq[e], dot);
      __nv_bfloat16* dst = dqkv + srow[t] * 3 * g.C + head * kHd;
#pragma unroll
      for (int e = 0; e < kHd; ++e) dst[e] = __float2bfloat16((dq[e] - qr[e] * dot) * sqn[t]);
    }
    __syncthreads();
    // ---- pass 2: key column j = t
    if (t < N) {
      float kr[kHd], vr[kHd], dk[kHd], dv[kHd];
#pragma unroll
      for (int e = 0; e < kHd; ++e) {
        kr[e] = sk[t][e];
        vr[e] = sv[t][e];
        dk[e] = dv[e] = 0.f;
      }
      const int my_reg = sreg[t];
#pragma unroll 2
      for (int i = 0; i < N; ++i) {
        const float c = dot32(kr, sq[i]);
        const float dp = dot32(vr, sdo[i]);
        const float s = c * scale + bias[((long long)head * N + i) * N + t] + (sreg[i] != my_reg ? -100.f : 0.f);
        const float p = __expf(s - sm[i]) * sl[i];
        const float ds = p * (dp - sdelta[i]);
        sdb[i][t] += ds;
        dls = fmaf(ds, c, dls);
        axpy32(dk, ds * scale, sq[i]);
        axpy32(dv, p, sdo[i]);
      }
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < kHd; ++e) dot = fmaf(kr[e], dk[e], dot);
      __nv_bfloat16* dst = dqkv + srow[t] * 3 * g.C + head * kHd;
#pragma unroll
      for (int e = 0; e < kHd; ++e) {
        dst[g.C + e] = __float2bfloat16((dk[e] - kr[e] * dot) * skn[t]);
        dst[2 * g.C + e] = __float2bfloat16(dv[e]);
      }
    }
  }
  // flush: d bias[head][i][t], d logit_scale[head] (zero when the clamp is active)
  if (t < N) {
    for (int i = 0; i < N; ++i) atomicAdd(dbias + ((long long)head * N + i) * N + t, sdb[i][t]);
  }
  s_red[t] = (t < N && ls < 4.6051702f) ? dls * scale : 0.f;
  __syncthreads();
  if (t == 0) {
    float a = 0.f;
    for (int i = 0; i < 64; ++i) a += s_red[i];
    atomicAdd(dlogit_scale + head, a);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// tcgen05 backward: all five contractions of the attention backward on the tensor cores, in ONE kernel.
// A CTA (512 threads) owns one head and walks pairs of windows stacked along M exactly like the forward kernel.
// Thread (r, h): r = stacked token row 0..127 (= TMEM lane), h = warp >> 2 = which quarter of the row's 64 key columns
// it owns in the softmax phase / which tensor it gathers (q, k, v, dO) / which gradient it finishes (dq, dk, dv).
//   round 1   S  [128 x 128] = Qhat . Khat^T          dP [128 x 128] = dO . V^T              (K = 32, K-major operands)
//   registers P = softmax(S * scale + bias + mask), delta = sum_j P dP, dS = P o (dP - delta); the four threads of a
//             row merge their (max, sum, sum P~ dP) partials through shared memory once (online-softmax merge);
//             dbias and dlogit_scale accumulate in registers across all windows of the CTA
//   round 2   dV [128 x 32] = P^T . dO    dKhat = dS^T . Qhat   (A = the bf16 P / dS tiles read MN-major, K = 128 rows)
//             dQhat [128 x 32] = dS . Khat                      (A K-major block-diagonal halves, as P . V in forward)
//   epilogue  through the cosine normalisation: dq = scale * (dQhat - qhat (qhat . dQhat)) / |q|, same for k.
// The next pair's rows are prefetched into registers while round 2 runs.
constexpr int kBwThreads = 512;
constexpr int kBwTiles = 8 * 16384;  // sQ sK sV sDO, sP[2], sDS[2]
constexpr int kBwSmem = kBwTiles + kMaxN * kBiasPitch * 4 + 128 * 4 * 16 + 128 * 4 + 64 + 32 * 256 * 4 + 1024;

// The UMMA issuer is the elected lane of warp 12 (dO role: no epilogue work), not a 17th warp: a fifth warp on one SM
// sub-partition caps the kernel at 96 registers per thread (16384 / (5 x 32); 112 and 120 do not launch) and the spills
// cost more than the dedicated warp gains (r5: 1078 vs 905 us for stage 1 at bs256).
__global__ void __launch_bounds__(kBwThreads, 1)
window_attn_bwd_tc_kernel(AttnGeom g, int groups, const __nv_bfloat16* __restrict__ qkv,
                          const float* __restrict__ logit_scale, const float* __restrict__ bias,
                          const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dqkv,
                          float* __restrict__ dbias, float* __restrict__ dlogit_scale, float* __restrict__ dcolsum) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t aQ = smem_u32(smem), aK = aQ + 16384, aV = aK + 16384, aDO = aV + 16384;
  const uint32_t aP = aDO + 16384, aDS = aP + 32768;
  const uint32_t aBias = aDS + 32768;                 // float [64][kBiasPitch]: fill_bias_table
  const uint32_t aX = aBias + kMaxN * kBiasPitch * 4;  // float4 [128 rows][4 quarters]: (max, sum, sum p~ dP, -)
  const uint32_t aReg = aX + 128 * 4 * 16;            // int [128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (aReg - aQ) + 128 * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 4);   // bar: [0] round 1 done, [1] round 2 done, [2] operands staged, [3] P / dS staged
  // column sums of dq / dv over every token this CTA sees (= the q_bias / v_bias gradients): float [32][256],
  // slot = (dv ? 128 : 0) + row, owned by one thread each
  const uint32_t aCol = aReg + 128 * 4 + 64;

  const int N = g.ws * g.ws;
  const int head = blockIdx.x % g.heads;
  const int grp = blockIdx.x / g.heads;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int h = warp >> 2;                      // column quarter / role
  const int r = (warp & 3) * 32 + (tid & 31);   // stacked row = TMEM lane
  const int prob = r >> 6;
  const int t = r & 63;
  const float ls = logit_scale[head];
  const float scale = __expf(fminf(ls, 4.6051702f));
  const float scale2 = scale * kLog2e;
  const int total_windows = g.B * g.nwy * g.nwx;
  const int total_pairs = (total_windows + 1) / 2;

  for (int i = tid; i < kBwTiles / 16; i += kBwThreads) sts128(aQ + i * 16, make_uint4(0, 0, 0, 0));
  for (int i = tid; i < 32 * 256 / 4; i += kBwThreads) sts128(aCol + i * 16, make_uint4(0, 0, 0, 0));
  fill_bias_table(aBias, bias + (long long)head * N * N, N, tid, kBwThreads);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], kBwThreads / 32);   // one arrival per worker warp
    mbar_init(&bar[3], kBwThreads / 32);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;  // S [0,128) dP [128,256) dQ [256,320) dK [320,384) dV [384,448)
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  constexpr uint32_t idesc_kk = make_idesc_bf16(128, 128, false, false);
  constexpr uint32_t idesc_mm = make_idesc_bf16(128, 32, true, true);    // N = 32: only the data half of the 128-byte B rows is fetched
  constexpr uint32_t idesc_km = make_idesc_bf16(128, 32, false, true);
  const uint32_t my_tile = aQ + h * 16384;  // tile this thread fills in the gather (q, k, v, dO by role)


  float db[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) db[j] = 0.f;
  float dls = 0.f;

  // ---- rows of a pair: token position, validity, mask region by the thread that owns the row; the 64 bytes of this
  // role's tensor are LOADED chunk-wise (lane l: 16-byte chunk l & 3 of the rows R0 + rmap(l >> 2) + 8 i — see the forward
  // kernel: 8 rows x 64 contiguous bytes per warp instruction)
  const int lane = tid & 31;
  const int lc = lane & 3;
  const int lrow8 = ((lane >> 2) & 1) * 4 + (lane >> 3);
  const int crow0 = (warp & 3) * 32 + lrow8;
  // the lanes that hold the chunks of THIS lane's own row (for the norm of roles 0, 1): row lane = rmap(j) + 8 i
  const int own_i = lane >> 3;
  const int own_src = (((lane & 7) >> 2) + ((lane & 3) << 1)) << 2;
  uint4 raw[4];
  long long nrow = 0;
  int nregion = 0;
  bool nvalid = false;
  WindowWalker ww;
  ww.init(g, grp * 2 + prob, groups * 2, t);
  auto fetch = [&]() {
    nvalid = t < N && ww.in_range(g);
    nrow = 0;
    nregion = 0;
    if (nvalid) nrow = ww.row(g, nregion);
    ww.advance(g);
    const int mine = nvalid ? static_cast<int>(nrow) : -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gr = __shfl_sync(0xffffffffu, mine, lrow8 + 8 * i);
      raw[i] = make_uint4(0, 0, 0, 0);
      if (gr >= 0) {
        const uint4* src = h < 3 ? reinterpret_cast<const uint4*>(qkv + (long long)gr * 3 * g.C + h * g.C + head * kHd)
                                 : reinterpret_cast<const uint4*>(dout + (long long)gr * g.C + head * kHd);
        raw[i] = __ldg(src + lc);
      }
    }
  };
  fetch();

  uint32_t phase = 0;
  for (int pair = grp; pair < total_pairs; pair += groups, phase ^= 1) {
    const long long row = nrow;
    const int region = nregion;
    const bool valid = nvalid;
    float inv_norm = 0.f;  // 1 / |q| or 1 / |k| of this thread's own row (roles 0, 1)
    // ---- stage this role's operand rows (q and k normalised; absent rows were fetched as zeros)
    if (h < 2) {
      float inv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t wv[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) ss = fmaf(bf16_lo(wv[e]), bf16_lo(wv[e]), fmaf(bf16_hi(wv[e]), bf16_hi(wv[e]), ss));
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        inv[i] = rsqrtf(fmaxf(ss, 1e-24f));   // 1 / max(|x|, 1e-12) (F.normalize), one MUFU
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v = __shfl_sync(0xffffffffu, inv[i], own_src);
        if (own_i == i) inv_norm = v;
      }
      // split-bf16 operands exactly as in the forward kernel: sQ = [q_hi | q_lo], sK = [k_hi | k_hi], sV = [v | k_lo]
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cr = crow0 + 8 * i;
        const uint32_t wv[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
        uint32_t o[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x0 = bf16_lo(wv[e]) * inv[i], x1 = bf16_hi(wv[e]) * inv[i];
          o[e] = pack_bf16x2(x0, x1);
          lo[e] = pack_bf16x2(x0 - bf16_lo(o[e]), x1 - bf16_hi(o[e]));
        }
        sts128(my_tile + sw128_off(cr, lc), make_uint4(o[0], o[1], o[2], o[3]));
        if (h == 0) {
          sts128(aQ + sw128_off(cr, lc + 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
        } else {
          sts128(aK + sw128_off(cr, lc + 4), make_uint4(o[0], o[1], o[2], o[3]));
          sts128(aV + sw128_off(cr, lc + 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
        }
      }
      if (h == 0) sts_f32(aReg + r * 4, __int_as_float(region));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) sts128(my_tile + sw128_off(crow0 + 8 * i, lc), raw[i]);
    }
    fence_proxy_async_smem();
    tc_fence_before();   // orders this thread's TMEM reads of the previous pair before the issuer's next UMMAs
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar[2]);
    if (warp == 12 && elect_one()) {   // this warp waits for round 1 anyway: the blocking UMMA issue costs it nothing
      mbar_wait(&bar[2], phase);
      tc_fence_after();
      // round 1: S = Qhat Khat^T, dP = dO V^T
#pragma unroll
      for (int k = 0; k < 4; ++k)   // [q_hi | q_lo] . [k_hi | k_hi]
        umma_bf16(tmem, make_smem_desc_sw128(aQ + k * 32, 16, 1024), make_smem_desc_sw128(aK + k * 32, 16, 1024),
                  idesc_kk, k);
#pragma unroll
      for (int k = 0; k < 2; ++k)   // q_hi . k_lo
        umma_bf16(tmem, make_smem_desc_sw128(aQ + k * 32, 16, 1024),
                  make_smem_desc_sw128(aV + (k + 2) * 32, 16, 1024), idesc_kk, 1u);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        umma_bf16(tmem + 128, make_smem_desc_sw128(aDO + k * 32, 16, 1024),
                  make_smem_desc_sw128(aV + k * 32, 16, 1024), idesc_kk, k);
      umma_commit(&bar[0]);
    }
    __syncwarp();
    mbar_wait(&bar[0], phase);
    tc_fence_after();
    // ---- softmax statistics of this quarter-row, merged across the four quarters
    uint32_t sraw[16], graw[16];
    tmem_ld_32x32b_x16(tmem + lane_addr + prob * 64 + h * 16, sraw);
    tmem_ld_32x32b_x16(tmem + 128 + lane_addr + prob * 64 + h * 16, graw);
    tmem_ld_wait();
    float e[16];
    const float mx = logits_row<4>(sraw, e, aBias + (t * kBiasPitch + h * 16) * 4, aReg + (prob * 64 + h * 16) * 4, scale2,
                                   region, g.shift != 0);
    float lsum = 0.f, dsum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      e[j] = ex2_approx(e[j] - mx);   // 0 for the padded keys (-inf in the bias table)
      lsum += e[j];
      dsum = fmaf(e[j], __uint_as_float(graw[j]), dsum);
    }
    sts128(aX + (r * 4 + h) * 16, make_uint4(__float_as_uint(mx), __float_as_uint(lsum), __float_as_uint(dsum), 0u));
    asm volatile("bar.sync 1, 512;" ::: "memory");   // workers only: the four quarters of every row have published
    float factor = 0.f, delta = 0.f;
    {
      uint4 x[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) x[q4] = lds128(aX + (r * 4 + q4) * 16);
      float m = -INFINITY;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) m = fmaxf(m, __uint_as_float(x[q4].x));
      float l = 0.f, d = 0.f;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float f = ex2_approx(__uint_as_float(x[q4].x) - m);
        l = fmaf(__uint_as_float(x[q4].y), f, l);
        d = fmaf(__uint_as_float(x[q4].z), f, d);
      }
      if (valid && l > 0.f) {
        const float inv = 1.f / l;
        delta = d * inv;
        factor = ex2_approx(mx - m) * inv;
      }
    }
    // ---- P and dS of this quarter-row: bf16 tiles for round 2, dbias / dlogit_scale partials in registers
    {
      uint32_t pp[8], dd[8];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float p0 = e[j] * factor, p1 = e[j + 1] * factor;
        const float d0 = p0 * (__uint_as_float(graw[j]) - delta), d1 = p1 * (__uint_as_float(graw[j + 1]) - delta);
        db[j] += d0;
        db[j + 1] += d1;
        dls = fmaf(d0, __uint_as_float(sraw[j]), fmaf(d1, __uint_as_float(sraw[j + 1]), dls));
        pp[j >> 1] = pack_bf16x2(p0, p1);
        dd[j >> 1] = pack_bf16x2(d0, d1);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        sts128(aP + prob * 16384 + sw128_off(r, 2 * h + c), make_uint4(pp[4 * c], pp[4 * c + 1], pp[4 * c + 2], pp[4 * c + 3]));
        sts128(aDS + prob * 16384 + sw128_off(r, 2 * h + c), make_uint4(dd[4 * c], dd[4 * c + 1], dd[4 * c + 2], dd[4 * c + 3]));
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar[3]);
    if (warp == 12 && elect_one()) {
      mbar_wait(&bar[3], phase);
      tc_fence_after();
      // round 2
#pragma unroll
      for (int k = 0; k < 8; ++k)   // dV = P^T dO: K = the 128 stacked query rows
        umma_bf16(tmem + 384, make_smem_desc_sw128(aP + k * 2048, 16384, 1024),
                  make_smem_desc_sw128(aDO + k * 2048, 8192, 1024), idesc_mm, k);
#pragma unroll
      for (int k = 0; k < 8; ++k)   // dKhat = dS^T Qhat
        umma_bf16(tmem + 320, make_smem_desc_sw128(aDS + k * 2048, 16384, 1024),
                  make_smem_desc_sw128(aQ + k * 2048, 8192, 1024), idesc_mm, k);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {   // dQhat = dS Khat, block-diagonal halves
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem + 256, make_smem_desc_sw128(aDS + kb * 16384 + k * 32, 16, 1024),
                    make_smem_desc_sw128(aK + kb * 8192 + k * 2048, 8192, 1024), idesc_km, (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(&bar[1]);
    }
    __syncwarp();
    fetch();   // the next pair's rows travel while round 2 runs
    mbar_wait(&bar[1], phase);
    tc_fence_after();
    if (h < 3) {
      uint32_t acc[32];
      tmem_ld_32x32b_x32(tmem + 256 + h * 64 + lane_addr, acc);
      tmem_ld_wait();
      if (valid) {
        float mul = 1.f;
        if (h < 2) {
          // through x_hat = x / |x|:  dx = (dx_hat - x_hat (x_hat . dx_hat)) / |x|, with the logit scale folded in
          uint4 hat[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) hat[c] = lds128(my_tile + sw128_off(r, c));
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t wv[4] = {hat[c].x, hat[c].y, hat[c].z, hat[c].w};
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2)
              dot = fmaf(bf16_lo(wv[q2]), __uint_as_float(acc[c * 8 + 2 * q2]),
                         fmaf(bf16_hi(wv[q2]), __uint_as_float(acc[c * 8 + 2 * q2 + 1]), dot));
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t wv[4] = {hat[c].x, hat[c].y, hat[c].z, hat[c].w};
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2) {
              acc[c * 8 + 2 * q2] = __float_as_uint(fmaf(-bf16_lo(wv[q2]), dot, __uint_as_float(acc[c * 8 + 2 * q2])));
              acc[c * 8 + 2 * q2 + 1] = __float_as_uint(fmaf(-bf16_hi(wv[q2]), dot, __uint_as_float(acc[c * 8 + 2 * q2 + 1])));
            }
          }
          mul = scale * inv_norm;
        }
        if (dcolsum != nullptr && h != 1) {
          const uint32_t slot = aCol + ((h >> 1) * 128 + r) * 4;
#pragma unroll
          for (int e0 = 0; e0 < 32; e0 += 8) {   // loads of a batch back to back: the asm statements keep their order
            float cur[8];
#pragma unroll
            for (int e2 = 0; e2 < 8; ++e2) cur[e2] = lds_f32(slot + (e0 + e2) * 1024);
#pragma unroll
            for (int e2 = 0; e2 < 8; ++e2) sts_f32(slot + (e0 + e2) * 1024, fmaf(__uint_as_float(acc[e0 + e2]), mul, cur[e2]));
          }
        }
        uint4* dst = reinterpret_cast<uint4*>(dqkv + row * 3 * g.C + h * g.C + head * kHd);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_uint4(pack_bf16x2(__uint_as_float(acc[c * 8]) * mul, __uint_as_float(acc[c * 8 + 1]) * mul),
                              pack_bf16x2(__uint_as_float(acc[c * 8 + 2]) * mul, __uint_as_float(acc[c * 8 + 3]) * mul),
                              pack_bf16x2(__uint_as_float(acc[c * 8 + 4]) * mul, __uint_as_float(acc[c * 8 + 5]) * mul),
                              pack_bf16x2(__uint_as_float(acc[c * 8 + 6]) * mul, __uint_as_float(acc[c * 8 + 7]) * mul));
      }
    }
    // no CTA barrier: this thread has seen round 2 complete, so the operand tiles may be restaged; the lanes of a warp
    // restage each other's rows, hence the warp barrier after the epilogue's reads of the own row
    __syncwarp();
  }
  tc_fence_before();
  asm volatile("bar.sync 1, 512;" ::: "memory");   // workers: column-sum slots complete, TMEM reads done
  // ---- flush the on-chip accumulators: d bias[head][t][key], d logit_scale[head] (zero while the clamp is active)
  if (t < N) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int key = h * 16 + j;
      if (key < N) atomicAdd(dbias + ((long long)head * N + t) * N + key, db[j]);
    }
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) dls += __shfl_xor_sync(0xffffffffu, dls, s);
  if ((tid & 31) == 0 && ls < 4.6051702f) atomicAdd(dlogit_scale + head, dls * scale);
  if (dcolsum != nullptr && tid < 64) {   // the last loop iteration ended with a __syncthreads
    const int part = tid >> 5, e2 = tid & 31;   // part 0: dq columns, part 1: dv columns
    float tsum = 0.f;
    for (int i = 0; i < 128; ++i) tsum += lds_f32(aCol + (e2 * 256 + part * 128 + ((i + e2) & 127)) * 4);
    atomicAdd(dcolsum + part * 2 * g.C + head * kHd + e2, tsum);
  }
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace tok

using namespace tok;

// grid = SMs x resident CTAs per SM: a grid-stride kernel launched with a few more CTAs than fit runs a second,
// mostly empty wave (1184 CTAs on 888 slots cost the GELU passes a third of their bandwidth)
template <typename Kern>
static int full_wave_ctas(Kern kernel, int threads, size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
  return 148 * n;
}
template <int G, int CH>
static long long ln_fwd_ctas() {
  static const int n = full_wave_ctas(layernorm_fwd_vec_kernel<G, CH>, 256, 0);
  return n;
}
template <int G, int CH>
static long long ln_bwd_ctas() {
  static const int n = full_wave_ctas(layernorm_bwd_vec_kernel<G, CH>, 128, 4 * 3 * 8 * CH * G * sizeof(float));
  return n;
}

// ------------------------------------------------------------------------------------------------------------------
// Swin-V2 continuous position bias (timm WindowAttention: cpb_mlp = Linear(2, 512) -> ReLU -> Linear(512, heads, no bias) on
// the log-spaced relative coordinate table, gathered by relative_position_index, 16 * sigmoid): one small kernel chain
// instead of torch's sgemm + index + sort + reduce launches (r1: ~3 % of the Swin-T step).
//   T = (2 ws - 1)^2 table entries, N = ws^2 tokens, e(i, j) = (yi - yj + ws - 1) * (2 ws - 1) + (xi - xj + ws - 1)
namespace tok {
namespace {
constexpr int kCpbHidden = 512;

// hidden[e][k] = relu(w1[k] . coords[e] + b1[k]);  t[e][h] = w2[h] . hidden[e].  One CTA (128 threads) per entry.
__global__ void __launch_bounds__(128)
cpb_table_kernel(const float* __restrict__ coords, const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ w2, float* __restrict__ hidden, float* __restrict__ table, int heads) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float sh[kCpbHidden];
  const int e = blockIdx.x;
  const float c0 = coords[2 * e], c1 = coords[2 * e + 1];
  for (int k = threadIdx.x; k < kCpbHidden; k += 128) {
    const float v = fmaxf(fmaf(w1[2 * k], c0, fmaf(w1[2 * k + 1], c1, b1[k])), 0.f);
    sh[k] = v;
    hidden[(long long)e * kCpbHidden + k] = v;
  }
  __syncthreads();
  // one warp per head (r5: the head loop used to run on the whole CTA with two barriers per head — 48 barriers and as many
  // dependent L2 round trips for 24 heads in a kernel whose whole job is 169 x 512 x heads FMAs)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < heads; h += 4) {
    float t = 0.f;
#pragma unroll 4
    for (int k = lane; k < kCpbHidden; k += 32) t = fmaf(w2[(long long)h * kCpbHidden + k], sh[k], t);
    t = warp_sum(t);
    if (lane == 0) table[(long long)e * heads + h] = t;
  }
}

__device__ __forceinline__ int cpb_entry(int i, int j, int ws) {
  const int yi = i / ws, xi = i - yi * ws, yj = j / ws, xj = j - yj * ws;
  return (yi - yj + ws - 1) * (2 * ws - 1) + (xi - xj + ws - 1);
}

// bias[h][i][j] = 16 * sigmoid(table[e(i, j)][h])
__global__ void __launch_bounds__(256)
cpb_gather_kernel(const float* __restrict__ table, float* __restrict__ bias, int heads, int ws) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int N = ws * ws;
  const long long total = (long long)heads * N * N;
  for (long long o = blockIdx.x * 256LL + threadIdx.x; o < total; o += gridDim.x * 256LL) {
    const int j = (int)(o % N);
    const int i = (int)((o / N) % N);
    const int h = (int)(o / ((long long)N * N));
    const float t = table[(long long)cpb_entry(i, j, ws) * heads + h];
    bias[o] = 16.f / (1.f + __expf(-t));
  }
}

// dtable[e][h] = 16 s (1 - s) * sum over the (i, j) with e(i, j) = e of dbias[h][i][j].  One CTA per entry.
__global__ void __launch_bounds__(128)
cpb_scatter_kernel(const float* __restrict__ dbias, const float* __restrict__ table, float* __restrict__ dtable,
                   int heads, int ws) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int N = ws * ws;
  const int e = blockIdx.x;
  const int dy = e / (2 * ws - 1) - (ws - 1), dx = e % (2 * ws - 1) - (ws - 1);
  // token j = (yj, xj) pairs with i = (yj + dy, xj + dx) when that is inside the window
  const int y0 = max(0, -dy), y1 = min(ws, ws - dy), x0 = max(0, -dx), x1 = min(ws, ws - dx);
  const int ny = y1 - y0, nx = x1 - x0, cnt = ny > 0 && nx > 0 ? ny * nx : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < heads; h += 4) {   // one warp per head, no CTA barriers
    float acc = 0.f;
    for (int t = lane; t < cnt; t += 32) {
      const int yj = y0 + t / nx, xj = x0 + t % nx;
      const int j = yj * ws + xj, i = (yj + dy) * ws + (xj + dx);
      acc += dbias[((long long)h * N + i) * N + j];
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float s = 1.f / (1.f + __expf(-table[(long long)e * heads + h]));
      dtable[(long long)e * heads + h] = 16.f * s * (1.f - s) * acc;
    }
  }
}

// MLP backward.  One CTA (128 threads) per hidden unit k:
//   dw2[h][k] = sum_e dt[e][h] hidden[e][k];  dhid[e] = [hidden > 0] sum_h dt[e][h] w2[h][k]
//   dw1[k][0..1] = sum_e dhid[e] coords[e];   db1[k] = sum_e dhid[e]
__global__ void __launch_bounds__(128)
cpb_mlp_bwd_kernel(const float* __restrict__ dtable, const float* __restrict__ hidden, const float* __restrict__ coords,
                   const float* __restrict__ w2, float* __restrict__ dw1, float* __restrict__ db1,
                   float* __restrict__ dw2, int T, int heads) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float red[4][4];
  const int k = blockIdx.x;
  float a0 = 0.f, a1 = 0.f, ab = 0.f;
  for (int e = threadIdx.x; e < T; e += 128) {
    const float hv = hidden[(long long)e * kCpbHidden + k];
    if (hv > 0.f) {
      float d = 0.f;
      for (int h = 0; h < heads; ++h) d = fmaf(dtable[(long long)e * heads + h], w2[(long long)h * kCpbHidden + k], d);
      a0 = fmaf(d, coords[2 * e], a0);
      a1 = fmaf(d, coords[2 * e + 1], a1);
      ab += d;
    }
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  ab = warp_sum(ab);
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5][0] = a0;
    red[threadIdx.x >> 5][1] = a1;
    red[threadIdx.x >> 5][2] = ab;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    dw1[2 * k] += red[0][0] + red[1][0] + red[2][0] + red[3][0];
    dw1[2 * k + 1] += red[0][1] + red[1][1] + red[2][1] + red[3][1];
    db1[k] += red[0][2] + red[1][2] + red[2][2] + red[3][2];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < heads; h += 4) {   // one warp per head, no CTA barriers
    float t = 0.f;
    for (int e = lane; e < T; e += 32) t = fmaf(dtable[(long long)e * heads + h], hidden[(long long)e * kCpbHidden + k], t);
    t = warp_sum(t);
    if (lane == 0) dw2[(long long)h * kCpbHidden + k] += t;
  }
}
}  // namespace
}  // namespace tok

extern "C" {

int tok_layernorm_fwd(long long rows, int C, const void* x, const float* gamma, const float* beta, float eps,
                      const void* residual, const float* rowscale, int rows_per_sample, void* out, float* mean,
                      float* rstd, void* stream) {
  if (rows <= 0 || C <= 0 || C > 32 * kLnMaxPerLane) return set_error(TOK_ERR_INVALID, "layernorm_fwd: 1 <= C <= 1024");
  if (rowscale && rows_per_sample <= 0) return set_error(TOK_ERR_INVALID, "layernorm_fwd: rows_per_sample");
  int G, CH;
  if (ln_vec_shape(C, &G, &CH)) {
    const long long rows_per_cta = 8LL * (32 / G);
    const long long want = (rows + rows_per_cta - 1) / rows_per_cta;
#define TOK_LN_FWD_V(GG, CC)                                                                                      \
  (void)launch_pdl(layernorm_fwd_vec_kernel<GG, CC>, dim3((unsigned)(want < ln_fwd_ctas<GG, CC>() ? want : ln_fwd_ctas<GG, CC>())), dim3(256), 0, \
                                     (cudaStream_t)stream,                                                       \
      rows, (const __nv_bfloat16*)x, gamma, beta, eps, (const __nv_bfloat16*)residual, rowscale,                   \
      rows_per_sample > 0 ? rows_per_sample : 1, (__nv_bfloat16*)out, mean, rstd)
#define TOK_LN_FWD_G(CC)                                                                                          \
  do {                                                                                                             \
    if (G == 4) TOK_LN_FWD_V(4, CC);                                                                               \
    else if (G == 8) TOK_LN_FWD_V(8, CC);                                                                          \
    else if (G == 16) TOK_LN_FWD_V(16, CC);                                                                        \
    else TOK_LN_FWD_V(32, CC);                                                                                     \
  } while (0)
    if (CH == 3) TOK_LN_FWD_G(3);
    else if (CH == 4) TOK_LN_FWD_G(4);
    else if (CH == 2) TOK_LN_FWD_G(2);
    else TOK_LN_FWD_G(1);
#undef TOK_LN_FWD_G
#undef TOK_LN_FWD_V
    TOK_CHECK_LAUNCH("layernorm_fwd_vec");
    return TOK_OK;
  }
  const long long threads = rows * 32;
  const unsigned grid = (unsigned)((threads + 127) / 128);
#define TOK_LN_FWD(P)                                                                                              \
  (void)launch_pdl(layernorm_fwd_kernel<P>, dim3(grid), dim3(128), 0, (cudaStream_t)stream,                                                   \
      rows, C, (const __nv_bfloat16*)x, gamma, beta, eps, (const __nv_bfloat16*)residual, rowscale,                \
      rows_per_sample > 0 ? rows_per_sample : 1, (__nv_bfloat16*)out, mean, rstd)
  if (C <= 128) TOK_LN_FWD(4);
  else if (C <= 256) TOK_LN_FWD(8);
  else if (C <= 512) TOK_LN_FWD(16);
  else TOK_LN_FWD(32);
#undef TOK_LN_FWD
  TOK_CHECK_LAUNCH("layernorm_fwd");
  return TOK_OK;
}

int tok_layernorm_has_dxsum(int C) {
  int G, CH;
  return ln_vec_shape(C, &G, &CH) ? 1 : 0;
}

int tok_layernorm_bwd(long long rows, int C, const void* x, const float* gamma, const float* mean, const float* rstd,
                      const void* dout, const float* rowscale, int rows_per_sample, void* dx, float* dgamma,
                      float* dbeta, float* dxsum, void* stream) {
  if (rows <= 0 || C <= 0 || C > 32 * kLnMaxPerLane) return set_error(TOK_ERR_INVALID, "layernorm_bwd: 1 <= C <= 1024");
  int G, CH;
  if (ln_vec_shape(C, &G, &CH)) {
    const long long rows_per_pass = 4LL * (32 / G);
    const long long want = (rows + rows_per_pass - 1) / rows_per_pass;
#define TOK_LN_BWD_V(GG, CC)                                                                                      \
  (void)launch_pdl(layernorm_bwd_vec_kernel<GG, CC>, dim3((unsigned)(want < ln_bwd_ctas<GG, CC>() ? want : ln_bwd_ctas<GG, CC>())), dim3(128), \
                                     4 * 3 * C * sizeof(float), (cudaStream_t)stream,                            \
      rows, (const __nv_bfloat16*)x, gamma, mean, rstd, (const __nv_bfloat16*)dout, rowscale,                      \
      rows_per_sample > 0 ? rows_per_sample : 1, (__nv_bfloat16*)dx, dgamma, dbeta, dxsum)
#define TOK_LN_BWD_G(CC)                                                                                          \
  do {                                                                                                             \
    if (G == 4) TOK_LN_BWD_V(4, CC);                                                                               \
    else if (G == 8) TOK_LN_BWD_V(8, CC);                                                                          \
    else if (G == 16) TOK_LN_BWD_V(16, CC);                                                                        \
    else TOK_LN_BWD_V(32, CC);                                                                                     \
  } while (0)
    if (CH == 3) TOK_LN_BWD_G(3);
    else if (CH == 4) TOK_LN_BWD_G(4);
    else if (CH == 2) TOK_LN_BWD_G(2);
    else TOK_LN_BWD_G(1);
#undef TOK_LN_BWD_G
#undef TOK_LN_BWD_V
    TOK_CHECK_LAUNCH("layernorm_bwd_vec");
    return TOK_OK;
  }
  if (dxsum) return set_error(TOK_ERR_INVALID, "layernorm_bwd: dxsum needs a vector-path width (tok_layernorm_has_dxsum)");
  long long ctas = 148LL * 8;
  long long rpc = (rows + ctas - 1) / ctas;
  if (rpc < 4) rpc = 4;
  ctas = (rows + rpc - 1) / rpc;
#define TOK_LN_BWD(P)                                                                                              \
  (void)launch_pdl(layernorm_bwd_kernel<P>, dim3((unsigned)ctas), dim3(128), 2 * C * sizeof(float), (cudaStream_t)stream,                     \
      rows, C, (const __nv_bfloat16*)x, gamma, mean, rstd, (const __nv_bfloat16*)dout, rowscale,                   \
      rows_per_sample > 0 ? rows_per_sample : 1, (__nv_bfloat16*)dx, dgamma, dbeta, (int)rpc)
  if (C <= 128) TOK_LN_BWD(4);
  else if (C <= 256) TOK_LN_BWD(8);
  else if (C <= 512) TOK_LN_BWD(16);
  else TOK_LN_BWD(32);
#undef TOK_LN_BWD
  TOK_CHECK_LAUNCH("layernorm_bwd");
  return TOK_OK;
}

int tok_gelu_fwd(long long n, const void* x, void* y, void* stream) {
  if (n <= 0 || (n & 1)) return set_error(TOK_ERR_INVALID, "gelu: element count must be positive and even");
  if ((n & 7) == 0) {
    long long blocks = (n / 8 + 255) / 256;
    static const int wave = full_wave_ctas(gelu_fwd_vec_kernel, 256, 0);
    if (blocks > wave) blocks = wave;
    (void)launch_pdl(gelu_fwd_vec_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const uint4*)x, (uint4*)y, n / 8);
    TOK_CHECK_LAUNCH("gelu_fwd_vec");
    return TOK_OK;
  }
  long long blocks = (n / 2 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  (void)launch_pdl(gelu_fwd_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat162*)x, (__nv_bfloat162*)y, n / 2);
  TOK_CHECK_LAUNCH("gelu_fwd");
  return TOK_OK;
}
int tok_gelu_bwd(long long n, int C, const void* x, const void* dy, void* dx, float* dbias, void* stream) {
  if (n <= 0 || (n & 1)) return set_error(TOK_ERR_INVALID, "gelu: element count must be positive and even");
  if (C > 0 && (C % 128) == 0 && (n % C) == 0) {
    const long long rows = n / C;
    const int gx = C / 128;
    static const int wave = full_wave_ctas(gelu_bwd_vec_kernel, 256, 0);
    long long gy = wave / gx;   // never more CTAs than one resident wave
    if (gy < 1) gy = 1;
    if (gy > (rows + 15) / 16) gy = (rows + 15) / 16;
    (void)launch_pdl(gelu_bwd_vec_kernel, dim3(dim3((unsigned)gx, (unsigned)gy)), dim3(256), 0, (cudaStream_t)stream, 
        rows, C, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, dbias);
    TOK_CHECK_LAUNCH("gelu_bwd_vec");
    return TOK_OK;
  }
  if (dbias) return set_error(TOK_ERR_INVALID, "gelu_bwd: the fused bias gradient needs C %% 128 == 0 (C=%d)", C);
  long long blocks = (n / 2 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  (void)launch_pdl(gelu_bwd_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat162*)x, (const __nv_bfloat162*)dy,
                                                                     (__nv_bfloat162*)dx, n / 2);
  TOK_CHECK_LAUNCH("gelu_bwd");
  return TOK_OK;
}

// patch rows (+ the bf16 dy row in the backward) of one CTA, or the backward's [49][E] merge buffer, whichever is larger
static int pe_smem_bytes(int W, int E) {
  const int a = (W / 4) * kPePitch * 4 + (W / 4) * E * 2, b = E * (kPeK + 1) * 4;
  return a > b ? a : b;
}

int tok_patchify(int B, int C, int H, int W, int patch, int src_is_bf16, const void* image, void* dst, void* stream) {
  if (B <= 0 || C != 3 || patch != 4 || H <= 0 || W <= 0 || (H % 4) || (W % 4))
    return set_error(TOK_ERR_INVALID, "patchify: 3-channel images and 4x4 patches (got C=%d patch=%d H=%d W=%d)", C, patch, H, W);
  const long long total = (long long)B * (H / 4) * (W / 4) * 4;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (src_is_bf16)
    (void)launch_pdl(patchify_kernel<__nv_bfloat16, 4, 3>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                     (const __nv_bfloat16*)image, (__nv_bfloat16*)dst, B, H, W, total);
  else
    (void)launch_pdl(patchify_kernel<float, 4, 3>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                     (const float*)image, (__nv_bfloat16*)dst, B, H, W, total);
  TOK_CHECK_LAUNCH("patchify");
  return TOK_OK;
}

int tok_patch_embed_supported(int Cin, int patch, int H, int W, int E) {
  if (Cin != 3 || patch != 4 || H <= 0 || W <= 0 || (H % 4) || (W % 4)) return 0;
  if (E < 32 || (E % 32) || 2 * E > 512) return 0;
  if (pe_smem_bytes(W, E) > 160 * 1024) return 0;
  return 1;
}

int tok_patch_embed_fwd(int B, int H, int W, int E, const float* image, const float* weight, const float* bias,
                        const int* wstride, void* tokens, void* stream) {
  if (B <= 0 || !tok_patch_embed_supported(3, 4, H, W, E))
    return set_error(TOK_ERR_INVALID, "patch_embed: unsupported shape (tok_patch_embed_supported)");
  const int smem = pe_smem_bytes(W, E);
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(patch_embed_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(patch_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "patch_embed: %s", cudaGetErrorString(e));
    configured = smem;
  }
  long long ctas = (long long)B * (H / 4);
  if (ctas > 148 * 3) ctas = 148 * 3;
  (void)launch_pdl(patch_embed_fwd_kernel, dim3((unsigned)ctas), dim3(2 * E), smem, (cudaStream_t)stream, 
      image, weight, bias, B, H, W, E, wstride[0], wstride[1], wstride[2], wstride[3], (__nv_bfloat16*)tokens);
  TOK_CHECK_LAUNCH("patch_embed_fwd");
  return TOK_OK;
}

int tok_patch_embed_bwd(int B, int H, int W, int E, const float* image, const void* dtokens, const int* wstride,
                        float* dweight, float* dbias, void* stream) {
  if (B <= 0 || !tok_patch_embed_supported(3, 4, H, W, E))
    return set_error(TOK_ERR_INVALID, "patch_embed: unsupported shape (tok_patch_embed_supported)");
  const int smem = pe_smem_bytes(W, E);
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(patch_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "patch_embed: %s", cudaGetErrorString(e));
    configured = smem;
  }
  long long ctas = (long long)B * (H / 4);
  if (ctas > 148 * 3) ctas = 148 * 3;
  (void)launch_pdl(patch_embed_bwd_kernel, dim3((unsigned)ctas), dim3(2 * E), smem, (cudaStream_t)stream, 
      image, (const __nv_bfloat16*)dtokens, B, H, W, E, wstride[0], wstride[1], wstride[2], wstride[3], dweight, dbias);
  TOK_CHECK_LAUNCH("patch_embed_bwd");
  return TOK_OK;
}

int tok_patch_merge(int B, int H, int W, int C, const void* src, void* dst, int inverse, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || C <= 0 || (C % 8))
    return set_error(TOK_ERR_INVALID, "patch_merge: even H, W and C %% 8 == 0 required (H=%d W=%d C=%d)", H, W, C);
  const long long total = (long long)B * H * W * (C / 8);
  long long blocks = (total + 255) / 256;
  static const int wave = full_wave_ctas(patch_merge_kernel, 256, 0);
  if (blocks > wave) blocks = wave;
  (void)launch_pdl(patch_merge_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, B, H, W, C / 8, (const uint4*)src, (uint4*)dst, inverse);
  TOK_CHECK_LAUNCH("patch_merge");
  return TOK_OK;
}

static int attn_geom(AttnGeom* g, int B, int H, int W, int C, int heads, int ws, int shift) {
  if (B <= 0 || H <= 0 || W <= 0 || heads <= 0 || ws <= 0) return set_error(TOK_ERR_INVALID, "window_attn: bad shape");
  if (C != heads * kHd) return set_error(TOK_ERR_INVALID, "window_attn: head dimension must be 32 (C=%d heads=%d)", C, heads);
  if (ws * ws > kBigMaxN) return set_error(TOK_ERR_INVALID, "window_attn: window %d exceeds the 24x24 limit", ws);
  if ((H % ws) || (W % ws) || shift < 0 || shift >= ws) return set_error(TOK_ERR_INVALID, "window_attn: H, W must be multiples of the window; 0 <= shift < window");
  if ((long long)B * H * W >= (1LL << 31)) return set_error(TOK_ERR_INVALID, "window_attn: B*H*W must be below 2^31 (token rows are exchanged as 32-bit indices)");
  g->B = B; g->H = H; g->W = W; g->C = C; g->heads = heads; g->ws = ws; g->shift = shift;
  g->nwy = H / ws; g->nwx = W / ws;
  return TOK_OK;
}

static long long* g_attn_prof = nullptr;   // TOK_ATTN_PROFILE=1: phase clocks of the last forward launch (bring-up aid)
int tok_debug_attn_profile(long long* host_out, int max_entries) {
  if (!g_attn_prof || max_entries <= 0) return 0;
  const int n = max_entries < 64 ? max_entries : 64;
  if (cudaDeviceSynchronize() != cudaSuccess) return 0;
  if (cudaMemcpy(host_out, g_attn_prof, n * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return n;
}

int tok_window_attn_fwd(int B, int H, int W, int C, int heads, int ws, int shift, const void* qkv,
                        const float* logit_scale, const float* bias, void* out, void* stream) {
  AttnGeom g;
  int rc = attn_geom(&g, B, H, W, C, heads, ws, shift);
  if (rc) return rc;
  if (ws * ws > kMaxN) {   // windows 12 / 16 / 24: the large-window CUDA-core kernel
    if (C % 8) return set_error(TOK_ERR_INVALID, "window_attn: C must be a multiple of 8");
    const int smem = (2 * ws * ws * kPitch) * 4 + ws * ws * 4 + 64;
    static int configured_big = 0;
    if (configured_big < smem) {
      cudaError_t e = cudaFuncSetAttribute(window_attn_big_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBigSmem);
      if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "window_attn_fwd(big): %s", cudaGetErrorString(e));
      configured_big = kBigSmem;
    }
    const long long ctas = (long long)B * g.nwy * g.nwx * heads;
    (void)launch_pdl(window_attn_big_fwd_kernel, dim3((unsigned)ctas), dim3(kBigThreads), smem, (cudaStream_t)stream, 
        g, (const __nv_bfloat16*)qkv, logit_scale, bias, (__nv_bfloat16*)out);
    TOK_CHECK_LAUNCH("window_attn_big_fwd");
    return TOK_OK;
  }
  static const bool use_cuda_cores = getenv("TOK_ATTN_CUDA_CORES") != nullptr;  // bring-up aid: the round-1 fp32 kernel
  static const bool use_v1 = getenv("TOK_ATTN_FWD_V1") != nullptr;   // the one-thread-per-row tcgen05 kernel
  if (!use_cuda_cores && !use_v1 && (C % 8) == 0) {
    static bool configured2 = false;
    if (!configured2) {
      cudaError_t e = cudaFuncSetAttribute(window_attn_fwd_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTc2Smem);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(window_attn_fwd_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTc2Smem);
      if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "window_attn_fwd: %s", cudaGetErrorString(e));
      configured2 = true;
    }
    const int pairs = (B * g.nwy * g.nwx + 1) / 2;
    int groups = (148 * 2) / heads;   // two CTAs per SM (shared memory, registers); one more CTA would be a second wave
    if (groups < 1) groups = 1;
    if (groups > pairs) groups = pairs;
    static const bool want_prof = getenv("TOK_ATTN_PROFILE") != nullptr;
    if (want_prof && !g_attn_prof) {
      cudaMalloc(&g_attn_prof, 64 * sizeof(long long));
      cudaMemset(g_attn_prof, 0, 64 * sizeof(long long));
    }
    if (want_prof)
      (void)launch_pdl(window_attn_fwd_tc2_kernel<true>, dim3((unsigned)(groups * heads)), dim3(kTc2Threads + 32), kTc2Smem,
                       (cudaStream_t)stream, g, groups, (const __nv_bfloat16*)qkv, logit_scale, bias, (__nv_bfloat16*)out,
                       g_attn_prof, getenv("TOK_ATTN_DBG") ? atoi(getenv("TOK_ATTN_DBG")) : 0);
    else
      (void)launch_pdl(window_attn_fwd_tc2_kernel<false>, dim3((unsigned)(groups * heads)), dim3(kTc2Threads + 32), kTc2Smem,
                       (cudaStream_t)stream, g, groups, (const __nv_bfloat16*)qkv, logit_scale, bias, (__nv_bfloat16*)out,
                       (long long*)nullptr, 0);
    TOK_CHECK_LAUNCH("window_attn_fwd_tc2");
    return TOK_OK;
  }
  if (!use_cuda_cores && (C % 8) == 0) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(window_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
      if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "window_attn_fwd: %s", cudaGetErrorString(e));
      configured = true;
    }
    const int pairs = (B * g.nwy * g.nwx + 1) / 2;
    int groups = (148 * 2) / heads;   // two CTAs per SM (shared memory); one more CTA would be a second wave
    if (groups < 1) groups = 1;
    if (groups > pairs) groups = pairs;
    (void)launch_pdl(window_attn_fwd_tc_kernel, dim3((unsigned)(groups * heads)), dim3(kTcThreads), kTcSmem, (cudaStream_t)stream, 
        g, groups, (const __nv_bfloat16*)qkv, logit_scale, bias, (__nv_bfloat16*)out);
    TOK_CHECK_LAUNCH("window_attn_fwd_tc");
    return TOK_OK;
  }
  const long long ctas = (long long)B * g.nwy * g.nwx * heads;
  (void)launch_pdl(window_attn_fwd_kernel, dim3((unsigned)ctas), dim3(64), 0, (cudaStream_t)stream, g, (const __nv_bfloat16*)qkv, logit_scale, bias,
                                                                         (__nv_bfloat16*)out);
  TOK_CHECK_LAUNCH("window_attn_fwd");
  return TOK_OK;
}

int tok_window_attn_bwd(int B, int H, int W, int C, int heads, int ws, int shift, const void* qkv,
                        const float* logit_scale, const float* bias, const void* dout, void* dqkv, float* dbias,
                        float* dlogit_scale, float* dqkv_colsum, void* stream) {
  AttnGeom g;
  int rc = attn_geom(&g, B, H, W, C, heads, ws, shift);
  if (rc) return rc;
  const int windows = B * g.nwy * g.nwx;
  if (ws * ws > kMaxN) {   // windows 12 / 16 / 24
    if (C % 8) return set_error(TOK_ERR_INVALID, "window_attn: C must be a multiple of 8");
    if (dqkv_colsum) return set_error(TOK_ERR_INVALID, "window_attn_bwd: dqkv_colsum is not produced for windows > 8x8");
    const int n = ws * ws;
    const int smem = (2 * n * kPitch + 4 * n) * 4 + n * 4 + 32 * 4 + 64;
    static bool configured_big = false;
    if (!configured_big) {
      cudaError_t e = cudaFuncSetAttribute(window_attn_big_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBigSmem);
      if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "window_attn_bwd(big): %s", cudaGetErrorString(e));
      configured_big = true;
    }
    (void)launch_pdl(window_attn_big_bwd_kernel, dim3((unsigned)((long long)windows * heads)), dim3(kBigThreads), smem, (cudaStream_t)stream, 
        g, (const __nv_bfloat16*)qkv, logit_scale, bias, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dqkv, dbias,
        dlogit_scale);
    TOK_CHECK_LAUNCH("window_attn_big_bwd");
    return TOK_OK;
  }
  const char* cc = getenv("TOK_ATTN_BWD_CUDA_CORES");  // bring-up aid: the round-1 fp32 kernel (read per call)
  if (!(cc && cc[0] == '1') && (C % 8) == 0) {
    static bool configured_tc = false;
    if (!configured_tc) {
      cudaError_t e = cudaFuncSetAttribute(window_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwSmem);
      if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "window_attn_bwd_tc: %s", cudaGetErrorString(e));
      configured_tc = true;
    }
    const int pairs = (windows + 1) / 2;
    int groups = 148 / heads;   // one CTA per SM (shared memory), never a second wave
    if (groups < 1) groups = 1;
    if (groups > pairs) groups = pairs;
    (void)launch_pdl(window_attn_bwd_tc_kernel, dim3((unsigned)(groups * heads)), dim3(kBwThreads), kBwSmem, (cudaStream_t)stream, 
        g, groups, (const __nv_bfloat16*)qkv, logit_scale, bias, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dqkv, dbias,
        dlogit_scale, dqkv_colsum);
    TOK_CHECK_LAUNCH("window_attn_bwd_tc");
    return TOK_OK;
  }
  if (dqkv_colsum) return set_error(TOK_ERR_INVALID, "window_attn_bwd: dqkv_colsum is only produced by the tcgen05 kernel");
  int groups = (148 * 16 + heads - 1) / heads;
  if (groups > windows) groups = windows;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(window_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnBwdSmem);
    if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "window_attn_bwd: %s", cudaGetErrorString(e));
    configured = true;
  }
  (void)launch_pdl(window_attn_bwd_kernel, dim3((unsigned)(groups * heads)), dim3(64), kAttnBwdSmem, (cudaStream_t)stream, 
      g, groups, (const __nv_bfloat16*)qkv, logit_scale, bias, (const __nv_bfloat16*)dout, (__nv_bfloat16*)dqkv, dbias,
      dlogit_scale);
  TOK_CHECK_LAUNCH("window_attn_bwd");
  return TOK_OK;
}

int tok_cpb_bias_fwd(int ws, int heads, int hidden_dim, const float* coords, const float* w1, const float* b1,
                     const float* w2, float* hidden, float* table, float* bias, void* stream) {
  if (ws <= 0 || heads <= 0 || hidden_dim != kCpbHidden)
    return set_error(TOK_ERR_INVALID, "cpb_bias: the cpb MLP of timm's WindowAttention has 512 hidden units");
  const int T = (2 * ws - 1) * (2 * ws - 1);
  cudaStream_t st = (cudaStream_t)stream;
  (void)launch_pdl(cpb_table_kernel, dim3(T), dim3(128), 0, st, coords, w1, b1, w2, hidden, table, heads);
  const long long total = (long long)heads * ws * ws * ws * ws;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  (void)launch_pdl(cpb_gather_kernel, dim3((unsigned)blocks), dim3(256), 0, st, table, bias, heads, ws);
  TOK_CHECK_LAUNCH("cpb_bias_fwd");
  return TOK_OK;
}

int tok_cpb_bias_bwd(int ws, int heads, int hidden_dim, const float* coords, const float* w2, const float* hidden,
                     const float* table, const float* dbias, float* dtable, float* dw1, float* db1, float* dw2,
                     void* stream) {
  if (ws <= 0 || heads <= 0 || hidden_dim != kCpbHidden)
    return set_error(TOK_ERR_INVALID, "cpb_bias: the cpb MLP of timm's WindowAttention has 512 hidden units");
  const int T = (2 * ws - 1) * (2 * ws - 1);
  cudaStream_t st = (cudaStream_t)stream;
  (void)launch_pdl(cpb_scatter_kernel, dim3(T), dim3(128), 0, st, dbias, table, dtable, heads, ws);
  (void)launch_pdl(cpb_mlp_bwd_kernel, dim3(kCpbHidden), dim3(128), 0, st, dtable, hidden, coords, w2, dw1, db1, dw2, T, heads);
  TOK_CHECK_LAUNCH("cpb_bias_bwd");
  return TOK_OK;
}

}  // extern "C"
