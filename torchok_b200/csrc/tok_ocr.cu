// tok_ocr.cu — the object-context head and the U-Net decoder glue.
//
// Reference call sites replaced:
//   SpatialGather_Module.forward   torchok/models/heads/segmentation/ocr.py:37-46   softmax over the H*W positions of every
//                                  class map, context[b, k, :] = sum_hw p[b, k, hw] * feats[b, hw, :]
//   ObjectAttentionBlock.forward   ocr.py:77-101 (scale 1)  sim = softmax_k(key_channels^-.5 * query . key), out = sim . value
//   nn.Dropout2d                   ocr.py:126  per-(sample, channel) scale
//   DecoderBlock.forward           torchok/models/necks/segmentation/unet.py:40-58  nearest x2 upsample (+ nearest resize of
//                                  the skip) + torch.cat, written straight into the padded NHWC concat buffer
// Both products have a tiny dimension (the number of classes K, tens) against H*W pixels, so they are HBM / latency
// bound gathers, not GEMMs: CUDA-core kernels with the small operand resident in shared memory, every large tensor read
// once (twice for the gather: statistics pass + accumulation pass).  All activations NHWC bf16 with channel pitch % 8.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/tokb200.h"
#include "tok_internal.h"
#include "tok_ptx.cuh"

namespace tok {
namespace {

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
inline unsigned grid_for(long long total, int per = 256) {
  long long b = (total + per - 1) / per;
  const long long cap = 148LL * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

// ------------------------------------------------------------------------------------------------ nearest + concat
// dst[n, h, w, off + c] = src[n, floor(h * hi / ho), floor(w * wi / wo), c]   (F.interpolate mode='nearest')
__global__ void __launch_bounds__(256)
nearest_fwd_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long total, int hi, int wi, int ho, int wo,
                   int cvec, int dvec, int doff) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int w = (int)(pix % wo);
    pix /= wo;
    const int h = (int)(pix % ho);
    const long long n = pix / ho;
    const int sh = (int)(((long long)h * hi) / ho), sw = (int)(((long long)w * wi) / wo);
    dst[((n * ho + h) * wo + w) * dvec + doff + cv] = __ldg(src + ((n * hi + sh) * wi + sw) * cvec + cv);
  }
}
// dsrc[n, sh, sw, c] = sum of dout over the destination pixels that read (sh, sw): rows ceil(sh*ho/hi) .. ceil((sh+1)*ho/hi)-1
__global__ void __launch_bounds__(256)
nearest_bwd_kernel(const uint4* __restrict__ dout, uint4* __restrict__ dsrc, long long total, int hi, int wi, int ho,
                   int wo, int cvec, int dvec, int doff) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int sw = (int)(pix % wi);
    pix /= wi;
    const int sh = (int)(pix % hi);
    const long long n = pix / hi;
    const int h0 = (int)(((long long)sh * ho + hi - 1) / hi), h1 = (int)(((long long)(sh + 1) * ho + hi - 1) / hi);
    const int w0 = (int)(((long long)sw * wo + wi - 1) / wi), w1 = (int)(((long long)(sw + 1) * wo + wi - 1) / wi);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) {
        float f[8];
        unpack8(__ldg(dout + ((n * ho + h) * wo + w) * dvec + doff + cv), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
    dsrc[i] = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------------ Dropout2d scale
__global__ void __launch_bounds__(256)
channel_scale_kernel(const uint4* __restrict__ x, const float* __restrict__ scale, uint4* __restrict__ out,
                     long long total, long long hw, int cvec) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const long long n = i / (hw * cvec);
    float f[8];
    unpack8(__ldg(x + i), f);
    const float* s = scale + (n * cvec + cv) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= __ldg(s + j);
    out[i] = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------ spatial gather
// pass 1: per (b, class) max and sum of exp over the H*W positions.  grid (B, ceil(Kp / 8)); a thread walks positions with
// one 8-class vector of the logits per position (online max / sum), then the CTA merges its 256 partial pairs.
__global__ void __launch_bounds__(256)
gather_stats_kernel(const uint4* __restrict__ logits, float* __restrict__ stats, int hw, int kvec, int K) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int b = blockIdx.x, kv = blockIdx.y;
  __shared__ float sm[256][8], ss[256][8];
  float m[8], s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m[j] = -INFINITY;
    s[j] = 0.f;
  }
  const uint4* base = logits + (long long)b * hw * kvec + kv;
  for (int p = threadIdx.x; p < hw; p += 256) {
    float f[8];
    unpack8(__ldg(base + (long long)p * kvec), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float nm = fmaxf(m[j], f[j]);
      s[j] = s[j] * __expf(m[j] - nm) + __expf(f[j] - nm);
      m[j] = nm;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sm[threadIdx.x][j] = m[j];
    ss[threadIdx.x][j] = s[j];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int j = threadIdx.x;
    float M = -INFINITY;
    for (int t = 0; t < 256; ++t) M = fmaxf(M, sm[t][j]);
    float S = 0.f;
    for (int t = 0; t < 256; ++t)
      if (ss[t][j] > 0.f) S += ss[t][j] * __expf(sm[t][j] - M);
    const int k = kv * 8 + j;
    if (k < K) {
      stats[((long long)b * K + k) * 2] = M;
      stats[((long long)b * K + k) * 2 + 1] = S;
    }
  }
}

// pass 2: ctx[b, k, :] += sum over a slab of positions of p[b, k, hw] * feats[b, hw, :]   (fp32 atomics, then cast)
// grid (B, ceil(K / 8), splits); thread = (channel vector cv, position lane); 8 classes x 8 channels per thread.
__global__ void __launch_bounds__(256)
gather_ctx_kernel(const uint4* __restrict__ feats, const __nv_bfloat16* __restrict__ logits,
                  const float* __restrict__ stats, float* __restrict__ ctx, int hw, int cvec, int kp, int K, int per_split) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int b = blockIdx.x, k0 = blockIdx.y * 8;
  const int lanes = 256 / cvec;
  const int cv = threadIdx.x % cvec, pl = threadIdx.x / cvec;
  float mx[8], inv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = k0 + j;
    mx[j] = k < K ? stats[((long long)b * K + k) * 2] : 0.f;
    inv[j] = k < K ? 1.f / stats[((long long)b * K + k) * 2 + 1] : 0.f;
  }
  float acc[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
  const int p0 = blockIdx.z * per_split;
  const int p1 = min(hw, p0 + per_split);
  if (pl < lanes) {
    for (int p = p0 + pl; p < p1; p += lanes) {
      float f[8];
      unpack8(__ldg(feats + ((long long)b * hw + p) * cvec + cv), f);
      const __nv_bfloat16* lp = logits + ((long long)b * hw + p) * kp + k0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // the probability the reference stores in bf16 under autocast is NOT materialised there either (fp32 softmax)
        const float pr = (k0 + j < K) ? __expf(__bfloat162float(lp[j]) - mx[j]) * inv[j] : 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(pr, f[c], acc[j][c]);
      }
    }
  }
  // lanes of the same channel vector are 'cvec' threads apart: reduce through shared memory, then one atomic per value
  __shared__ float red[256 * 8];
  for (int j = 0; j < 8; ++j) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) red[threadIdx.x * 8 + c] = (pl < lanes) ? acc[j][c] : 0.f;
    __syncthreads();
    if (k0 + j < K) {
      for (int o = threadIdx.x; o < cvec * 8; o += 256) {
        const int v = o / 8, c = o % 8;
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += red[(l * cvec + v) * 8 + c];
        atomicAdd(ctx + (((long long)b * K + k0 + j) * cvec + v) * 8 + c, t);
      }
    }
  }
}

__global__ void cast_f32_bf16_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}

// backward: one warp per position.  dctx[b] (K x C, bf16) and D[k] = sum_c dctx[k, c] * ctx[k, c] sit in shared memory.
//   dp_k = dctx[k, :] . feats[hw, :]     dlogit[hw, k] = p_k * (dp_k - D_k)     dfeats[hw, :] = sum_k p_k * dctx[k, :]
__global__ void __launch_bounds__(256)
gather_bwd_kernel(const __nv_bfloat16* __restrict__ feats, const __nv_bfloat16* __restrict__ logits,
                  const float* __restrict__ stats, const __nv_bfloat16* __restrict__ ctx,
                  const __nv_bfloat16* __restrict__ dctx, __nv_bfloat16* __restrict__ dfeats,
                  __nv_bfloat16* __restrict__ dlogits, int hw, int C, int kp, int K, int rows_per_cta) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ float sh[];
  float* s_d = sh;               // [K][C] dctx as fp32
  float* s_D = sh + K * C;       // [K]
  float* s_mx = s_D + K;         // [K]
  float* s_inv = s_mx + K;       // [K]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < K * C; i += 256) s_d[i] = __bfloat162float(dctx[(long long)b * K * C + i]);
  for (int k = threadIdx.x; k < K; k += 256) {
    s_mx[k] = stats[((long long)b * K + k) * 2];
    s_inv[k] = 1.f / stats[((long long)b * K + k) * 2 + 1];
  }
  __syncthreads();
  for (int k = warp; k < K; k += 8) {
    float t = 0.f;
    for (int c = lane; c < C; c += 32) t += s_d[k * C + c] * __bfloat162float(ctx[((long long)b * K + k) * C + c]);
    t = warp_sum(t);
    if (lane == 0) s_D[k] = t;
  }
  __syncthreads();
  const int p0 = blockIdx.y * rows_per_cta, p1 = min(hw, p0 + rows_per_cta);
  for (int p = p0 + warp; p < p1; p += 8) {
    const __nv_bfloat16* fr = feats + ((long long)b * hw + p) * C;
    const __nv_bfloat16* lr = logits + ((long long)b * hw + p) * kp;
    __nv_bfloat16* dl = dlogits + ((long long)b * hw + p) * kp;
    __nv_bfloat16* dfr = dfeats + ((long long)b * hw + p) * C;
    for (int c0 = 0; c0 < C; c0 += 32 * 8) {   // 8 channels per lane and sweep; C <= 2048 in practice
      const int c = c0 + lane * 8;
      float f[8], o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
      const bool on = c < C;
      if (on) unpack8(__ldg(reinterpret_cast<const uint4*>(fr + c)), f);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
      for (int k = 0; k < K; ++k) {
        const float pr = __expf(__bfloat162float(lr[k]) - s_mx[k]) * s_inv[k];
        if (on) {
          const float* dk = s_d + k * C + c;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(pr, dk[j], o[j]);
        }
        if (C <= 256) {   // single sweep: the dot product is complete here
          float t = 0.f;
          if (on) {
            const float* dk = s_d + k * C + c;
#pragma unroll
            for (int j = 0; j < 8; ++j) t = fmaf(dk[j], f[j], t);
          }
          t = warp_sum(t);
          if (lane == 0) dl[k] = __float2bfloat16(pr * (t - s_D[k]));
        }
      }
      if (on) *reinterpret_cast<uint4*>(dfr + c) = pack8(o);
    }
    if (C > 256) {   // several sweeps: the dot products need all of them
      for (int k = 0; k < K; ++k) {
        float t = 0.f;
        for (int c = lane * 8; c < C; c += 256) {
          float f[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(fr + c)), f);
          const float* dk = s_d + k * C + c;
#pragma unroll
          for (int j = 0; j < 8; ++j) t = fmaf(dk[j], f[j], t);
        }
        t = warp_sum(t);
        const float pr = __expf(__bfloat162float(lr[k]) - s_mx[k]) * s_inv[k];
        if (lane == 0) dl[k] = __float2bfloat16(pr * (t - s_D[k]));
      }
    }
    if (lane == 0)
      for (int k = K; k < kp; ++k) dl[k] = __float2bfloat16(0.f);   // pad lanes of the logits gradient stay zero
  }
}

// ------------------------------------------------------------------------------------------------ object attention
// One warp per position; key / value [K][Kc] of the image resident in shared memory (fp32).  Kc <= 256.
//   s_k = scale * q . key_k   sim = softmax_k(s)   out = sum_k sim_k * value_k
template <bool BWD>
__global__ void __launch_bounds__(256)
object_attn_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ key,
                   const __nv_bfloat16* __restrict__ value, __nv_bfloat16* __restrict__ out,
                   const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dq, float* __restrict__ dkey,
                   float* __restrict__ dvalue, int hw, int Kc, int K, float scale, int rows_per_cta) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  extern __shared__ float sh[];
  float* s_key = sh;                    // [K][Kc]
  float* s_val = s_key + K * Kc;        // [K][Kc]
  float* s_sim = s_val + K * Kc;        // [8 warps][K]
  float* s_ds = s_sim + 8 * K;          // [8 warps][K]      (backward)
  float* s_dk = s_ds + 8 * K;           // [K][Kc] partial dkey   (backward)
  float* s_dv = s_dk + K * Kc;          // [K][Kc] partial dvalue (backward)
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < K * Kc; i += 256) {
    s_key[i] = __bfloat162float(key[(long long)b * K * Kc + i]);
    s_val[i] = __bfloat162float(value[(long long)b * K * Kc + i]);
    if (BWD) s_dk[i] = s_dv[i] = 0.f;
  }
  __syncthreads();
  const int per = (Kc + 31) / 32;       // channels per lane (<= 8)
  const int p0 = blockIdx.y * rows_per_cta, p1 = min(hw, p0 + rows_per_cta);
  float* sim = s_sim + warp * K;
  float* dsv = s_ds + warp * K;
  for (int p = p0 + warp; p < p1; p += 8) {
    const long long row = ((long long)b * hw + p) * Kc;
    float qv[8], gv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = lane + 32 * j;
      qv[j] = (j < per && c < Kc) ? __bfloat162float(q[row + c]) : 0.f;
      gv[j] = (BWD && j < per && c < Kc) ? __bfloat162float(dout[row + c]) : 0.f;
    }
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < per && lane + 32 * j < Kc) t = fmaf(qv[j], s_key[k * Kc + lane + 32 * j], t);
      t = warp_sum(t) * scale;
      t = bf(t);                          // the reference's autocast stores the scaled similarity in bf16
      if (lane == 0) sim[k] = t;
      mx = fmaxf(mx, t);
    }
    __syncwarp();
    float sum = 0.f;
    for (int k = 0; k < K; ++k) sum += __expf(sim[k] - mx);
    const float inv = 1.f / sum;
    if (!BWD) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
      for (int k = 0; k < K; ++k) {
        const float pr = bf(__expf(sim[k] - mx) * inv);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < per && lane + 32 * j < Kc) o[j] = fmaf(pr, s_val[k * Kc + lane + 32 * j], o[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < per && lane + 32 * j < Kc) out[row + lane + 32 * j] = __float2bfloat16(o[j]);
    } else {
      // dsim_k = dout . value_k ; ds_k = sim_k * (dsim_k - sum_j sim_j dsim_j) * scale
      float dot = 0.f;
      for (int k = 0; k < K; ++k) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < per && lane + 32 * j < Kc) t = fmaf(gv[j], s_val[k * Kc + lane + 32 * j], t);
        t = warp_sum(t);
        const float pr = __expf(sim[k] - mx) * inv;
        if (lane == 0) dsv[k] = t;
        dot = fmaf(pr, t, dot);
      }
      __syncwarp();
      float dqv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) dqv[j] = 0.f;
      for (int k = 0; k < K; ++k) {
        const float pr = __expf(sim[k] - mx) * inv;
        const float ds = pr * (dsv[k] - dot) * scale;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = lane + 32 * j;
          if (j < per && c < Kc) {
            dqv[j] = fmaf(ds, s_key[k * Kc + c], dqv[j]);
            atomicAdd(&s_dk[k * Kc + c], ds * qv[j]);
            atomicAdd(&s_dv[k * Kc + c], pr * gv[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < per && lane + 32 * j < Kc) dq[row + lane + 32 * j] = __float2bfloat16(dqv[j]);
      __syncwarp();
    }
  }
  if (BWD) {
    __syncthreads();
    for (int i = threadIdx.x; i < K * Kc; i += 256) {
      atomicAdd(dkey + (long long)b * K * Kc + i, s_dk[i]);
      atomicAdd(dvalue + (long long)b * K * Kc + i, s_dv[i]);
    }
  }
}

}  // namespace
}  // namespace tok

using namespace tok;

extern "C" {

int tok_nearest_fwd(int n, int hi, int wi, int c, int ho, int wo, const void* src, void* dst, int dst_c,
                    int dst_c_offset, void* stream) {
  if (n <= 0 || hi <= 0 || wi <= 0 || ho <= 0 || wo <= 0 || c <= 0 || (c % 8) || (dst_c % 8) || (dst_c_offset % 8) ||
      dst_c_offset + c > dst_c)
    return set_error(TOK_ERR_INVALID, "nearest_fwd: bad shape (channel counts and offsets must be multiples of 8)");
  const long long total = (long long)n * ho * wo * (c / 8);
  (void)launch_pdl(nearest_fwd_kernel, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, (const uint4*)src, (uint4*)dst, total, hi, wi, ho,
                                                                       wo, c / 8, dst_c / 8, dst_c_offset / 8);
  TOK_CHECK_LAUNCH("nearest_fwd");
  return TOK_OK;
}

int tok_nearest_bwd(int n, int hi, int wi, int c, int ho, int wo, const void* dout, int dout_c, int dout_c_offset,
                    void* dsrc, void* stream) {
  if (n <= 0 || hi <= 0 || wi <= 0 || ho <= 0 || wo <= 0 || c <= 0 || (c % 8) || (dout_c % 8) || (dout_c_offset % 8) ||
      dout_c_offset + c > dout_c)
    return set_error(TOK_ERR_INVALID, "nearest_bwd: bad shape (channel counts and offsets must be multiples of 8)");
  const long long total = (long long)n * hi * wi * (c / 8);
  (void)launch_pdl(nearest_bwd_kernel, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, (const uint4*)dout, (uint4*)dsrc, total, hi, wi,
                                                                       ho, wo, c / 8, dout_c / 8, dout_c_offset / 8);
  TOK_CHECK_LAUNCH("nearest_bwd");
  return TOK_OK;
}

int tok_channel_scale(int n, long long hw, int c, const void* x, const float* scale, void* out, void* stream) {
  if (n <= 0 || hw <= 0 || c <= 0 || (c % 8)) return set_error(TOK_ERR_INVALID, "channel_scale: bad shape");
  const long long total = (long long)n * hw * (c / 8);
  (void)launch_pdl(channel_scale_kernel, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, (const uint4*)x, scale, (uint4*)out, total, hw,
                                                                         c / 8);
  TOK_CHECK_LAUNCH("channel_scale");
  return TOK_OK;
}

int tok_spatial_gather_fwd(int b, int hw, int c, int k, int kp, const void* feats, const void* logits, float* stats,
                           float* ctx_f32, void* ctx, void* stream) {
  if (b <= 0 || hw <= 0 || c <= 0 || (c % 8) || c > 2048 || k <= 0 || kp < k || (kp % 8))
    return set_error(TOK_ERR_INVALID, "spatial_gather_fwd: need C %% 8 == 0, C <= 2048, class pitch %% 8 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  const int cvec = c / 8;
  (void)launch_pdl(gather_stats_kernel, dim3(dim3(b, kp / 8)), dim3(256), 0, st, (const uint4*)logits, stats, hw, kp / 8, k);
  cudaMemsetAsync(ctx_f32, 0, (size_t)b * k * c * 4, st);
  const int kgroups = (k + 7) / 8;
  int splits = (148 * 4 + b * kgroups - 1) / (b * kgroups);
  const int lanes = 256 / cvec > 0 ? 256 / cvec : 1;
  const int max_splits = (hw + lanes * 4 - 1) / (lanes * 4);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int per_split = (hw + splits - 1) / splits;
  if (cvec > 256) return set_error(TOK_ERR_INVALID, "spatial_gather_fwd: C too large");
  (void)launch_pdl(gather_ctx_kernel, dim3(dim3(b, kgroups, splits)), dim3(256), 0, st, (const uint4*)feats, (const __nv_bfloat16*)logits, stats,
                                                             ctx_f32, hw, cvec, kp, k, per_split);
  const long long n = (long long)b * k * c;
  (void)launch_pdl(cast_f32_bf16_rows_kernel, dim3(grid_for(n)), dim3(256), 0, st, ctx_f32, (__nv_bfloat16*)ctx, n);
  TOK_CHECK_LAUNCH("spatial_gather_fwd");
  return TOK_OK;
}

int tok_spatial_gather_bwd(int b, int hw, int c, int k, int kp, const void* feats, const void* logits,
                           const float* stats, const void* ctx, const void* dctx, void* dfeats, void* dlogits,
                           void* stream) {
  if (b <= 0 || hw <= 0 || c <= 0 || (c % 8) || c > 2048 || k <= 0 || kp < k || (kp % 8))
    return set_error(TOK_ERR_INVALID, "spatial_gather_bwd: need C %% 8 == 0, C <= 2048, class pitch %% 8 == 0");
  const size_t smem = ((size_t)k * c + 3 * (size_t)k) * 4;
  if (smem > 200 * 1024) return set_error(TOK_ERR_INVALID, "spatial_gather_bwd: K x C = %d x %d exceeds shared memory", k, c);
  cudaError_t e = cudaFuncSetAttribute(gather_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "spatial_gather_bwd: %s", cudaGetErrorString(e));
  int gy = (148 * 2 + b - 1) / b;
  if (gy > (hw + 7) / 8) gy = (hw + 7) / 8;
  const int rows_per_cta = (hw + gy - 1) / gy;
  (void)launch_pdl(gather_bwd_kernel, dim3(dim3(b, gy)), dim3(256), smem, (cudaStream_t)stream, 
      (const __nv_bfloat16*)feats, (const __nv_bfloat16*)logits, stats, (const __nv_bfloat16*)ctx,
      (const __nv_bfloat16*)dctx, (__nv_bfloat16*)dfeats, (__nv_bfloat16*)dlogits, hw, c, kp, k, rows_per_cta);
  TOK_CHECK_LAUNCH("spatial_gather_bwd");
  return TOK_OK;
}

static int attn_geometry(int b, int hw, int kc, int k, bool bwd, size_t* smem, dim3* grid, int* rows_per_cta) {
  if (b <= 0 || hw <= 0 || kc <= 0 || kc > 256 || k <= 0)
    return set_error(TOK_ERR_INVALID, "object_attention: need key_channels <= 256");
  *smem = ((size_t)(bwd ? 4 : 2) * k * kc + 16 * (size_t)k) * 4;
  if (*smem > 200 * 1024)
    return set_error(TOK_ERR_INVALID, "object_attention: K x key_channels = %d x %d exceeds shared memory", k, kc);
  int gy = (148 * 2 + b - 1) / b;
  if (gy > (hw + 7) / 8) gy = (hw + 7) / 8;
  *rows_per_cta = (hw + gy - 1) / gy;
  *grid = dim3(b, gy);
  return TOK_OK;
}

int tok_object_attn_fwd(int b, int hw, int kc, int k, float scale, const void* q, const void* key, const void* value,
                        void* out, void* stream) {
  size_t smem;
  dim3 grid;
  int rpc;
  int rc = attn_geometry(b, hw, kc, k, false, &smem, &grid, &rpc);
  if (rc) return rc;
  cudaError_t e = cudaFuncSetAttribute(object_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "object_attn_fwd: %s", cudaGetErrorString(e));
  (void)launch_pdl(object_attn_kernel<false>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, 
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)key, (const __nv_bfloat16*)value, (__nv_bfloat16*)out, nullptr,
      nullptr, nullptr, nullptr, hw, kc, k, scale, rpc);
  TOK_CHECK_LAUNCH("object_attn_fwd");
  return TOK_OK;
}

int tok_object_attn_bwd(int b, int hw, int kc, int k, float scale, const void* q, const void* key, const void* value,
                        const void* dout, void* dq, float* dkey_f32, float* dvalue_f32, void* dkey, void* dvalue,
                        void* stream) {
  size_t smem;
  dim3 grid;
  int rpc;
  int rc = attn_geometry(b, hw, kc, k, true, &smem, &grid, &rpc);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaFuncSetAttribute(object_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "object_attn_bwd: %s", cudaGetErrorString(e));
  const long long n = (long long)b * k * kc;
  cudaMemsetAsync(dkey_f32, 0, n * 4, st);
  cudaMemsetAsync(dvalue_f32, 0, n * 4, st);
  (void)launch_pdl(object_attn_kernel<true>, dim3(grid), dim3(256), smem, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)key,
                                                    (const __nv_bfloat16*)value, nullptr, (const __nv_bfloat16*)dout,
                                                    (__nv_bfloat16*)dq, dkey_f32, dvalue_f32, hw, kc, k, scale, rpc);
  (void)launch_pdl(cast_f32_bf16_rows_kernel, dim3(grid_for(n)), dim3(256), 0, st, dkey_f32, (__nv_bfloat16*)dkey, n);
  (void)launch_pdl(cast_f32_bf16_rows_kernel, dim3(grid_for(n)), dim3(256), 0, st, dvalue_f32, (__nv_bfloat16*)dvalue, n);
  TOK_CHECK_LAUNCH("object_attn_bwd");
  return TOK_OK;
}

}  // extern "C"
