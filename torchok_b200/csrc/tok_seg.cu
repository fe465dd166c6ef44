// tok_seg.cu — HBM-bound passes of the HRNet / segmentation path (NHWC bf16, 8 channels per 16-byte access).
//
// Reference call sites:
//   timm HighResolutionModule.forward (used by torchok/models/backbones/hrnet.py:167-192): y_i = sum_j fuse_ij(x_j) with
//     nn.Upsample(scale_factor=2^(j-i), mode='nearest') on the low-resolution terms, then ReLU        -> fuse_sum fwd/bwd
//   F.interpolate(mode='bilinear', align_corners=False) in HRNetSegmentationNeck.forward
//     (torchok/models/necks/segmentation/hrnet.py:35-38) followed by torch.cat, and in SegmentationHead.forward
//     (torchok/models/heads/segmentation/base.py:37)                                                  -> bilinear fwd/bwd
//   torch.nn.CrossEntropyLoss on (B, C, H, W) logits (torchok/losses/__init__.py:26)                   -> xent_small
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
inline int grid_for(long long items) {
  long long b = (items + 255) / 256;
  if (b > 148LL * 16) b = 148LL * 16;
  return (int)(b < 1 ? 1 : b);
}

struct FuseTerms {
  const uint4* src[4];
  uint4* dst[4];
  int shift[4];  // term t lives at resolution (H >> shift, W >> shift)
  int n;
};

// out[n,h,w,:] = relu(sum_t term_t[n, h >> s_t, w >> s_t, :]);  bits = (out > 0)
__global__ void __launch_bounds__(256)
fuse_sum_fwd_kernel(FuseTerms t, uint4* __restrict__ out, uint8_t* __restrict__ bits, long long total, int H, int W,
                    int cvec, int relu) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int w = (int)(pix % W);
    pix /= W;
    const int h = (int)(pix % H);
    const long long n = pix / H;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int k = 0; k < t.n; ++k) {
      const int s = t.shift[k];
      const long long o = ((n * (H >> s) + (h >> s)) * (W >> s) + (w >> s)) * cvec + cv;
      float f[8];
      unpack8(__ldg(t.src[k] + o), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    uint32_t b = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (relu) acc[j] = fmaxf(acc[j], 0.f);
      b |= (acc[j] > 0.f ? 1u : 0u) << j;
    }
    out[i] = pack8(acc);
    if (bits) bits[i] = (uint8_t)b;
  }
}

// dterm[n, h', w', :] = sum over the 2^s x 2^s patch of dout * mask   (one launch per term)
__global__ void __launch_bounds__(256)
fuse_sum_bwd_kernel(const uint4* __restrict__ dout, const uint8_t* __restrict__ bits, uint4* __restrict__ dterm,
                    long long total, int H, int W, int cvec, int shift) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int Hs = H >> shift, Ws = W >> shift, f = 1 << shift;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long pix = i / cvec;
    const int w = (int)(pix % Ws);
    pix /= Ws;
    const int h = (int)(pix % Hs);
    const long long n = pix / Hs;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int dh = 0; dh < f; ++dh) {
      for (int dw = 0; dw < f; ++dw) {
        const long long o = ((n * H + (h * f + dh)) * W + (w * f + dw)) * cvec + cv;
        float g[8];
        unpack8(__ldg(dout + o), g);
        const uint32_t b = bits ? bits[o] : 0xFFu;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += ((b >> j) & 1u) ? g[j] : 0.f;
      }
    }
    dterm[i] = pack8(acc);
  }
}

// PyTorch's upsample_bilinear2d source index (align_corners=False): src = max((dst + .5) * scale - .5, 0)
__device__ __forceinline__ void bilinear_taps(int o, float scale, int in_size, int& i0, int& i1, float& l1) {
  float s = (o + 0.5f) * scale - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

// dst[n, ho, wo, coff + c] = bilinear(src)[n, ho, wo, c]; dst has channel pitch dcvec (vectors), src pitch scvec.
__global__ void __launch_bounds__(256)
bilinear_fwd_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long total, int Hi, int Wi, int Ho,
                    int Wo, int scvec, int dcvec, int coff_vec, float sh, float sw) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % scvec);
    long long pix = i / scvec;
    const int wo = (int)(pix % Wo);
    pix /= Wo;
    const int ho = (int)(pix % Ho);
    const long long n = pix / Ho;
    int h0, h1, w0, w1;
    float lh, lw;
    bilinear_taps(ho, sh, Hi, h0, h1, lh);
    bilinear_taps(wo, sw, Wi, w0, w1, lw);
    float a[8], b[8], c[8], d[8], r[8];
    unpack8(__ldg(src + ((n * Hi + h0) * Wi + w0) * scvec + cv), a);
    unpack8(__ldg(src + ((n * Hi + h0) * Wi + w1) * scvec + cv), b);
    unpack8(__ldg(src + ((n * Hi + h1) * Wi + w0) * scvec + cv), c);
    unpack8(__ldg(src + ((n * Hi + h1) * Wi + w1) * scvec + cv), d);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      r[j] = (1.f - lh) * ((1.f - lw) * a[j] + lw * b[j]) + lh * ((1.f - lw) * c[j] + lw * d[j]);
    dst[((n * Ho + ho) * Wo + wo) * dcvec + coff_vec + cv] = pack8(r);
  }
}

// Backward as a GATHER over the input grid: every input pixel sums the output pixels whose taps touch it (no atomics,
// deterministic).  For an integer up-scale factor f the candidates are the <= 2f+1 output rows/cols around it.
__global__ void __launch_bounds__(256)
bilinear_bwd_kernel(const uint4* __restrict__ dout, uint4* __restrict__ dsrc, long long total, int Hi, int Wi, int Ho,
                    int Wo, int scvec, int dcvec, int coff_vec, float sh, float sw) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int fh = (int)ceilf(1.f / sh) + 1, fw = (int)ceilf(1.f / sw) + 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % scvec);
    long long pix = i / scvec;
    const int wi = (int)(pix % Wi);
    pix /= Wi;
    const int hi = (int)(pix % Hi);
    const long long n = pix / Hi;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int ho_c = (int)((hi + 0.5f) / sh), wo_c = (int)((wi + 0.5f) / sw);
    for (int ho = max(ho_c - fh, 0); ho <= min(ho_c + fh, Ho - 1); ++ho) {
      int h0, h1;
      float lh;
      bilinear_taps(ho, sh, Hi, h0, h1, lh);
      const float wh = (h0 == hi ? 1.f - lh : 0.f) + (h1 == hi ? lh : 0.f);
      if (wh == 0.f) continue;
      for (int wo = max(wo_c - fw, 0); wo <= min(wo_c + fw, Wo - 1); ++wo) {
        int w0, w1;
        float lw;
        bilinear_taps(wo, sw, Wi, w0, w1, lw);
        const float ww = (w0 == wi ? 1.f - lw : 0.f) + (w1 == wi ? lw : 0.f);
        if (ww == 0.f) continue;
        float g[8];
        unpack8(__ldg(dout + ((n * Ho + ho) * Wo + wo) * dcvec + coff_vec + cv), g);
        const float k = wh * ww;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(k, g[j], acc[j]);
      }
    }
    dsrc[i] = pack8(acc);
  }
}

// Cross entropy over many short rows (segmentation logits: rows = B*H*W, C <= 64): one thread per row.
__global__ void __launch_bounds__(256)
xent_small_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ target,
                  float* __restrict__ loss_sum, float* __restrict__ count, __nv_bfloat16* __restrict__ dlogits,
                  long long rows, int C, int ld, const float* __restrict__ inv_norm_dev, float gscale,
                  const float* __restrict__ gscale_dev, long long ignore_index) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float s_loss[8], s_cnt[8];
  float my_loss = 0.f, my_cnt = 0.f;
  if (gscale_dev) gscale *= __ldg(gscale_dev);
  if (inv_norm_dev && dlogits) gscale *= __ldg(inv_norm_dev);
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const __nv_bfloat16* lp = logits + r * ld;
    const long long t = target[r];
    const bool ignored = t == ignore_index;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __bfloat162float(lp[c]));
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += __expf(__bfloat162float(lp[c]) - mx);
    if (!ignored && !dlogits) {
      my_loss += logf(se) + mx - __bfloat162float(lp[t]);
      my_cnt += 1.f;
    }
    if (dlogits) {
      __nv_bfloat16* dp = dlogits + r * ld;
      const float inv = 1.f / se;
      for (int c = 0; c < ld; ++c) {
        float g = 0.f;
        if (!ignored && c < C) g = (__expf(__bfloat162float(lp[c]) - mx) * inv - (c == t ? 1.f : 0.f)) * gscale;
        dp[c] = __float2bfloat16(g);
      }
    }
  }
  if (!dlogits) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
      my_cnt += __shfl_xor_sync(0xffffffffu, my_cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
      s_loss[threadIdx.x >> 5] = my_loss;
      s_cnt[threadIdx.x >> 5] = my_cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < 8; ++w) {
        a += s_loss[w];
        b += s_cnt[w];
      }
      atomicAdd(loss_sum, a);
      atomicAdd(count, b);
    }
  }
}


// The same with the row held in registers: ld = 8 * LV channels are fetched as LV 16-byte vectors (one pass instead of
// three passes of 2-byte loads over a 48-byte row: 626 -> ~150 us for the 8.4 M pixels x 19 classes of the HRNet step).
template <int LV>
__global__ void __launch_bounds__(256)
xent_small_vec_kernel(const uint4* __restrict__ logits, const long long* __restrict__ target,
                      float* __restrict__ loss_sum, float* __restrict__ count, uint4* __restrict__ dlogits,
                      long long rows, int C, const float* __restrict__ inv_norm_dev, float gscale,
                      const float* __restrict__ gscale_dev, long long ignore_index) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float s_loss[8], s_cnt[8];
  float my_loss = 0.f, my_cnt = 0.f;
  if (gscale_dev) gscale *= __ldg(gscale_dev);
  if (inv_norm_dev && dlogits) gscale *= __ldg(inv_norm_dev);
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    float v[LV * 8];
#pragma unroll
    for (int q = 0; q < LV; ++q) {
      const uint4 u = __ldg(logits + r * LV + q);
      const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[q * 8 + 2 * e] = __uint_as_float(w4[e] << 16);
        v[q * 8 + 2 * e + 1] = __uint_as_float(w4[e] & 0xFFFF0000u);
      }
    }
    const long long t = target[r];
    const bool ignored = t == ignore_index;
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < LV * 8; ++c)
      if (c < C) mx = fmaxf(mx, v[c]);
    float se = 0.f, vt = 0.f;
#pragma unroll
    for (int c = 0; c < LV * 8; ++c) {
      if (c < C) {
        if (c == t) vt = v[c];
        v[c] = __expf(v[c] - mx);
        se += v[c];
      }
    }
    if (!ignored && !dlogits) {
      my_loss += logf(se) + mx - vt;
      my_cnt += 1.f;
    }
    if (dlogits) {
      const float inv = 1.f / se;
#pragma unroll
      for (int q = 0; q < LV; ++q) {
        uint32_t w4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float g2[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c = q * 8 + 2 * e + h;
            g2[h] = (!ignored && c < C) ? (v[c] * inv - (c == t ? 1.f : 0.f)) * gscale : 0.f;
          }
          __nv_bfloat162 b2 = __floats2bfloat162_rn(g2[0], g2[1]);
          w4[e] = *reinterpret_cast<uint32_t*>(&b2);
        }
        dlogits[r * LV + q] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
    }
  }
  if (!dlogits) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      my_loss += __shfl_xor_sync(0xffffffffu, my_loss, o);
      my_cnt += __shfl_xor_sync(0xffffffffu, my_cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
      s_loss[threadIdx.x >> 5] = my_loss;
      s_cnt[threadIdx.x >> 5] = my_cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < 8; ++w) {
        a += s_loss[w];
        b += s_cnt[w];
      }
      atomicAdd(loss_sum, a);
      atomicAdd(count, b);
    }
  }
}

// ---- Dice loss (multiclass, from logits) ------------------------------------------------------------------------------
// torchok/losses/segmentation/dice.py:85-188 (DiceLoss, mode='multiclass', from_logits=True) with soft_dice_score
// (dice.py:23-56) over dims (0, 2): per class c
//   I_c = sum p_c [t = c],  card_c = sum p_c + sum [t = c],  score_c = (2 I_c + smooth) / (max(card_c, eps) + smooth)
//   loss = mean_c [count_c > 0] (1 - score_c)      (or -log(max(score_c, eps)) with log_loss)
// where p = softmax(logits) per pixel.  Pass 1 (one thread per pixel, C <= 64) accumulates I, sum p, count; a
// one-warp finalize turns them into the scalar loss and the per-class coefficients a_c, b_c of
// dloss/dp_c(pixel) = a_c [t = c] + b_c; pass 2 recomputes the softmax and writes dlogits = p o (dp - sum_k p_k dp_k).
__global__ void __launch_bounds__(256)
dice_stats_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ target,
                  float* __restrict__ stats /* [3][64]: I, sum p, count */, long long rows, int C, int ld) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float sh[3][64];
  for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) (&sh[0][0])[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // warp-uniform trip count: every lane takes part in the shuffles, rows past the end contribute zeros
  for (long long base = blockIdx.x * (long long)blockDim.x + (threadIdx.x - lane); base < rows; base += stride) {
    const long long r = base + lane;
    const bool live = r < rows;
    const __nv_bfloat16* lp = logits + (live ? r : 0) * ld;
    const long long t = live ? target[r] : -1;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __bfloat162float(lp[c]));
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += __expf(__bfloat162float(lp[c]) - mx);
    const float inv = live ? 1.f / se : 0.f;
    for (int c = 0; c < C; ++c) {
      float pc = __expf(__bfloat162float(lp[c]) - mx) * inv;
      float ic = t == c ? pc : 0.f;
      const unsigned hit = __ballot_sync(0xffffffffu, t == c);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        pc += __shfl_xor_sync(0xffffffffu, pc, o);
        ic += __shfl_xor_sync(0xffffffffu, ic, o);
      }
      if (lane == 0) {
        atomicAdd(&sh[0][c], ic);
        atomicAdd(&sh[1][c], pc);
        atomicAdd(&sh[2][c], (float)__popc(hit));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) {
    const float v = (&sh[0][0])[i];
    if (v != 0.f) atomicAdd(stats + i, v);
  }
}

__global__ void dice_finalize_kernel(const float* __restrict__ stats, int C, float smooth, float eps, int log_loss,
                                     float* __restrict__ loss, float* __restrict__ coef /* [2][64]: a, b */) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  const int c = threadIdx.x;  // 64 threads
  float l = 0.f;
  if (c < C) {
    const float I = stats[c], card = stats[64 + c] + stats[128 + c], cnt = stats[128 + c];
    const float D = fmaxf(card, eps) + smooth;
    const float num = 2.f * I + smooth;
    const float score = num / D;
    const float mask = cnt > 0.f ? 1.f / C : 0.f;   // 1/C of the mean folded in
    float dscore = -1.f;                              // d loss_c / d score_c
    if (log_loss) {
      l = -logf(fmaxf(score, eps));
      dscore = score > eps ? -1.f / score : 0.f;
    } else {
      l = 1.f - score;
    }
    l *= mask;
    coef[c] = mask * dscore * 2.f / D;                                       // a_c: through I_c
    coef[64 + c] = card > eps ? -mask * dscore * num / (D * D) : 0.f;        // b_c: through card_c (clamp_min)
  } else {
    coef[c] = 0.f;
    coef[64 + c] = 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  __shared__ float part[2];
  if ((c & 31) == 0) part[c >> 5] = l;
  __syncthreads();
  if (c == 0) *loss = part[0] + part[1];
}

__global__ void __launch_bounds__(256)
dice_bwd_kernel(const __nv_bfloat16* __restrict__ logits, const long long* __restrict__ target,
                const float* __restrict__ coef, const float* __restrict__ gscale_dev,
                __nv_bfloat16* __restrict__ dlogits, long long rows, int C, int ld) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float sa[64], sb[64];
  if (threadIdx.x < 64) {
    sa[threadIdx.x] = coef[threadIdx.x];
    sb[threadIdx.x] = coef[64 + threadIdx.x];
  }
  __syncthreads();
  const float gs = gscale_dev ? __ldg(gscale_dev) : 1.f;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    const __nv_bfloat16* lp = logits + r * ld;
    const long long t = target[r];
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __bfloat162float(lp[c]));
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += __expf(__bfloat162float(lp[c]) - mx);
    const float inv = 1.f / se;
    float dot = 0.f;   // sum_k p_k dp_k
    for (int c = 0; c < C; ++c) {
      const float pc = __expf(__bfloat162float(lp[c]) - mx) * inv;
      dot = fmaf(pc, (t == c ? sa[c] : 0.f) + sb[c], dot);
    }
    __nv_bfloat16* dp = dlogits + r * ld;
    for (int c = 0; c < ld; ++c) {
      float g = 0.f;
      if (c < C) {
        const float pc = __expf(__bfloat162float(lp[c]) - mx) * inv;
        g = pc * ((t == c ? sa[c] : 0.f) + sb[c] - dot) * gs;
      }
      dp[c] = __float2bfloat16(g);
    }
  }
}

}  // namespace
}  // namespace tok

using namespace tok;

extern "C" {

int tok_fuse_sum_fwd(int n, int h, int w, int c, int nterms, const void* const* terms, const int* shifts, int relu,
                     void* out, void* bits, void* stream) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8)) return set_error(TOK_ERR_INVALID, "fuse_sum_fwd: bad shape");
  if (nterms < 1 || nterms > 4) return set_error(TOK_ERR_INVALID, "fuse_sum_fwd: 1..4 terms");
  FuseTerms t;
  t.n = nterms;
  for (int k = 0; k < nterms; ++k) {
    if (shifts[k] < 0 || (h % (1 << shifts[k])) || (w % (1 << shifts[k])))
      return set_error(TOK_ERR_INVALID, "fuse_sum_fwd: spatial size %dx%d is not divisible by 2^%d", h, w, shifts[k]);
    t.src[k] = (const uint4*)terms[k];
    t.shift[k] = shifts[k];
    t.dst[k] = nullptr;
  }
  const long long total = (long long)n * h * w * (c / 8);
  (void)launch_pdl(fuse_sum_fwd_kernel, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, t, (uint4*)out, (uint8_t*)bits, total, h, w,
                                                                        c / 8, relu);
  TOK_CHECK_LAUNCH("fuse_sum_fwd");
  return TOK_OK;
}

int tok_fuse_sum_bwd(int n, int h, int w, int c, int shift, const void* dout, const void* bits, void* dterm,
                     void* stream) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8) || shift < 0) return set_error(TOK_ERR_INVALID, "fuse_sum_bwd: bad shape");
  const long long total = (long long)n * (h >> shift) * (w >> shift) * (c / 8);
  (void)launch_pdl(fuse_sum_bwd_kernel, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, (const uint4*)dout, (const uint8_t*)bits,
                                                                        (uint4*)dterm, total, h, w, c / 8, shift);
  TOK_CHECK_LAUNCH("fuse_sum_bwd");
  return TOK_OK;
}

int tok_bilinear_fwd(int n, int hi, int wi, int c, int ho, int wo, const void* src, void* dst, int dst_c,
                     int dst_c_offset, void* stream) {
  if (n <= 0 || hi <= 0 || wi <= 0 || ho <= 0 || wo <= 0 || c <= 0 || (c % 8) || (dst_c % 8) || (dst_c_offset % 8) ||
      dst_c_offset + c > dst_c)
    return set_error(TOK_ERR_INVALID, "bilinear_fwd: bad shape (channel counts and offsets must be multiples of 8)");
  const long long total = (long long)n * ho * wo * (c / 8);
  (void)launch_pdl(bilinear_fwd_kernel, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, 
      (const uint4*)src, (uint4*)dst, total, hi, wi, ho, wo, c / 8, dst_c / 8, dst_c_offset / 8, (float)hi / ho,
      (float)wi / wo);
  TOK_CHECK_LAUNCH("bilinear_fwd");
  return TOK_OK;
}

int tok_bilinear_bwd(int n, int hi, int wi, int c, int ho, int wo, const void* dout, int dout_c, int dout_c_offset,
                     void* dsrc, void* stream) {
  if (n <= 0 || hi <= 0 || wi <= 0 || ho <= 0 || wo <= 0 || c <= 0 || (c % 8) || (dout_c % 8) ||
      (dout_c_offset % 8) || dout_c_offset + c > dout_c)
    return set_error(TOK_ERR_INVALID, "bilinear_bwd: bad shape (channel counts and offsets must be multiples of 8)");
  const long long total = (long long)n * hi * wi * (c / 8);
  (void)launch_pdl(bilinear_bwd_kernel, dim3(grid_for(total)), dim3(256), 0, (cudaStream_t)stream, 
      (const uint4*)dout, (uint4*)dsrc, total, hi, wi, ho, wo, c / 8, dout_c / 8, dout_c_offset / 8, (float)hi / ho,
      (float)wi / wo);
  TOK_CHECK_LAUNCH("bilinear_bwd");
  return TOK_OK;
}

int tok_softmax_xent_small(long long rows, int C, int ld, const void* logits, const long long* target,
                           float* loss_sum, float* count, void* dlogits, const float* inv_count_dev, float gscale,
                           const float* gscale_dev, long long ignore_index, void* stream) {
  if (rows <= 0 || C <= 0 || C > 64 || ld < C) return set_error(TOK_ERR_INVALID, "softmax_xent_small: 1 <= C <= 64, ld >= C");
  if (!dlogits && (!loss_sum || !count)) return set_error(TOK_ERR_INVALID, "softmax_xent_small: forward needs loss_sum and count");
  const bool vec = ld % 8 == 0 && ld <= 32 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dlogits) & 15) == 0;
#define TOK_XENT_VEC(LV)                                                                                          \
  (void)launch_pdl(xent_small_vec_kernel<LV>, dim3(grid_for(rows)), dim3(256), 0, (cudaStream_t)stream,                                     \
      (const uint4*)logits, target, loss_sum, count, (uint4*)dlogits, rows, C, inv_count_dev, gscale, gscale_dev, \
      ignore_index)
  if (vec && ld == 8) TOK_XENT_VEC(1);
  else if (vec && ld == 16) TOK_XENT_VEC(2);
  else if (vec && ld == 24) TOK_XENT_VEC(3);
  else if (vec && ld == 32) TOK_XENT_VEC(4);
  else
    (void)launch_pdl(xent_small_kernel, dim3(grid_for(rows)), dim3(256), 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)logits, target, loss_sum, count, (__nv_bfloat16*)dlogits, rows, C, ld, inv_count_dev,
        gscale, gscale_dev, ignore_index);
#undef TOK_XENT_VEC
  TOK_CHECK_LAUNCH("softmax_xent_small");
  return TOK_OK;
}

int tok_dice_stats(long long rows, int C, int ld, const void* logits, const long long* target, float* stats,
                   void* stream) {
  if (rows <= 0 || C <= 0 || C > 64 || ld < C || !stats) return set_error(TOK_ERR_INVALID, "dice_stats: 1 <= C <= 64, ld >= C");
  (void)launch_pdl(dice_stats_kernel, dim3(grid_for(rows)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)logits, target, stats, rows, C, ld);
  TOK_CHECK_LAUNCH("dice_stats");
  return TOK_OK;
}

int tok_dice_finalize(int C, const float* stats, float smooth, float eps, int log_loss, float* loss, float* coef,
                      void* stream) {
  if (C <= 0 || C > 64 || !stats || !loss || !coef) return set_error(TOK_ERR_INVALID, "dice_finalize: 1 <= C <= 64");
  (void)launch_pdl(dice_finalize_kernel, dim3(1), dim3(64), 0, (cudaStream_t)stream, stats, C, smooth, eps, log_loss, loss, coef);
  TOK_CHECK_LAUNCH("dice_finalize");
  return TOK_OK;
}

int tok_dice_bwd(long long rows, int C, int ld, const void* logits, const long long* target, const float* coef,
                 const float* gscale_dev, void* dlogits, void* stream) {
  if (rows <= 0 || C <= 0 || C > 64 || ld < C || !coef || !dlogits) return set_error(TOK_ERR_INVALID, "dice_bwd: 1 <= C <= 64, ld >= C");
  (void)launch_pdl(dice_bwd_kernel, dim3(grid_for(rows)), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)logits, target, coef, gscale_dev,
                                                                  (__nv_bfloat16*)dlogits, rows, C, ld);
  TOK_CHECK_LAUNCH("dice_bwd");
  return TOK_OK;
}

}  // extern "C"
