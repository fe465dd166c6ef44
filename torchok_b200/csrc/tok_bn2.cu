// tok_bn2.cu — second-generation BatchNorm passes: minimum-traffic forward apply (+ReLU bit mask) and backward
// reduce / apply that recover the ReLU mask without re-reading the activation.
//
// Reference call sites: torch.nn.BatchNorm2d + ReLU inside ConvBnAct (torchok/models/modules/bricks/convbnact.py:44-53)
// and the timm block tails `x += shortcut; x = act(x)` (blocks built by torchok/models/backbones/resnet.py:363-405);
// autograd's BatchNorm / ReLU backward is what the *_bwd_* kernels restate.
//
// Traffic per element of a [rows][C] bf16 tensor (2 B):
//   forward  apply          : read y (+residual), write out (+1 bit when the unit is a residual tail)
//   backward reduce (pass 1): read dout, y (+1 bit)                       — was dout, out, y
//   backward apply  (pass 2): read dout, y (+1 bit), write dy (+dres)     — was dout, out, y
// ReLU mask sources: MASK_NONE (no activation), MASK_Y (plain conv->BN->ReLU: out > 0 <=> scale*y+shift > 0, recomputed
// with the forward's own fp32 scale/shift and the same fmaf), MASK_BITS (residual tails: 1 bit / element written by
// the forward pass, 8 channels per byte in vector order).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/tokb200.h"
#include "tok_internal.h"
#include "tok_ptx.cuh"
#include "tok_bnfin.cuh"

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
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u8(const uint8_t* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

constexpr int kUnroll = 4;

enum { MASK_NONE = 0, MASK_Y = 1, MASK_BITS = 2 };

// One thread owns one 8-channel vector column (cv) and walks rows rl, rl+rlanes, ... of its CTA's row range.
struct Map {
  int cl, rl, rlanes, cv;
  bool active;
  long long r0, r1;
};
__device__ __forceinline__ Map make_map(long long rows, int cvec, int cvec_b, int rows_per_cta) {
  Map m;
  m.rlanes = 256 / cvec_b;
  m.cl = threadIdx.x % cvec_b;
  m.rl = threadIdx.x / cvec_b;
  m.cv = blockIdx.y * cvec_b + m.cl;
  m.active = m.rl < m.rlanes && m.cv < cvec;
  m.r0 = (long long)blockIdx.x * rows_per_cta;
  m.r1 = m.r0 + rows_per_cta;
  if (m.r1 > rows) m.r1 = rows;
  return m;
}

// Second half of the CTA reduction of the backward sums: part[thread][16] (8 x sum g, 8 x sum g*y per 8-channel vector)
// -> one thread per (channel vector, group of four values) adds the row lanes up and issues ONE 16-byte vector
// reduction.  r1 used one scalar atomicAdd per value and ~570 CTAs per channel: with up to 4096 scalar atomics per CTA
// on the same addresses the L2 atomic units serialised them into a 13-40 us tail per launch (profiles/r2: 55 us for
// a 2048 x 7x7 tail whose traffic needs 16 us); now a channel is shared by at most ~37-148 CTAs (plan(): 16 vectors per
// column block) and every request carries four values.
__device__ __forceinline__ void cta_reduce_red4(const float* part, int cvec, int cvec_b, float* sum_g, float* sum_gy) {
  const int rlanes = 256 / cvec_b;
  for (int o = threadIdx.x; o < cvec_b * 4; o += 256) {
    const int cl = o >> 2, k0 = (o & 3) * 4;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    for (int rl = 0; rl < rlanes; ++rl) {
      const float* q = part + (rl * cvec_b + cl) * 16 + k0;
      t0 += q[0];
      t1 += q[1];
      t2 += q[2];
      t3 += q[3];
    }
    const int cv = blockIdx.y * cvec_b + cl;
    if (cv < cvec) {
      float* dst = (k0 < 8 ? sum_g : sum_gy) + cv * 8 + (k0 & 7);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t0), "f"(t1), "f"(t2), "f"(t3)
                   : "memory");
    }
  }
}

template <int MASK>
__device__ __forceinline__ void apply_mask(float (&g)[8], const float (&yy)[8], const float (&sc)[8],
                                           const float (&sf)[8], uint32_t bits) {
  if (MASK == MASK_Y) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = fmaf(yy[j], sc[j], sf[j]) > 0.f ? g[j] : 0.f;
  } else if (MASK == MASK_BITS) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = ((bits >> j) & 1u) ? g[j] : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ forward apply + bits
// out = relu(y*scale + shift + residual), bits[vector] = (out > 0) per channel.
__global__ void __launch_bounds__(256)
bn_apply_bits_kernel(const uint4* __restrict__ y, const uint4* __restrict__ res, uint4* __restrict__ out,
                     uint8_t* __restrict__ bits, const float* __restrict__ scale, const float* __restrict__ shift,
                     long long rows, int cvec, int cvec_b, int rows_per_cta, const ApplyFin fin) {
  __shared__ int s_flag;
  pdl_wait();
  pdl_launch();
  const Map m = make_map(rows, cvec, cvec_b, rows_per_cta);
  float sc[8], sf[8];
  if (fin.fused) {
    if (m.cv < cvec) applyfin_coefs(fin, m.cv * 8, sc, sf);
    applyfin_publish(fin, blockIdx.x == 0 && blockIdx.y == 0, gridDim.x * gridDim.y, &s_flag);
  }
  if (!m.active) return;
  if (!fin.fused) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(scale + m.cv * 8 + j);
      sf[j] = __ldg(shift + m.cv * 8 + j);
    }
  }
  const long long step = (long long)m.rlanes * kUnroll;
  for (long long r = m.r0 + m.rl; r < m.r1; r += step) {
    uint4 vy[kUnroll], vr[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long rr = r + (long long)u * m.rlanes;
      if (rr < m.r1) {
        vy[u] = ldg_stream(y + rr * cvec + m.cv);
        vr[u] = ldg_stream(res + rr * cvec + m.cv);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long rr = r + (long long)u * m.rlanes;
      if (rr < m.r1) {
        float f[8], q[8];
        unpack8(vy[u], f);
        unpack8(vr[u], q);
        uint32_t b = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = fmaxf(fmaf(f[j], sc[j], sf[j]) + q[j], 0.f);
          b |= (f[j] > 0.f ? 1u : 0u) << j;
        }
        out[rr * cvec + m.cv] = pack8(f);
        bits[rr * cvec + m.cv] = (uint8_t)b;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward pass 1
// Optional fused finalize: the LAST CTA to finish (ticket counter in global memory, handed back zeroed) turns the
// per-channel sums into the coefficients of pass 2 and the gamma / beta gradients, so the 4-microsecond single-CTA
// bn_bwd_finalize launch (and its launch gap) between the two passes disappears: 53 launches per ResNet-50 step.
struct BwdFin {
  unsigned* counter;   // nullptr: no fused finalize
  float count;
  const float* mean;
  const float* invstd;
  const float* gamma;
  float* coef_a;
  float* coef_c1;
  float* coef_c0;
  float* dgamma;
  float* dbeta;
  int accumulate;
  int C;
  int Cv;   // channels that exist in gamma / dgamma / dbeta (C rounded DOWN from the padded pitch); pad lanes: gamma = 0
};

template <int MASK, bool HAS2>
__global__ void __launch_bounds__(256)
bn_bwd_reduce2_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ dout2, const uint4* __restrict__ y,
                      const uint8_t* __restrict__ bits, const float* __restrict__ scale,
                      const float* __restrict__ shift, float* __restrict__ sum_g, float* __restrict__ sum_gy,
                      long long rows, int cvec, int cvec_b, int rows_per_cta, const BwdFin fin) {
  __shared__ float part[256 * 16];
  __shared__ int s_last;
  pdl_wait();
  pdl_launch();
  const Map m = make_map(rows, cvec, cvec_b, rows_per_cta);
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  if (m.active) {
    float sc[8], sf[8];
    if (MASK == MASK_Y) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sc[j] = __ldg(scale + m.cv * 8 + j);
        sf[j] = __ldg(shift + m.cv * 8 + j);
      }
    }
    const long long step = (long long)m.rlanes * kUnroll;
    for (long long r = m.r0 + m.rl; r < m.r1; r += step) {
      uint4 vg[kUnroll], vy[kUnroll], v2[kUnroll];
      uint32_t vb[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long rr = r + (long long)u * m.rlanes;
        const long long idx = rr * cvec + m.cv;
        if (rr < m.r1) {
          vg[u] = ldg_stream(dout + idx);
          vy[u] = ldg_stream(y + idx);
          if (HAS2) v2[u] = ldg_stream(dout2 + idx);
          if (MASK == MASK_BITS) vb[u] = ldg_stream_u8(bits + idx);
        } else {
          vg[u] = make_uint4(0, 0, 0, 0);
          vy[u] = make_uint4(0, 0, 0, 0);
          if (HAS2) v2[u] = make_uint4(0, 0, 0, 0);
          if (MASK == MASK_BITS) vb[u] = 0;
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        float g[8], yy[8];
        unpack8(vg[u], g);
        unpack8(vy[u], yy);
        if (HAS2) {
          float g2[8];
          unpack8(v2[u], g2);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += g2[j];
        }
        apply_mask<MASK>(g, yy, sc, sf, MASK == MASK_BITS ? vb[u] : 0u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a1[j] += g[j];
          a2[j] = fmaf(g[j], yy[j], a2[j]);
        }
      }
    }
  }
  // CTA reduction without shared-memory atomics: part[thread][16], then one thread per (channel-vector, slot)
  float* mine = part + threadIdx.x * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mine[j] = a1[j];
    mine[8 + j] = a2[j];
  }
  __syncthreads();
  cta_reduce_red4(part, cvec, cvec_b, sum_g, sum_gy);
  if (fin.counter != nullptr) {
    __threadfence();   // this CTA's contributions are ordered before its ticket
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(fin.counter, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      for (int c = threadIdx.x; c < fin.C; c += 256) {
        const float sg = __ldcg(sum_g + c);
        const float mu = fin.mean[c];
        const float is = fin.invstd[c];
        const float sgx = is * (__ldcg(sum_gy + c) - mu * sg);
        sum_g[c] = 0.f;   // consumed: handed back zeroed (see bn_bwd_finalize_kernel)
        sum_gy[c] = 0.f;
        const bool real = c < fin.Cv;
        const float g = real ? (fin.gamma ? fin.gamma[c] : 1.f) : 0.f;
        const float a = g * is;
        const float k1 = sg / fin.count;
        const float k2 = sgx / fin.count;
        fin.coef_a[c] = a;
        fin.coef_c1[c] = -a * k2 * is;
        fin.coef_c0[c] = -a * k1 + a * k2 * is * mu;
        if (fin.dgamma && real) fin.dgamma[c] = fin.accumulate ? fin.dgamma[c] + sgx : sgx;
        if (fin.dbeta && real) fin.dbeta[c] = fin.accumulate ? fin.dbeta[c] + sg : sg;
      }
      if (threadIdx.x == 0) *fin.counter = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward pass 2
// dy = a*g + c1*y + c0 ; dres (optional) = g
template <int MASK, bool HAS2, bool DRES>
__global__ void __launch_bounds__(256)
bn_bwd_apply2_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ dout2, const uint4* __restrict__ y,
                     const uint8_t* __restrict__ bits, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ coef_a,
                     const float* __restrict__ coef_c1, const float* __restrict__ coef_c0, uint4* __restrict__ dy,
                     uint4* __restrict__ dres, long long rows, int cvec, int cvec_b, int rows_per_cta) {
  pdl_wait();
  pdl_launch();
  const Map m = make_map(rows, cvec, cvec_b, rows_per_cta);
  if (!m.active) return;
  float sc[8], sf[8], ca[8], c1[8], c0[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (MASK == MASK_Y) {
      sc[j] = __ldg(scale + m.cv * 8 + j);
      sf[j] = __ldg(shift + m.cv * 8 + j);
    }
    ca[j] = __ldg(coef_a + m.cv * 8 + j);
    c1[j] = __ldg(coef_c1 + m.cv * 8 + j);
    c0[j] = __ldg(coef_c0 + m.cv * 8 + j);
  }
  const long long step = (long long)m.rlanes * kUnroll;
  for (long long r = m.r0 + m.rl; r < m.r1; r += step) {
    uint4 vg[kUnroll], vy[kUnroll], v2[kUnroll];
    uint32_t vb[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long rr = r + (long long)u * m.rlanes;
      const long long idx = rr * cvec + m.cv;
      if (rr < m.r1) {
        vg[u] = ldg_stream(dout + idx);
        vy[u] = ldg_stream(y + idx);
        if (HAS2) v2[u] = ldg_stream(dout2 + idx);
        if (MASK == MASK_BITS) vb[u] = ldg_stream_u8(bits + idx);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long rr = r + (long long)u * m.rlanes;
      if (rr < m.r1) {
        const long long idx = rr * cvec + m.cv;
        float g[8], yy[8];
        unpack8(vg[u], g);
        unpack8(vy[u], yy);
        if (HAS2) {
          float g2[8];
          unpack8(v2[u], g2);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += g2[j];
        }
        apply_mask<MASK>(g, yy, sc, sf, MASK == MASK_BITS ? vb[u] : 0u);
        if (DRES) dres[idx] = pack8(g);
#pragma unroll
        for (int j = 0; j < 8; ++j) yy[j] = fmaf(ca[j], g[j], fmaf(c1[j], yy[j], c0[j]));
        dy[idx] = pack8(yy);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward, ONE launch
// Pass 1 + finalize + pass 2 in one kernel for tensors that fit the L2 (layer3 / layer4 of ResNet-50 at bs256: g and y
// are 26-100 MB): every CTA reduces its rows, the last one through the ticket turns the sums into the coefficients and
// flips a release word; the others wait for the flip (bounded spin) and then re-read THE SAME rows — from L2 now — for
// dy = a g + c1 y + c0.  Against two launches this removes the second HBM read of g and y, one launch and the gap
// between them; separate launches measured ~20 us each on these tensors whatever their size (ramp + atomics -> fence ->
// ticket -> finalize tail), i.e. ~1.9 ms of the 8.4 ms of BatchNorm time of a step.
// The grid is ONE wave of co-resident CTAs by construction (launcher: occupancy API), so the wait cannot deadlock on
// its own grid; CTAs of other streams only delay it.  flag protocol: every CTA samples the release word BEFORE it takes
// its ticket; the word changes only after ALL tickets are taken, so all CTAs sample the same value and wait for the
// other one.
template <int MASK, bool HAS2, bool DRES>
__global__ void __launch_bounds__(256)
bn_bwd_fused_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ dout2, const uint4* __restrict__ y,
                    const uint8_t* __restrict__ bits, const float* __restrict__ scale, const float* __restrict__ shift,
                    float* __restrict__ sum_g, float* __restrict__ sum_gy, uint4* __restrict__ dy,
                    uint4* __restrict__ dres, long long rows, int cvec, int cvec_b, int rows_per_cta, const BwdFin fin,
                    unsigned* release) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float part[256 * 16];
  __shared__ int s_last;
  __shared__ unsigned s_seen;
  const Map m = make_map(rows, cvec, cvec_b, rows_per_cta);
  if (threadIdx.x == 0) s_seen = *reinterpret_cast<volatile unsigned*>(release);
  float sc[8], sf[8];
  if (MASK == MASK_Y && m.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(scale + m.cv * 8 + j);
      sf[j] = __ldg(shift + m.cv * 8 + j);
    }
  }
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  const long long step = (long long)m.rlanes * kUnroll;
  if (m.active) {
    for (long long r = m.r0 + m.rl; r < m.r1; r += step) {
      uint4 vg[kUnroll], vy[kUnroll], v2[kUnroll];
      uint32_t vb[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long rr = r + (long long)u * m.rlanes;
        const long long idx = rr * cvec + m.cv;
        if (rr < m.r1) {
          vg[u] = ldg_stream(dout + idx);
          vy[u] = ldg_stream(y + idx);
          if (HAS2) v2[u] = ldg_stream(dout2 + idx);
          if (MASK == MASK_BITS) vb[u] = ldg_stream_u8(bits + idx);
        } else {
          vg[u] = make_uint4(0, 0, 0, 0);
          vy[u] = make_uint4(0, 0, 0, 0);
          if (HAS2) v2[u] = make_uint4(0, 0, 0, 0);
          if (MASK == MASK_BITS) vb[u] = 0;
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        float g[8], yy[8];
        unpack8(vg[u], g);
        unpack8(vy[u], yy);
        if (HAS2) {
          float g2[8];
          unpack8(v2[u], g2);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += g2[j];
        }
        apply_mask<MASK>(g, yy, sc, sf, MASK == MASK_BITS ? vb[u] : 0u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a1[j] += g[j];
          a2[j] = fmaf(g[j], yy[j], a2[j]);
        }
      }
    }
  }
  float* mine = part + threadIdx.x * 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mine[j] = a1[j];
    mine[8 + j] = a2[j];
  }
  __syncthreads();
  cta_reduce_red4(part, cvec, cvec_b, sum_g, sum_gy);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(fin.counter, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  const unsigned seen = s_seen;
  if (s_last) {
    __threadfence();
    for (int c = threadIdx.x; c < fin.C; c += 256) {
      const float sg = __ldcg(sum_g + c);
      const float mu = fin.mean[c];
      const float is = fin.invstd[c];
      const float sgx = is * (__ldcg(sum_gy + c) - mu * sg);
      sum_g[c] = 0.f;
      sum_gy[c] = 0.f;
      const bool real = c < fin.Cv;
      const float g = real ? (fin.gamma ? fin.gamma[c] : 1.f) : 0.f;
      const float a = g * is;
      const float k1 = sg / fin.count;
      const float k2 = sgx / fin.count;
      fin.coef_a[c] = a;
      fin.coef_c1[c] = -a * k2 * is;
      fin.coef_c0[c] = -a * k1 + a * k2 * is * mu;
      if (fin.dgamma && real) fin.dgamma[c] = fin.accumulate ? fin.dgamma[c] + sgx : sgx;
      if (fin.dbeta && real) fin.dbeta[c] = fin.accumulate ? fin.dbeta[c] + sg : sg;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      *fin.counter = 0u;
      __threadfence();
      atomicExch(release, seen ^ 1u);
    }
  } else if (threadIdx.x == 0) {
    // bounded wait for the release flip (~4 s): a scheduling surprise must end as a launch failure, not a hung GPU
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned*>(release) == seen) {
      if (clock64() - t0 > 8000000000LL) {
        printf("tokb200: bn_bwd_fused release wait timed out (block %d,%d)\n", blockIdx.x, blockIdx.y);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
  if (!m.active) return;
  // ---- pass 2 on the same rows (L2 hits): coefficients were written by another SM — read them past L1
  float ca[8], c1[8], c0[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ca[j] = __ldcg(fin.coef_a + m.cv * 8 + j);
    c1[j] = __ldcg(fin.coef_c1 + m.cv * 8 + j);
    c0[j] = __ldcg(fin.coef_c0 + m.cv * 8 + j);
  }
  for (long long r = m.r0 + m.rl; r < m.r1; r += step) {
    uint4 vg[kUnroll], vy[kUnroll], v2[kUnroll];
    uint32_t vb[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long rr = r + (long long)u * m.rlanes;
      const long long idx = rr * cvec + m.cv;
      if (rr < m.r1) {
        vg[u] = ldg_stream(dout + idx);
        vy[u] = ldg_stream(y + idx);
        if (HAS2) v2[u] = ldg_stream(dout2 + idx);
        if (MASK == MASK_BITS) vb[u] = ldg_stream_u8(bits + idx);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const long long rr = r + (long long)u * m.rlanes;
      if (rr < m.r1) {
        const long long idx = rr * cvec + m.cv;
        float g[8], yy[8];
        unpack8(vg[u], g);
        unpack8(vy[u], yy);
        if (HAS2) {
          float g2[8];
          unpack8(v2[u], g2);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += g2[j];
        }
        apply_mask<MASK>(g, yy, sc, sf, MASK == MASK_BITS ? vb[u] : 0u);
        if (DRES) dres[idx] = pack8(g);
#pragma unroll
        for (int j = 0; j < 8; ++j) yy[j] = fmaf(ca[j], g[j], fmaf(c1[j], yy[j], c0[j]));
        dy[idx] = pack8(yy);
      }
    }
  }
}

// dst[n, p*s, q*s, :] += src[n, p, q, :]  — merges the compact data gradient of a strided 1x1 (downsample) conv
// into the full-resolution gradient of the block input (replaces a zero-filled scatter + full-size addend read)
__global__ void __launch_bounds__(256)
strided_add_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long total, int P, int Q, int cvec,
                   int H, int W, int s) {
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
    const long long o = ((n * H + (long long)p * s) * W + (long long)q * s) * cvec + cv;
    float a[8], b[8];
    unpack8(ldg_stream(src + i), a);
    unpack8(dst[o], b);
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] += a[j];
    dst[o] = pack8(b);
  }
}

// ------------------------------------------------------------------------------------------------ fused ResNet stem
// conv7x7 output y -> BN -> ReLU -> maxpool 3x3/2 pad 1 in ONE pass (torchok/models/backbones/resnet.py:488-490,510,
// 542-545): reads y once, writes the pooled tensor + argmax slots (and `act` only when the caller needs the act1
// feature); the backward below gathers the pooled gradient on the fly, so neither the 112x112 activation nor its
// gradient ever makes a round trip through HBM.
__global__ void __launch_bounds__(256)
stem_bn_relu_pool_fwd_kernel(const uint4* __restrict__ y, const float* __restrict__ scale,
                             const float* __restrict__ shift, uint4* __restrict__ act, uint4* __restrict__ pooled,
                             uint2* __restrict__ arg, long long total, int H, int W, int P, int Q, int cvec) {
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
    float sc[8], sf[8], best[8];
    uint32_t bi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(scale + cv * 8 + j);
      sf[j] = __ldg(shift + cv * 8 + j);
      best[j] = -INFINITY;
      bi[j] = 0;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = p * 2 - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int w = q * 2 - 1 + c;
        if (w < 0 || w >= W) continue;
        const long long o = ((n * H + h) * W + w) * cvec + cv;
        float f[8];
        unpack8(__ldg(y + o), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = fmaxf(fmaf(f[j], sc[j], sf[j]), 0.f);
          f[j] = __bfloat162float(__float2bfloat16(f[j]));  // the value torch's pool would see (bf16 activation)
          if (f[j] > best[j]) {
            best[j] = f[j];
            bi[j] = r * 3 + c;
          }
        }
        if (act != nullptr && r >= 1 && c >= 1) act[o] = pack8(f);  // rows 2p, 2p+1 / cols 2q, 2q+1: unique owner
      }
    }
    pooled[i] = pack8(best);
    arg[i] = make_uint2(bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24),
                        bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24));
  }
}

// gradient reaching y's ReLU output at (n, h, w, cv): sum over the <= 4 pooling windows that contain the pixel of
// dpooled * [argmax slot == this pixel]  (+ the direct gradient of the act1 feature when it was used).
// Branch-free: window A = (h >> 1) always contains the row (tap 1 or 2), window B = A + 1 contains it (tap 0) iff h is
// odd; same along w.  All eight loads are issued before any use.
__device__ __forceinline__ void stem_pool_grad(float (&g)[8], const uint4* __restrict__ dpooled,
                                               const uint2* __restrict__ arg, const uint4* __restrict__ dact,
                                               long long n, int h, int w, int cv, int H, int W, int P, int Q, int cvec) {
  const int pA = h >> 1, qA = w >> 1;
  const int rA = h - (pA * 2 - 1), cA = w - (qA * 2 - 1);
  const bool hB = (h & 1) && (pA + 1 < P), wB = (w & 1) && (qA + 1 < Q);
  const int pB = hB ? pA + 1 : pA, qB = wB ? qA + 1 : qA;
  const long long base = n * P;
  const long long o00 = ((base + pA) * Q + qA) * cvec + cv, o01 = ((base + pA) * Q + qB) * cvec + cv;
  const long long o10 = ((base + pB) * Q + qA) * cvec + cv, o11 = ((base + pB) * Q + qB) * cvec + cv;
  const uint2 a00 = __ldg(arg + o00), a01 = __ldg(arg + o01), a10 = __ldg(arg + o10), a11 = __ldg(arg + o11);
  const uint4 d00 = __ldg(dpooled + o00), d01 = __ldg(dpooled + o01), d10 = __ldg(dpooled + o10),
              d11 = __ldg(dpooled + o11);
#pragma unroll
  for (int j = 0; j < 8; ++j) g[j] = 0.f;
  if (dact != nullptr) unpack8(ldg_stream(dact + ((n * H + h) * W + w) * cvec + cv), g);
  const uint32_t s00 = rA * 3 + cA, s01 = rA * 3 + 0, s10 = 0 * 3 + cA, s11 = 0;
  auto add = [&](const uint2& a, const uint4& dv, uint32_t slot, bool on) {
    float d[8];
    unpack8(dv, d);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (on && ((a.x >> (8 * j)) & 0xFF) == slot) g[j] += d[j];
      if (on && ((a.y >> (8 * j)) & 0xFF) == slot) g[4 + j] += d[4 + j];
    }
  };
  add(a00, d00, s00, true);
  add(a01, d01, s01, wB);
  add(a10, d10, s10, hB);
  add(a11, d11, s11, hB && wB);
}

// mode 0: sum_g / sum_gy reduction; mode 1: dy = a*g + c1*y + c0
template <int MODE>
__global__ void __launch_bounds__(256)
stem_bwd_kernel(const uint4* __restrict__ dpooled, const uint2* __restrict__ arg, const uint4* __restrict__ dact,
                const uint4* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                float* __restrict__ sum_g, float* __restrict__ sum_gy, const float* __restrict__ coef_a,
                const float* __restrict__ coef_c1, const float* __restrict__ coef_c0, uint4* __restrict__ dy,
                long long rows, int H, int W, int P, int Q, int cvec, int cvec_b, int rows_per_cta) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float part[MODE == 0 ? 256 * 16 : 1];
  const Map m = make_map(rows, cvec, cvec_b, rows_per_cta);
  float a1[8], a2[8], sc[8], sf[8], ca[8], c1[8], c0[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  if (m.active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(scale + m.cv * 8 + j);
      sf[j] = __ldg(shift + m.cv * 8 + j);
      if (MODE == 1) {
        ca[j] = __ldg(coef_a + m.cv * 8 + j);
        c1[j] = __ldg(coef_c1 + m.cv * 8 + j);
        c0[j] = __ldg(coef_c0 + m.cv * 8 + j);
      }
    }
    for (long long r = m.r0 + m.rl; r < m.r1; r += m.rlanes) {
      const int w = (int)(r % W);
      const long long t = r / W;
      const int h = (int)(t % H);
      const long long n = t / H;
      float g[8], yy[8];
      unpack8(ldg_stream(y + r * cvec + m.cv), yy);
      stem_pool_grad(g, dpooled, arg, dact, n, h, w, m.cv, H, W, P, Q, cvec);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = fmaf(yy[j], sc[j], sf[j]) > 0.f ? g[j] : 0.f;
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a1[j] += g[j];
          a2[j] = fmaf(g[j], yy[j], a2[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) yy[j] = fmaf(ca[j], g[j], fmaf(c1[j], yy[j], c0[j]));
        dy[r * cvec + m.cv] = pack8(yy);
      }
    }
  }
  if (MODE == 0) {
    float* mine = part + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[j] = a1[j];
      mine[8 + j] = a2[j];
    }
    __syncthreads();
    cta_reduce_red4(part, cvec, cvec_b, sum_g, sum_gy);
  }
}

// Second version of the stem backward: a thread owns the 2x2 block of conv-output pixels {2p, 2p+1} x {2q, 2q+1} of one
// 8-channel vector.  The four pooling windows {p, p+1} x {q, q+1} are exactly the windows that can have picked one of
// those pixels, so 4 argmax words + 4 pooled gradients + 4 y vectors serve 32 outputs (v1: 9 loads and two 64-bit
// div/mod chains PER PIXEL; 437 + 463 us for the 205 M element tensor whose traffic needs ~90 + ~130 us).
// Requires even H and W (the fused 3/2/1 pooling path) and cvec | 256.
template <int MODE>
__global__ void __launch_bounds__(256, 2)
stem_bwd2_kernel(const uint4* __restrict__ dpooled, const uint2* __restrict__ arg, const uint4* __restrict__ dact,
                 const uint4* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                 float* __restrict__ sum_g, float* __restrict__ sum_gy, const float* __restrict__ coef_a,
                 const float* __restrict__ coef_c1, const float* __restrict__ coef_c0, uint4* __restrict__ dy, int total,
                 int H, int W, int P, int Q, int cvec) {
  pdl_wait();   // programmatic dependent launch: see tok_ptx.cuh
  pdl_launch();
  __shared__ float part[MODE == 0 ? 256 * 16 : 1];
  const int cv = threadIdx.x % cvec;   // constant per thread: 256 and the grid stride are multiples of cvec
  float a1[8], a2[8], sc[8], sf[8], ca[8], c1[8], c0[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a1[j] = a2[j] = 0.f;
    sc[j] = __ldg(scale + cv * 8 + j);
    sf[j] = __ldg(shift + cv * 8 + j);
    if (MODE == 1) {
      ca[j] = __ldg(coef_a + cv * 8 + j);
      c1[j] = __ldg(coef_c1 + cv * 8 + j);
      c0[j] = __ldg(coef_c0 + cv * 8 + j);
    }
  }
  const int H2 = H >> 1, W2 = W >> 1;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    int t = i / cvec;
    const int qA = t % W2;
    t /= W2;
    const int pA = t % H2;
    const int n = t / H2;
    const bool hB = pA + 1 < P, wB = qA + 1 < Q;
    const int pB = hB ? pA + 1 : pA, qB = wB ? qA + 1 : qA;
    const long long wbase = (long long)n * P;
    const long long o00 = ((wbase + pA) * Q + qA) * cvec + cv, o01 = ((wbase + pA) * Q + qB) * cvec + cv;
    const long long o10 = ((wbase + pB) * Q + qA) * cvec + cv, o11 = ((wbase + pB) * Q + qB) * cvec + cv;
    const long long y00 = (((long long)n * H + 2 * pA) * W + 2 * qA) * cvec + cv;
    const long long yo[4] = {y00, y00 + cvec, y00 + (long long)W * cvec, y00 + (long long)W * cvec + cvec};
    const uint2 a00 = __ldg(arg + o00), a01 = __ldg(arg + o01), a10 = __ldg(arg + o10), a11 = __ldg(arg + o11);
    const uint4 d00 = __ldg(dpooled + o00), d01 = __ldg(dpooled + o01), d10 = __ldg(dpooled + o10),
                d11 = __ldg(dpooled + o11);
    uint4 vy[4], va[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      vy[k] = ldg_stream(y + yo[k]);
      if (dact != nullptr) va[k] = ldg_stream(dact + yo[k]);
    }
    float p00[8], p01[8], p10[8], p11[8];
    unpack8(d00, p00);
    unpack8(d01, p01);
    unpack8(d10, p10);
    unpack8(d11, p11);
    // pooling slot (3 * r + c) of pixel k = (dh, dw) inside each window that contains it
    //   window (pA, qA): r = 1 + dh, c = 1 + dw        window (pA, qB): r = 1 + dh, c = 0   (odd w only)
    //   window (pB, qA): r = 0, c = 1 + dw (odd h only)  window (pB, qB): slot 0             (odd h and odd w)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int dh = k >> 1, dw = k & 1;
      float g[8], yy[8];
      unpack8(vy[k], yy);
      if (dact != nullptr) unpack8(va[k], g);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
      }
      const uint32_t s00 = (1 + dh) * 3 + (1 + dw), s01 = (1 + dh) * 3, s10 = 1 + dw, s11 = 0;
      const bool on01 = dw == 1 && wB, on10 = dh == 1 && hB, on11 = on01 && on10;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t w00 = ((j < 4 ? a00.x : a00.y) >> (8 * (j & 3))) & 0xFF;
        const uint32_t w01 = ((j < 4 ? a01.x : a01.y) >> (8 * (j & 3))) & 0xFF;
        const uint32_t w10 = ((j < 4 ? a10.x : a10.y) >> (8 * (j & 3))) & 0xFF;
        const uint32_t w11 = ((j < 4 ? a11.x : a11.y) >> (8 * (j & 3))) & 0xFF;
        float v = g[j];
        v += (w00 == s00) ? p00[j] : 0.f;
        v += (on01 && w01 == s01) ? p01[j] : 0.f;
        v += (on10 && w10 == s10) ? p10[j] : 0.f;
        v += (on11 && w11 == s11) ? p11[j] : 0.f;
        g[j] = fmaf(yy[j], sc[j], sf[j]) > 0.f ? v : 0.f;
      }
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a1[j] += g[j];
          a2[j] = fmaf(g[j], yy[j], a2[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) yy[j] = fmaf(ca[j], g[j], fmaf(c1[j], yy[j], c0[j]));
        dy[yo[k]] = pack8(yy);
      }
    }
  }
  if (MODE == 0) {
    float* mine = part + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[j] = a1[j];
      mine[8 + j] = a2[j];
    }
    __syncthreads();
    cta_reduce_red4(part, cvec, cvec, sum_g, sum_gy);   // blockIdx.y == 0: column block 0 covers all cvec vectors
  }
}

struct Grid2 {
  dim3 grid;
  int cvec, cvec_b, rows_per_cta;
};
Grid2 plan(long long rows, int C, int ctas_per_sm, int max_cvec_b = 256) {
  Grid2 g;
  g.cvec = C / 8;
  g.cvec_b = g.cvec < max_cvec_b ? g.cvec : max_cvec_b;
  const int gy = (g.cvec + g.cvec_b - 1) / g.cvec_b;
  const int rlanes = 256 / g.cvec_b;
  long long ctas = 148LL * ctas_per_sm / gy;
  if (ctas < 1) ctas = 1;
  long long rpc = (rows + ctas - 1) / ctas;
  const long long quantum = (long long)rlanes * kUnroll;
  rpc = (rpc + quantum - 1) / quantum * quantum;  // whole unrolled iterations: no ragged tail inside a CTA
  ctas = (rows + rpc - 1) / rpc;
  g.grid = dim3((unsigned)ctas, gy);
  g.rows_per_cta = (int)rpc;
  return g;
}

}  // namespace
}  // namespace tok

using namespace tok;

extern "C" {

static int launch_apply_bits(long long rows, int C, const void* y, const float* scale, const float* shift,
                             const void* residual, void* out, void* bits, const ApplyFin& fin, void* stream) {
  if (C <= 0 || (C % 8)) return set_error(TOK_ERR_INVALID, "bn_apply_bits: C must be a positive multiple of 8 (got %d)", C);
  if (rows <= 0 || !residual || !bits) return set_error(TOK_ERR_INVALID, "bn_apply_bits: rows, residual and bits are required");
  const Grid2 g = plan(rows, C, 6);
  cudaError_t le = launch_pdl(bn_apply_bits_kernel, g.grid, dim3(256), 0, (cudaStream_t)stream, (const uint4*)y,
                              (const uint4*)residual, (uint4*)out, (uint8_t*)bits, scale, shift, rows, g.cvec, g.cvec_b,
                              g.rows_per_cta, fin);
  if (le != cudaSuccess) return set_error(TOK_ERR_CUDA, "bn_apply_bits: %s", cudaGetErrorString(le));
  return TOK_OK;
}

int tok_bn_apply_bits(long long rows, int C, const void* y, const float* scale, const float* shift,
                      const void* residual, void* out, void* bits, void* stream) {
  ApplyFin fin;
  memset(&fin, 0, sizeof(fin));
  return launch_apply_bits(rows, C, y, scale, shift, residual, out, bits, fin, stream);
}

int tok_bn_apply_bits_train(long long rows, int C, const void* y, float* sum, float* sqsum, const float* gamma,
                            const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                            float* scale, float* shift, float* save_mean, float* save_invstd, unsigned* counter,
                            const void* residual, void* out, void* bits, void* stream) {
  if (!sum || !sqsum || !scale || !shift || !save_mean || !save_invstd || !counter)
    return set_error(TOK_ERR_INVALID, "bn_apply_bits_train: accumulators, outputs and the ticket counter are required");
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
  return launch_apply_bits(rows, C, y, scale, shift, residual, out, bits, fin, stream);
}

int tok_bn_apply_bits_chain(long long rows, int C, int c_valid, const void* y, float* sum, float* sqsum,
                            const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                            float* running_var, float* scale, float* shift, float* save_mean, float* save_invstd,
                            float* zero_ptr, int zero_n, const void* residual, void* out, void* bits, void* stream) {
  if (!sum || !sqsum || !scale || !shift || !save_mean || !save_invstd || c_valid <= 0 || c_valid > C || zero_n < 0 ||
      (zero_n > 0 && !zero_ptr))
    return set_error(TOK_ERR_INVALID, "bn_apply_bits_chain: accumulators and outputs are required");
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
  return launch_apply_bits(rows, C, y, scale, shift, residual, out, bits, fin, stream);
}

int tok_strided_add(int n, int h, int w, int c, int stride, const void* src_compact, void* dst, void* stream) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8) || stride <= 0)
    return set_error(TOK_ERR_INVALID, "strided_add: bad shape (c must be a positive multiple of 8)");
  const int P = (h - 1) / stride + 1, Q = (w - 1) / stride + 1;
  const long long total = (long long)n * P * Q * (c / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  (void)launch_pdl(strided_add_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const uint4*)src_compact, (uint4*)dst, total,
                                                                       P, Q, c / 8, h, w, stride);
  TOK_CHECK_LAUNCH("strided_add");
  return TOK_OK;
}

int tok_stem_bn_relu_pool_fwd(int n, int h, int w, int c, const void* y, const float* scale, const float* shift,
                              void* act, void* pooled, void* argmax, void* stream) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8)) return set_error(TOK_ERR_INVALID, "stem_bn_relu_pool_fwd: bad shape");
  const int P = (h + 2 - 3) / 2 + 1, Q = (w + 2 - 3) / 2 + 1;
  if (act && ((h & 1) || (w & 1))) return set_error(TOK_ERR_INVALID, "stem_bn_relu_pool_fwd: act output needs even H, W");
  const long long total = (long long)n * P * Q * (c / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  (void)launch_pdl(stem_bn_relu_pool_fwd_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, 
      (const uint4*)y, scale, shift, (uint4*)act, (uint4*)pooled, (uint2*)argmax, total, h, w, P, Q, c / 8);
  TOK_CHECK_LAUNCH("stem_bn_relu_pool_fwd");
  return TOK_OK;
}

int tok_stem_bwd_reduce(int n, int h, int w, int c, const void* dpooled, const void* argmax, const void* dact,
                        const void* y, const float* scale, const float* shift, float* sum_g, float* sum_gy,
                        void* stream) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8)) return set_error(TOK_ERR_INVALID, "stem_bwd_reduce: bad shape");
  const int P = (h + 2 - 3) / 2 + 1, Q = (w + 2 - 3) / 2 + 1;
  const long long rows = (long long)n * h * w;
  if (!(h & 1) && !(w & 1) && 256 % (c / 8) == 0 && !getenv("TOK_STEM_BWD_V1")) {
    const int total = n * (h / 2) * (w / 2) * (c / 8);
    static int occ = 0;
    if (occ == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stem_bwd2_kernel<0>, 256, 0) != cudaSuccess || occ < 1))
      occ = 2;
    int grid = 148 * occ;
    if (grid > (total + 255) / 256) grid = (total + 255) / 256;
    (void)launch_pdl(stem_bwd2_kernel<0>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
        (const uint4*)dpooled, (const uint2*)argmax, (const uint4*)dact, (const uint4*)y, scale, shift, sum_g, sum_gy,
        nullptr, nullptr, nullptr, nullptr, total, h, w, P, Q, c / 8);
    TOK_CHECK_LAUNCH("stem_bwd_reduce");
    return TOK_OK;
  }
  const Grid2 g = plan(rows, c, 6);
  (void)launch_pdl(stem_bwd_kernel<0>, dim3(g.grid), dim3(256), 0, (cudaStream_t)stream, 
      (const uint4*)dpooled, (const uint2*)argmax, (const uint4*)dact, (const uint4*)y, scale, shift, sum_g, sum_gy,
      nullptr, nullptr, nullptr, nullptr, rows, h, w, P, Q, g.cvec, g.cvec_b, g.rows_per_cta);
  TOK_CHECK_LAUNCH("stem_bwd_reduce");
  return TOK_OK;
}

int tok_stem_bwd_apply(int n, int h, int w, int c, const void* dpooled, const void* argmax, const void* dact,
                       const void* y, const float* scale, const float* shift, const float* coef_a,
                       const float* coef_c1, const float* coef_c0, void* dy, void* stream) {
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || (c % 8)) return set_error(TOK_ERR_INVALID, "stem_bwd_apply: bad shape");
  const int P = (h + 2 - 3) / 2 + 1, Q = (w + 2 - 3) / 2 + 1;
  const long long rows = (long long)n * h * w;
  if (!(h & 1) && !(w & 1) && 256 % (c / 8) == 0 && !getenv("TOK_STEM_BWD_V1")) {
    const int total = n * (h / 2) * (w / 2) * (c / 8);
    static int occ = 0;
    if (occ == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stem_bwd2_kernel<1>, 256, 0) != cudaSuccess || occ < 1))
      occ = 2;
    int grid = 148 * occ * 2;
    if (grid > (total + 255) / 256) grid = (total + 255) / 256;
    (void)launch_pdl(stem_bwd2_kernel<1>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, 
        (const uint4*)dpooled, (const uint2*)argmax, (const uint4*)dact, (const uint4*)y, scale, shift, nullptr, nullptr,
        coef_a, coef_c1, coef_c0, (uint4*)dy, total, h, w, P, Q, c / 8);
    TOK_CHECK_LAUNCH("stem_bwd_apply");
    return TOK_OK;
  }
  const Grid2 g = plan(rows, c, 6);
  (void)launch_pdl(stem_bwd_kernel<1>, dim3(g.grid), dim3(256), 0, (cudaStream_t)stream, 
      (const uint4*)dpooled, (const uint2*)argmax, (const uint4*)dact, (const uint4*)y, scale, shift, nullptr, nullptr,
      coef_a, coef_c1, coef_c0, (uint4*)dy, rows, h, w, P, Q, g.cvec, g.cvec_b, g.rows_per_cta);
  TOK_CHECK_LAUNCH("stem_bwd_apply");
  return TOK_OK;
}

#define TOK_BN2_DISPATCH_MASK(KERNEL, ...)                                                    \
  do {                                                                                        \
    if (mask_mode == MASK_NONE) { KERNEL(MASK_NONE, __VA_ARGS__); }                           \
    else if (mask_mode == MASK_Y) { KERNEL(MASK_Y, __VA_ARGS__); }                            \
    else { KERNEL(MASK_BITS, __VA_ARGS__); }                                                  \
  } while (0)

// column-block width (16-byte vectors) of the backward reduction; TOK_BN_REDUCE_CVB overrides (tuning aid)
static int reduce_cvb(int C) {
  static const int forced = getenv("TOK_BN_REDUCE_CVB") ? atoi(getenv("TOK_BN_REDUCE_CVB")) : 0;
  (void)C;
  return forced > 0 ? forced : 16;
}

static int launch_bwd_reduce2(long long rows, int C, const void* dout, const void* dout2, const void* y, int mask_mode,
                              const void* bits, const float* scale, const float* shift, float* sum_g, float* sum_gy,
                              const BwdFin& fin, void* stream) {
  if (C <= 0 || (C % 8)) return set_error(TOK_ERR_INVALID, "bn_bwd_reduce2: C must be a positive multiple of 8 (got %d)", C);
  if (rows <= 0) return set_error(TOK_ERR_INVALID, "bn_bwd_reduce2: no rows");
  if (mask_mode < 0 || mask_mode > 2 || (mask_mode == MASK_BITS && !bits) || (mask_mode == MASK_Y && (!scale || !shift)))
    return set_error(TOK_ERR_INVALID, "bn_bwd_reduce2: mask_mode %d needs its operands", mask_mode);
  cudaStream_t st = (cudaStream_t)stream;
  // The grid is ONE wave of the CTAs that are really resident (occupancy API, per instantiation).  ncu r2: planned for 4
  // CTAs per SM while 77 registers admit 3, the 523-CTA grid ran as 1.18 waves and its 79-CTA tail nearly doubled the
  // kernel time (35 us for 103 MB).
#define K_REDUCE(M, H2)                                                                                        \
  {                                                                                                            \
    static int occ = 0;                                                                                        \
    if (occ == 0 &&                                                                                            \
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_reduce2_kernel<M, H2>, 256, 0) != cudaSuccess || occ < 1)) \
      occ = 2;                                                                                                 \
    const Grid2 g = plan(rows, C, occ, reduce_cvb(C));                                                        \
    le = launch_pdl(bn_bwd_reduce2_kernel<M, H2>, g.grid, dim3(256), 0, st, (const uint4*)dout, (const uint4*)dout2, \
                    (const uint4*)y, (const uint8_t*)bits, scale, shift, sum_g, sum_gy, rows, g.cvec, g.cvec_b,     \
                    g.rows_per_cta, fin);                                                                           \
  }
  cudaError_t le = cudaSuccess;
  if (dout2) TOK_BN2_DISPATCH_MASK(K_REDUCE, true);
  else TOK_BN2_DISPATCH_MASK(K_REDUCE, false);
#undef K_REDUCE
  if (le != cudaSuccess) return set_error(TOK_ERR_CUDA, "bn_bwd_reduce2: %s", cudaGetErrorString(le));
  return TOK_OK;
}

int tok_bn_bwd_reduce2(long long rows, int C, const void* dout, const void* dout2, const void* y, int mask_mode,
                       const void* bits, const float* scale, const float* shift, float* sum_g, float* sum_gy,
                       void* stream) {
  BwdFin fin;
  memset(&fin, 0, sizeof(fin));
  return launch_bwd_reduce2(rows, C, dout, dout2, y, mask_mode, bits, scale, shift, sum_g, sum_gy, fin, stream);
}

int tok_bn_bwd_reduce2_finalize(long long rows, int C, const void* dout, const void* dout2, const void* y, int mask_mode,
                                const void* bits, const float* scale, const float* shift, float* sum_g, float* sum_gy,
                                const float* save_mean, const float* save_invstd, const float* gamma, float* coef_a,
                                float* coef_c1, float* coef_c0, float* dgamma, float* dbeta, int accumulate,
                                unsigned* counter, void* stream) {
  return tok_bn_bwd_reduce2_finalize_cv(rows, C, C, dout, dout2, y, mask_mode, bits, scale, shift, sum_g, sum_gy, save_mean,
                                        save_invstd, gamma, coef_a, coef_c1, coef_c0, dgamma, dbeta, accumulate, counter,
                                        stream);
}

int tok_bn_bwd_reduce2_finalize_cv(long long rows, int C, int c_valid, const void* dout, const void* dout2, const void* y,
                                   int mask_mode, const void* bits, const float* scale, const float* shift, float* sum_g,
                                   float* sum_gy, const float* save_mean, const float* save_invstd, const float* gamma,
                                   float* coef_a, float* coef_c1, float* coef_c0, float* dgamma, float* dbeta,
                                   int accumulate, unsigned* counter, void* stream) {
  if (c_valid <= 0 || c_valid > C) return set_error(TOK_ERR_INVALID, "bn_bwd_reduce2_finalize: bad valid channel count");
  if (!counter || !save_mean || !save_invstd || !coef_a || !coef_c1 || !coef_c0)
    return set_error(TOK_ERR_INVALID, "bn_bwd_reduce2_finalize: counter, saved statistics and coefficient outputs are required");
  BwdFin fin;
  fin.counter = counter;
  fin.count = (float)rows;
  fin.mean = save_mean;
  fin.invstd = save_invstd;
  fin.gamma = gamma;
  fin.coef_a = coef_a;
  fin.coef_c1 = coef_c1;
  fin.coef_c0 = coef_c0;
  fin.dgamma = dgamma;
  fin.dbeta = dbeta;
  fin.accumulate = accumulate;
  fin.C = C;
  fin.Cv = c_valid;
  return launch_bwd_reduce2(rows, C, dout, dout2, y, mask_mode, bits, scale, shift, sum_g, sum_gy, fin, stream);
}

int tok_bn_bwd_apply2(long long rows, int C, const void* dout, const void* dout2, const void* y, int mask_mode,
                      const void* bits, const float* scale, const float* shift, const float* coef_a,
                      const float* coef_c1, const float* coef_c0, void* dy, void* dres, void* stream) {
  if (C <= 0 || (C % 8)) return set_error(TOK_ERR_INVALID, "bn_bwd_apply2: C must be a positive multiple of 8 (got %d)", C);
  if (rows <= 0) return set_error(TOK_ERR_INVALID, "bn_bwd_apply2: no rows");
  if (mask_mode < 0 || mask_mode > 2 || (mask_mode == MASK_BITS && !bits) || (mask_mode == MASK_Y && (!scale || !shift)))
    return set_error(TOK_ERR_INVALID, "bn_bwd_apply2: mask_mode %d needs its operands", mask_mode);
  const Grid2 g = plan(rows, C, 6);
  cudaStream_t st = (cudaStream_t)stream;
#define K_APPLY(M, H2, DR)                                                                                          \
  le = launch_pdl(bn_bwd_apply2_kernel<M, H2, DR>, g.grid, dim3(256), 0, st, (const uint4*)dout, (const uint4*)dout2, \
                  (const uint4*)y, (const uint8_t*)bits, scale, shift, coef_a, coef_c1, coef_c0, (uint4*)dy,         \
                  (uint4*)dres, rows, g.cvec, g.cvec_b, g.rows_per_cta)
  cudaError_t le = cudaSuccess;
  if (dout2 && dres) TOK_BN2_DISPATCH_MASK(K_APPLY, true, true);
  else if (dout2) TOK_BN2_DISPATCH_MASK(K_APPLY, true, false);
  else if (dres) TOK_BN2_DISPATCH_MASK(K_APPLY, false, true);
  else TOK_BN2_DISPATCH_MASK(K_APPLY, false, false);
#undef K_APPLY
  if (le != cudaSuccess) return set_error(TOK_ERR_CUDA, "bn_bwd_apply2: %s", cudaGetErrorString(le));
  return TOK_OK;
}

int tok_bn_bwd_fused_cv(long long rows, int C, int c_valid, const void* dout, const void* dout2, const void* y,
                        int mask_mode, const void* bits, const float* scale, const float* shift, float* sum_g,
                        float* sum_gy, const float* save_mean, const float* save_invstd, const float* gamma,
                        float* coef_a, float* coef_c1, float* coef_c0, float* dgamma, float* dbeta, int accumulate,
                        unsigned* counter, unsigned* release, void* dy, void* dres, void* stream) {
  if (C <= 0 || (C % 8) || rows <= 0 || c_valid <= 0 || c_valid > C)
    return set_error(TOK_ERR_INVALID, "bn_bwd_fused: bad shape (rows %lld C %d c_valid %d)", rows, C, c_valid);
  if (!counter || !release || !save_mean || !save_invstd || !coef_a || !coef_c1 || !coef_c0 || !dy)
    return set_error(TOK_ERR_INVALID, "bn_bwd_fused: counter, release word, saved statistics, coefficient and dy buffers are required");
  if (mask_mode < 0 || mask_mode > 2 || (mask_mode == MASK_BITS && !bits) || (mask_mode == MASK_Y && (!scale || !shift)))
    return set_error(TOK_ERR_INVALID, "bn_bwd_fused: mask_mode %d needs its operands", mask_mode);
  BwdFin fin;
  memset(&fin, 0, sizeof(fin));
  fin.counter = counter;
  fin.count = (float)rows;
  fin.mean = save_mean;
  fin.invstd = save_invstd;
  fin.gamma = gamma;
  fin.coef_a = coef_a;
  fin.coef_c1 = coef_c1;
  fin.coef_c0 = coef_c0;
  fin.dgamma = dgamma;
  fin.dbeta = dbeta;
  fin.accumulate = accumulate;
  fin.C = C;
  fin.Cv = c_valid;
  cudaStream_t st = (cudaStream_t)stream;
#define K_FUSED(M, H2, DR)                                                                                          \
  {                                                                                                                 \
    static int occ = 0;                                                                                             \
    if (occ == 0 &&                                                                                                 \
        (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_fused_kernel<M, H2, DR>, 256, 0) != cudaSuccess || \
         occ < 1))                                                                                                  \
      occ = 1;                                                                                                      \
    const Grid2 g = plan(rows, C, occ, 16);                                                                         \
    if ((long long)g.grid.x * g.grid.y > 148LL * occ)                                                               \
      return set_error(TOK_ERR_INVALID, "bn_bwd_fused: grid %u x %u exceeds one resident wave", g.grid.x, g.grid.y); \
    (void)launch_pdl(bn_bwd_fused_kernel<M, H2, DR>, dim3(g.grid), dim3(256), 0, st,                                                          \
        (const uint4*)dout, (const uint4*)dout2, (const uint4*)y, (const uint8_t*)bits, scale, shift, sum_g, sum_gy, \
        (uint4*)dy, (uint4*)dres, rows, g.cvec, g.cvec_b, g.rows_per_cta, fin, release);                            \
  }
  if (dout2 && dres) TOK_BN2_DISPATCH_MASK(K_FUSED, true, true);
  else if (dout2) TOK_BN2_DISPATCH_MASK(K_FUSED, true, false);
  else if (dres) TOK_BN2_DISPATCH_MASK(K_FUSED, false, true);
  else TOK_BN2_DISPATCH_MASK(K_FUSED, false, false);
#undef K_FUSED
  TOK_CHECK_LAUNCH("bn_bwd_fused");
  return TOK_OK;
}

}  // extern "C"
