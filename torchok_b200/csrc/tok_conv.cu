// tok_conv.cu — im2col-free implicit-GEMM convolution for sm_100a.
//
//   conv_fwd_kernel   : forward and data-gradient.  Pixels are GEMM rows (M), output channels are GEMM columns (N),
//                       the reduction runs over (filter tap, input-channel block).  A tiles arrive by TMA im2col
//                       (or plain 2-D TMA for 1x1/linear), B tiles by 2-D TMA from the [Cout][R*S*Cin] weight
//                       matrix, tcgen05.mma accumulates into TMEM, 4 epilogue warps drain TMEM -> bf16 -> HBM and
//                       produce the BatchNorm batch statistics (sum, sum of squares per channel) on the way out.
//   conv_wgrad_kernel : weight gradient.  Output channels are GEMM rows, input channels GEMM columns, pixels are the
//                       reduction; both operands are "MN-major" views of the same NHWC tiles.  Split-K over pixels,
//                       fp32 reductions into the [Cout][R*S*Cin] gradient.
//
// Reference call sites this replaces (all torch.nn.Conv2d dispatches): torchok/models/modules/bricks/convbnact.py:38-53,
// torchok/models/backbones/resnet.py:480-510 and the timm BasicBlock/Bottleneck convs built in resnet.py:363-405.
#include <stdlib.h>

#include "tok_conv.cuh"
#include "tok_internal.h"
#include "tok_ptx.cuh"

namespace tok {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATile = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kNumThreads = 192;               // warp0 TMA, warp1 MMA (+TMEM alloc), warps 2-5 epilogue

__device__ __forceinline__ void pixel_coords(const PixelSrc& s, int m, int& w, int& h, int& n) {
  const int pq = s.P * s.Q;
  n = m / pq;
  const int rem = m - n * pq;
  const int p = rem / s.Q;
  const int q = rem - p * s.Q;
  w = q * s.stride - s.pad;
  h = p * s.stride - s.pad;
}

// ------------------------------------------------------------------------------------------------------------------
template <int BN, int STAGES, bool B_MN>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const ConvFwdParams p) {
  constexpr int kBTile = BN * kBlockK * 2;
  constexpr int kStage = kATile + kBTile;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float* s_sum = reinterpret_cast<float*>(tmem_slot + 2);
  float* s_sq = s_sum + BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int n_t = blockIdx.x % n_tiles;
  const int m_t = blockIdx.x / n_tiles;
  const int m0 = m_t * kBlockM;
  const int n0 = n_t * BN;
  const int cin_chunks = (p.Cin + kBlockK - 1) / kBlockK;
  const int taps = p.a.R * p.a.S;
  const int num_kb = taps * cin_chunks;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < 2 * BN; i += 128) s_sum[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int w0 = 0, h0 = 0, img = 0;
      if (p.a.im2col) pixel_coords(p.a, m0, w0, h0, img);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStage;
        uint8_t* sb = sa + kATile;
        mbar_arrive_expect_tx(&full_bar[stage], kStage);
        const int tap = kb / cin_chunks;
        const int kc = (kb - tap * cin_chunks) * kBlockK;
        if (p.a.im2col) {
          const int r = tap / p.a.S;
          const int s = tap - r * p.a.S;
          tma_load_im2col_4d(&tmA, &full_bar[stage], sa, kc, w0, h0, img, static_cast<uint16_t>(s * p.a.dil),
                             static_cast<uint16_t>(r * p.a.dil));
        } else {
          tma_load_2d(&tmA, &full_bar[stage], sa, kc, m0);
        }
        const int wtap = p.use_tapmap ? p.tapmap[tap] : (p.flip_taps ? (taps - 1 - tap) : tap);
        if (!B_MN) {
          // weight matrix seen as [N rows][taps*Cin] : K-major B tile
          tma_load_2d(&tmB, &full_bar[stage], sb, wtap * p.Cin + kc, n0);
        } else {
          // weight matrix seen as [K rows = reduction channel][taps*N] : MN-major B tile, 64-column chunks
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_2d(&tmB, &full_bar[stage], sb + j * 8192, wtap * p.N + n0 + j * 64, kc);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, false, B_MN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * kStage);
        const uint32_t b_addr = a_addr + kATile;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_addr + k * p.mn_kadv, p.mn_lbo, p.mn_sbo)
                                      : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ---------------------------------------------------------------- epilogue: TMEM -> regs -> bf16 -> HBM (+stats)
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    long long out_row = m;
    if (p.scatter && row_ok) {
      const int pq = p.sc_P * p.sc_Q;
      const int img = m / pq;
      const int rem = m - img * pq;
      const int pp = rem / p.sc_Q;
      const int qq = rem - pp * p.sc_Q;
      out_row = (static_cast<long long>(img) * p.sc_H + static_cast<long long>(pp) * p.sc_sh + p.sc_oh) * p.sc_W +
                static_cast<long long>(qq) * p.sc_sw + p.sc_ow;
    }
    const bool want_stats = p.col_sum != nullptr;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (col0 >= p.N) continue;  // warp-uniform
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
      }
      const long long off = out_row * p.ldo + col0;
      if (p.addend != nullptr && row_ok) {
        const uint4* ap = reinterpret_cast<const uint4*>(p.addend + off);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (col0 + g * 8 < p.N) {
            const uint4 a = __ldg(ap + g);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[g * 8 + 2 * e] += bf16_lo(aw[e]);
              v[g * 8 + 2 * e + 1] += bf16_hi(aw[e]);
            }
          }
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      uint32_t packed[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
      if (row_ok) {
        uint4* op = reinterpret_cast<uint4*>(p.out + off);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (col0 + g * 8 < p.N) op[g] = make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
      }
      if (want_stats) {
        float s1[32], s2[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float lo = row_ok ? bf16_lo(packed[j]) : 0.f;
          const float hi = row_ok ? bf16_hi(packed[j]) : 0.f;
          s1[2 * j] = lo;
          s1[2 * j + 1] = hi;
          s2[2 * j] = lo * lo;
          s2[2 * j + 1] = hi * hi;
        }
        const float t1 = warp_transpose_reduce32(s1, lane);
        const float t2 = warp_transpose_reduce32(s2, lane);
        atomicAdd(&s_sum[c * 32 + lane], t1);
        atomicAdd(&s_sq[c * 32 + lane], t2);
      }
    }
    tc_fence_before();
    if (want_stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = threadIdx.x - 64; i < BN; i += 128) {
        if (n0 + i < p.N) {
          atomicAdd(p.col_sum + n0 + i, s_sum[i]);
          atomicAdd(p.col_sqsum + n0 + i, s_sq[i]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent variant of conv_fwd_kernel: one CTA per SM walks the (n_tile, m_tile) list.
//   warp 0      TMA producer: operand ring, and (ABUFS == 2) the addend tile of each output tile, one tile ahead
//   warp 1      MMA issuer; the accumulator is double-buffered in TMEM so tile i+1 is computed while tile i drains
//   warps 2..9  EIGHT epilogue warps: TMEM lane quarter = warp & 3, column half = (warp - 2) >> 2.  The bf16 tile is
//               staged in 128B-swizzled shared memory and written with one TMA store per 64-column block (coalesced,
//               edge-clipped by the tensor map).  With CBUFS == 2 the staging tile is double-buffered, so the store of
//               tile i drains while tile i+1 is converted.  An `addend` tile (residual gradient of a dgrad) arrives by
//               TMA too — prefetched into its own double buffer (ABUFS == 2) or, for the 128x256 tile that has no
//               shared memory to spare, loaded into the staging buffer and updated in place — so no thread ever waits
//               on a global load.  BatchNorm statistics are read back from the staged tile: warp w owns 16-byte chunk
//               w (8 channels) of every 64-column block, so each (channel, statistic) has exactly one owner lane and
//               the running sums live in shared memory without atomics until one global atomic per channel at the
//               end of an n_tile.
// `p.scatter` (strided 1x1 dgrad) keeps the direct row-scatter store.
constexpr int kEpiWarps = 8;
constexpr int kPersistThreads = 64 + 32 * kEpiWarps;  // 320

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// BRES > 0: the WEIGHT SLAB of the current n-tile (all k-blocks, BRES bytes at most) stays resident in shared memory
// while the CTA walks the m-tiles of that n-tile; the ring then carries the pixel operand only.  ncu r2
// (profiles/r2_conv3x3_c64_kernel.md): the short-reduction layers are bound by what an SM can ingest (~27-40 B/clk), and
// re-fetching the 8-32 KB weight tile with every k-block was a third to a half of it.
template <int BN, int STAGES, bool B_MN, int CBUFS, int ABUFS, int BRES, int EPI = 0>
__global__ void __launch_bounds__(kPersistThreads, 1)
conv_fwd_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD,
                        const ConvFwdParams p) {
  static_assert(CBUFS >= 1 && CBUFS <= 3, "staging buffers");
  static_assert(ABUFS == 0 || ABUFS == 2, "addend buffers");
  constexpr int kBTile = BN * kBlockK * 2;
  constexpr int kStage = kATile + (BRES > 0 ? 0 : kBTile);
  constexpr int kCTile = kBlockM * BN * 2;
  constexpr int kBlocks = BN / 64;              // 64-column staging blocks per tile
  constexpr int kChunks = BN / 32;              // 32-column TMEM chunks per tile
  constexpr int kChunksPerWarp = kChunks / 2;
  constexpr int kTmemCols = 2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512);   // allocations are powers of two
  static_assert(BN % 64 == 0 && 2 * BN <= 512, "tile width");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem + STAGES * kStage;     // resident weight slab (BRES bytes)
  uint8_t* smem_c = smem_w + BRES;
  uint8_t* smem_d = smem_c + CBUFS * kCTile;    // prefetched addend tiles (ABUFS of them)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_d + ABUFS * kCTile);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* addend_full_bar = tmem_empty_bar + 2; // [2]
  uint64_t* addend_empty_bar = addend_full_bar + 2;  // [2]
  uint64_t* wres_full_bar = addend_empty_bar + 2; // [1] slab landed
  uint64_t* wres_empty_bar = wres_full_bar + 1;   // [1] every UMMA of the slab's n-tile has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wres_empty_bar + 1);
  float* s_stat = reinterpret_cast<float*>(tmem_slot + 4);  // [sum | sqsum][BN], one owner lane per slot; 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int total_tiles = n_tiles * m_tiles;
  // Tile order.  Default (m_group == 0): n-tile major — every m-tile of one 128/256-column strip, then the next strip, so
  // the BatchNorm partial sums of a strip stay in registers and a resident weight slab is loaded once per strip.  With
  // several strips that re-reads the A operand once per strip, from HBM when it is larger than L2 (Swin stage-1 qkv:
  // 154 MB x 3).  m_group = G > 0 walks groups of G m-tiles, strip by strip inside a group, so the group's A tiles are
  // still in L2 when the next strip wants them.
  const int mg = (BRES == 0 && p.m_group > 0 && p.m_group < m_tiles) ? p.m_group : m_tiles;
  const int group_tiles = mg * n_tiles;
  auto decode_tile = [&](int t, int& n0, int& m0) {
    const int g = t / group_tiles;
    const int r = t - g * group_tiles;
    const int left = m_tiles - g * mg;
    const int gm = left < mg ? left : mg;
    const int nt = r / gm;
    n0 = nt * BN;
    m0 = (g * mg + (r - nt * gm)) * kBlockM;
  };
  const int cin_chunks = (p.Cin + kBlockK - 1) / kBlockK;
  const int taps = p.a.R * p.a.S;
  const int num_kb = taps * cin_chunks;
  const bool has_addend = p.addend != nullptr && !p.scatter;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmD);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], kEpiWarps);  // one arrive per epilogue warp
      mbar_init(&addend_full_bar[i], 1);
      mbar_init(&addend_empty_bar[i], kEpiWarps);
    }
    mbar_init(wres_full_bar, 1);
    mbar_init(wres_empty_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < 2 * BN; i += 32 * kEpiWarps) s_stat[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // operands of earlier kernels are touched only below
  pdl_launch();

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      int li = 0;
      int wres_n0 = -1;
      uint32_t wres_gen = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
        int n0, m0;
        decode_tile(t, n0, m0);
        if (ABUFS == 2 && has_addend) {
          // addend tile of THIS output tile; the epilogue is at most two tiles behind, so this runs a tile ahead
          const int ab = li & 1;
          mbar_wait(&addend_empty_bar[ab], ((li >> 1) & 1) ^ 1);
          uint32_t bytes = 0;
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            if (n0 + j * 64 < p.N) bytes += kBlockM * 128;
          mbar_arrive_expect_tx(&addend_full_bar[ab], bytes);
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            if (n0 + j * 64 < p.N)
              tma_load_2d(&tmD, &addend_full_bar[ab], smem_d + ab * kCTile + j * (kBlockM * 128), n0 + j * 64, m0);
        }
        int w0 = 0, h0 = 0, img = 0;
        if (p.a.im2col) pixel_coords(p.a, m0, w0, h0, img);
        if (BRES > 0 && n0 != wres_n0) {
          // new n-tile: the MMA warp commits wres_empty after the last tile of the old one.  (A parity wait on the
          // accumulator barrier of tile li-1 is NOT enough: with one k-block per tile the producer runs several tiles
          // ahead and the parity aliases — found by scripts/check_big_conv.py.)
          if (wres_n0 >= 0) {
            mbar_wait(wres_empty_bar, wres_gen & 1);
            ++wres_gen;
          }
          mbar_arrive_expect_tx(wres_full_bar, static_cast<uint32_t>(num_kb) * kBTile);
          for (int kb = 0; kb < num_kb; ++kb) {
            const int tap = kb / cin_chunks;
            const int kc = (kb - tap * cin_chunks) * kBlockK;
            const int wtap = p.use_tapmap ? p.tapmap[tap] : (p.flip_taps ? (taps - 1 - tap) : tap);
            uint8_t* sb = smem_w + kb * kBTile;
            if (!B_MN) {
              tma_load_2d(&tmB, wres_full_bar, sb, wtap * p.Cin + kc, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(&tmB, wres_full_bar, sb + j * 8192, wtap * p.N + n0 + j * 64, kc);
            }
          }
          wres_n0 = n0;
        }
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStage;
          uint8_t* sb = sa + kATile;
          mbar_arrive_expect_tx(&full_bar[stage], kStage);
          const int tap = kb / cin_chunks;
          const int kc = (kb - tap * cin_chunks) * kBlockK;
          if (p.a.im2col) {
            const int r = tap / p.a.S;
            const int s = tap - r * p.a.S;
            tma_load_im2col_4d(&tmA, &full_bar[stage], sa, kc, w0, h0, img, static_cast<uint16_t>(s * p.a.dil),
                               static_cast<uint16_t>(r * p.a.dil));
          } else {
            tma_load_2d(&tmA, &full_bar[stage], sa, kc, m0);
          }
          if (BRES == 0) {
            const int wtap = p.use_tapmap ? p.tapmap[tap] : (p.flip_taps ? (taps - 1 - tap) : tap);
            if (!B_MN) {
              tma_load_2d(&tmB, &full_bar[stage], sb, wtap * p.Cin + kc, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(&tmB, &full_bar[stage], sb + j * 8192, wtap * p.N + n0 + j * 64, kc);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, false, B_MN);
      uint32_t it = 0;
      int li = 0;
      int wres_n0 = -1;
      uint32_t wres_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
        const int buf = li & 1;
        mbar_wait(&tmem_empty_bar[buf], ((li >> 1) & 1) ^ 1);
        tc_fence_after();
        if (BRES > 0) {
          const int n0 = (t / m_tiles) * BN;
          if (n0 != wres_n0) {     // the producer reloads the slab exactly at these tiles
            mbar_wait(wres_full_bar, wres_phase);
            tc_fence_after();
            wres_phase ^= 1;
            wres_n0 = n0;
          }
        }
        const uint32_t acc = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStage);
          const uint32_t b_addr = BRES > 0 ? smem_u32(smem_w + kb * kBTile) : a_addr + kATile;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_addr + k * p.mn_kadv, p.mn_lbo, p.mn_sbo)
                                        : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(acc, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
        }
        umma_commit(&tmem_full_bar[buf]);
        if (BRES > 0) {   // last tile of this CTA in the n-tile: hand the slab back once these UMMAs have retired
          const int tn = t + gridDim.x;
          if (tn < total_tiles && (tn / m_tiles) * BN != wres_n0) umma_commit(wres_empty_bar);
        }
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue warps
    const int q = warp & 3;                 // TMEM lane quarter
    const int half = (warp - 2) >> 2;       // which interleaved set of 32-column chunks this warp drains
    const int ew = warp - 2;                // 0..7: 16-byte chunk owned in the statistics pass
    const int row = q * 32 + lane;
    const bool want_stats = p.col_sum != nullptr;
    const bool bias_smem = p.bias != nullptr && !want_stats && !p.scatter;   // bias table in the statistics slots
    const bool leader = threadIdx.x == 64;
    int li = 0;
    int prev_n0 = -1;
    // BatchNorm statistics: per-thread partial sums over the rows {lane, lane+32, lane+64, lane+96} of 16-byte chunk
    // `ew` of every 64-column block, carried in REGISTERS across all tiles of an n_tile ([0..7] sums, [8..15] squares)
    float sacc[kBlocks][16];
#pragma unroll
    for (int h = 0; h < kBlocks; ++h) {
#pragma unroll
      for (int j = 0; j < 16; ++j) sacc[h][j] = 0.f;
    }
    // cross-lane reduction of the partials (halving exchange: lane pair (2k, 2k+1) ends with the total of value k),
    // owner lanes publish to shared memory, then one global atomic per channel
    const uint32_t s_stat_s = smem_u32(s_stat);
    auto flush_stats = [&](int n0_done) {
      const int k = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
#pragma unroll
      for (int h = 0; h < kBlocks; ++h) {
#pragma unroll
        for (int s = 16, cnt = 8; s >= 2; s >>= 1, cnt >>= 1) {
          const bool upper = (lane & s) != 0;
#pragma unroll
          for (int i = 0; i < cnt; ++i) {
            const float send = upper ? sacc[h][i] : sacc[h][i + cnt];
            const float keep = upper ? sacc[h][i + cnt] : sacc[h][i];
            sacc[h][i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
          }
        }
        const float tot = sacc[h][0] + __shfl_xor_sync(0xffffffffu, sacc[h][0], 1);
        if ((lane & 1) == 0) sts_f32(s_stat_s + ((k >> 3) * BN + h * 64 + ew * 8 + (k & 7)) * 4, tot);
#pragma unroll
        for (int j = 0; j < 16; ++j) sacc[h][j] = 0.f;
      }
      epi_bar();
      for (int i = threadIdx.x - 64; i < BN; i += 32 * kEpiWarps) {
        if (n0_done + i < p.N) {
          atomicAdd(p.col_sum + n0_done + i, lds_f32(s_stat_s + i * 4));
          atomicAdd(p.col_sqsum + n0_done + i, lds_f32(s_stat_s + (BN + i) * 4));
        }
      }
    };
    // Statistics pass over a staged tile (see below).  With two staging buffers it is DEFERRED: the pass over tile t runs
    // in iteration t + 1, between the issue of that tile's first TMEM loads and their wait, so the ~200 instructions of
    // shared-memory reads and FMAs fill the tcgen05.ld latency instead of following the store (r3: the epilogue, not HBM,
    // bounds the K <= 256 layers: drain 714 + statistics 931 clocks per 128x128 tile back to back on 8 warps).
    auto stats_pass = [&](uint32_t cb, int rows_valid) {
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r2 = lane + rr * 32;
        uint4 vv[kBlocks];
#pragma unroll
        for (int h = 0; h < kBlocks; ++h)
          vv[h] = r2 < rows_valid ? lds128(cb + h * (kBlockM * 128) + r2 * 128 + ((ew ^ (r2 & 7)) * 16))
                                  : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int h = 0; h < kBlocks; ++h) {
          const uint32_t w4[4] = {vv[h].x, vv[h].y, vv[h].z, vv[h].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = bf16_lo(w4[e]), hi = bf16_hi(w4[e]);
            sacc[h][2 * e] += lo;
            sacc[h][2 * e + 1] += hi;
            sacc[h][8 + 2 * e] = fmaf(lo, lo, sacc[h][8 + 2 * e]);
            sacc[h][8 + 2 * e + 1] = fmaf(hi, hi, sacc[h][8 + 2 * e + 1]);
          }
        }
      }
    };
    const bool kDefer = CBUFS >= 2 && p.defer_stats != 0;
    bool pend = false;
    uint32_t pend_cb = 0;
    int pend_rows = 0, pend_n0 = -1;
    const bool prof = p.prof != nullptr && (threadIdx.x == 64 || threadIdx.x == 96 + 128);
    long long pt[6] = {0, 0, 0, 0, 0, 0};
    long long tp = prof ? clock64() : 0;
#define TOK_PROF(i)                      \
  if (prof) {                            \
    const long long now = clock64();     \
    pt[i] += now - tp;                   \
    tp = now;                            \
  }
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
      int n0, m0;
      decode_tile(t, n0, m0);
      const int buf = li & 1;
      const int m = m0 + row;
      const bool row_ok = m < p.M;
      uint8_t* cbuf = smem_c + (CBUFS >= 2 ? (li % CBUFS) * kCTile : 0);
      // (A) single staging buffer: every epilogue thread must have finished reading it (statistics pass of the
      //     previous tile).  With two buffers the readers of this buffer (two tiles ago) are behind barrier (B) of
      //     the previous tile already.
      if (CBUFS == 1) epi_bar();
      if (leader) {
        // the TMA store that last read this staging buffer must have drained it
        if (CBUFS == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        else if (CBUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else tma_store_wait_read();
        if (ABUFS == 0 && has_addend) {
          uint32_t bytes = 0;
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            if (n0 + j * 64 < p.N) bytes += kBlockM * 128;
          mbar_arrive_expect_tx(&addend_full_bar[0], bytes);
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            if (n0 + j * 64 < p.N) tma_load_2d(&tmD, &addend_full_bar[0], cbuf + j * (kBlockM * 128), n0 + j * 64, m0);
        }
      }
      if (!kDefer && want_stats && prev_n0 >= 0 && prev_n0 != n0) flush_stats(prev_n0);  // rare: a finished n_tile
      if (bias_smem && prev_n0 != n0) {
        // new strip: its bias into the (otherwise unused) statistics slots; every reader of the old values is behind
        // barrier (C) of the previous tile, barrier (B) below publishes the new ones
        for (int i = threadIdx.x - 64; i < BN; i += 32 * kEpiWarps) s_stat[i] = n0 + i < p.N ? __ldg(p.bias + n0 + i) : 0.f;
      }
      prev_n0 = n0;
      epi_bar();  // (B) staging buffer free for this tile's writers; statistics slots consistent
      TOK_PROF(0)
      long long out_row = m;
      if (p.scatter && row_ok) {
        const int pq = p.sc_P * p.sc_Q;
        const int img = m / pq;
        const int rem = m - img * pq;
        const int pp = rem / p.sc_Q;
        const int qq = rem - pp * p.sc_Q;
        out_row = (static_cast<long long>(img) * p.sc_H + static_cast<long long>(pp) * p.sc_sh + p.sc_oh) * p.sc_W +
                  static_cast<long long>(qq) * p.sc_sw + p.sc_ow;
      }
      // ReLU mask bits of the addend row (tok_conv.cuh: addend_bits): up to 32 bytes per row and tile, fetched with 16-byte
      // loads BEFORE the accumulator wait (4-byte loads per chunk inside the loop cost 8 sectors per useful one and sat
      // on the critical path: +30 % on the 512 -> 128 @28 dgrad)
      uint4 mine = make_uint4(~0u, ~0u, ~0u, ~0u);   // mask words of THIS warp's chunks: chunk (2 i + half) -> component i
      // (instantiations that never see an addend — ABUFS == 0 below the 256-wide tile — compile the mask away)
      if ((ABUFS == 2 || BN == 256) && has_addend && p.addend_bits != nullptr && row_ok) {
        const uint8_t* bp = p.addend_bits + static_cast<long long>(m) * (p.ldo >> 3) + (n0 >> 3);
        int nbytes = (p.N - n0) >> 3;
        if (nbytes > BN / 8) nbytes = BN / 8;
        uint4 t0 = make_uint4(~0u, ~0u, ~0u, ~0u), t1 = t0;
        if ((reinterpret_cast<uintptr_t>(bp) & 15) == 0 && nbytes >= 16) {
          t0 = __ldg(reinterpret_cast<const uint4*>(bp));
          if (BN == 256 && nbytes >= 32) t1 = __ldg(reinterpret_cast<const uint4*>(bp) + 1);
          else if (BN == 256 && nbytes > 16) {
            const unsigned int* wp = reinterpret_cast<const unsigned int*>(bp) + 4;
            if (nbytes >= 20) t1.x = __ldg(wp);
            if (nbytes >= 24) t1.y = __ldg(wp + 1);
            if (nbytes >= 28) t1.z = __ldg(wp + 2);
          }
        } else {
          const unsigned int* wp = reinterpret_cast<const unsigned int*>(bp);
          if (nbytes >= 4) t0.x = __ldg(wp);
          if (nbytes >= 8) t0.y = __ldg(wp + 1);
          if (nbytes >= 12) t0.z = __ldg(wp + 2);
          if (nbytes >= 16) t0.w = __ldg(wp + 3);
          if (BN == 256) {
            if (nbytes >= 20) t1.x = __ldg(wp + 4);
            if (nbytes >= 24) t1.y = __ldg(wp + 5);
            if (nbytes >= 28) t1.z = __ldg(wp + 6);
            if (nbytes >= 32) t1.w = __ldg(wp + 7);
          }
        }
        mine.x = half ? t0.y : t0.x;
        mine.y = half ? t0.w : t0.z;
        if (BN == 256) {
          mine.z = half ? t1.y : t1.x;
          mine.w = half ? t1.w : t1.z;
        }
      }
      mbar_wait(&tmem_full_bar[buf], (li >> 1) & 1);
      tc_fence_after();
      TOK_PROF(1)
      const uint32_t cbuf_s = smem_u32(cbuf);
      uint32_t abuf_s = cbuf_s;  // in-place addend by default
      if (has_addend) {
        if (ABUFS == 2) {
          mbar_wait(&addend_full_bar[li & 1], (li >> 1) & 1);
          abuf_s = smem_u32(smem_d + (li & 1) * kCTile);
        } else {
          mbar_wait(&addend_full_bar[0], li & 1);
        }
      }
      TOK_PROF(2)
      // TMEM loads are issued two at a time before the wait (the chunks are independent; two keeps the register
      // footprint of the 256-column tile inside the 168-register budget of a 320-thread CTA)
      constexpr int kInFlight = (kChunksPerWarp % 2) ? 1 : 2;   // the 192-wide tile has three chunks per warp
#pragma unroll 1
      for (int c0 = 0; c0 < kChunksPerWarp; c0 += kInFlight) {
      uint32_t r[kInFlight][32];
#pragma unroll
      for (int cc = 0; cc < kInFlight; ++cc)
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + ((c0 + cc) * 2 + half) * 32,
                           r[cc]);
      uint32_t mbits[kInFlight];
#pragma unroll
      for (int cc = 0; cc < kInFlight; ++cc) {
        const int w = c0 + cc;                        // this warp's (c0 + cc)-th chunk: word w of `mine`
        mbits[cc] = w == 0 ? mine.x : (w == 1 ? mine.y : (w == 2 ? mine.z : mine.w));
      }
      if (kDefer && c0 == 0 && pend) {   // the previous tile's statistics, while this tile's first TMEM loads travel
        stats_pass(pend_cb, pend_rows);
        pend = false;
        if (pend_n0 != n0) flush_stats(pend_n0);   // rare: a finished n_tile (every epilogue thread takes this branch)
      }
      tmem_ld_wait();
      if (c0 + kInFlight >= kChunksPerWarp) {
        // accumulator is in registers: hand the TMEM buffer back to the MMA warp right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
      }
#pragma unroll
      for (int cc = 0; cc < kInFlight; ++cc) {
        const int c = (c0 + cc) * 2 + half;
        const int col0 = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[cc][j]);
        if (bias_smem) {
          // this strip's bias from shared memory (filled at the strip change below): the 8 LDG.128 per chunk that used to
          // stand here sat behind tcgen05.wait::ld on the epilogue's critical path — +38 % on the Swin stage-1 qkv
          // projection, +25 % on fc1 (r5: scripts/linear_shapes.py, bias vs no bias)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 b4 = lds128(s_stat_s + (c * 32 + 4 * j) * 4);
            fadd2(v[4 * j], v[4 * j + 1], __uint_as_float(b4.x), __uint_as_float(b4.y));
            fadd2(v[4 * j + 2], v[4 * j + 3], __uint_as_float(b4.z), __uint_as_float(b4.w));
          }
        } else if (p.bias != nullptr) {
          if (col0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
            // 8 x LDG.128 instead of 32 scalar loads per thread and chunk (the linear layers of the Swin blocks spent
            // as many instructions fetching the bias as converting the tile)
            const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(bp + j);
              fadd2(v[4 * j], v[4 * j + 1], b4.x, b4.y);   // packed fp32 adds: the bias costs the short-K linear layers
              fadd2(v[4 * j + 2], v[4 * j + 3], b4.z, b4.w);   // 20-27 % of their time on the 8 epilogue warps (r5)
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
          }
        }
        // 128B-swizzled staging: 64-column blocks of [128 rows][128 B]; 16-byte chunk index XOR (row & 7)
        const int blk_off = (c >> 1) * (kBlockM * 128) + row * 128;
        if (has_addend) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int chunk = ((c & 1) * 4 + g) ^ (row & 7);
            const uint4 a = lds128(abuf_s + blk_off + chunk * 16);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
            const uint32_t mb = mbits[cc] >> (g * 8);   // byte g of the word: channels g*8 .. g*8+7 of this chunk
            if (EPI == 1) {
              // GELU backward: the "addend" tile is the GELU input h; out = bf16(dgrad) * gelu'(h), the product the separate
              // tok_gelu_bwd pass formed from the stored bf16 gradient (same rounding points)
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t r2 = pack_bf16x2(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
                v[g * 8 + 2 * e] = bf16_lo(r2) * gelu_grad(bf16_lo(aw[e]));
                v[g * 8 + 2 * e + 1] = bf16_hi(r2) * gelu_grad(bf16_hi(aw[e]));
              }
              continue;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[g * 8 + 2 * e] += (mb >> (2 * e)) & 1u ? bf16_lo(aw[e]) : 0.f;
              v[g * 8 + 2 * e + 1] += (mb >> (2 * e + 1)) & 1u ? bf16_hi(aw[e]) : 0.f;
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        if (p.scatter) {
          if (row_ok && col0 < p.N) {
            const long long off = out_row * p.ldo + col0;
            if (p.addend != nullptr) {  // accumulate into existing values (addend aliases out): rare, direct loads
              const uint4* ap = reinterpret_cast<const uint4*>(p.addend + off);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (col0 + g * 8 < p.N) {
                  const uint4 a = __ldg(ap + g);
                  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    packed[4 * g + e] = pack_bf16x2(v[g * 8 + 2 * e] + bf16_lo(aw[e]),
                                                    v[g * 8 + 2 * e + 1] + bf16_hi(aw[e]));
                  }
                }
              }
            }
            uint4* op = reinterpret_cast<uint4*>(p.out + off);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (col0 + g * 8 < p.N)
                op[g] = make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
          }
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int chunk = ((c & 1) * 4 + g) ^ (row & 7);
            sts128(cbuf_s + blk_off + chunk * 16,
                   make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]));
          }
        }
      }
      }
      if (ABUFS == 2 && has_addend) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&addend_empty_bar[li & 1]);
      }
      TOK_PROF(3)
      if (!p.scatter) {
        fence_proxy_async_smem();
        epi_bar();  // (C) tile staged
        TOK_PROF(4)
        if (leader) {
#pragma unroll
          for (int j = 0; j < kBlocks; ++j)
            if (n0 + j * 64 < p.N) tma_store_2d(&tmC, cbuf + j * (kBlockM * 128), n0 + j * 64, m0);
          tma_store_commit();
        }
        if (want_stats) {
          // per-channel sum / sum of squares of the STORED bf16 values, read back from the staged tile:
          // warp ew owns 16-byte chunk ew of each 64-column block; lane l reads rows l, l+32, l+64, l+96
          int rows_valid = p.M - m0;
          if (rows_valid > kBlockM) rows_valid = kBlockM;
          if (kDefer) {
            pend = true;
            pend_cb = cbuf_s;
            pend_rows = rows_valid;
            pend_n0 = n0;
          } else {
            stats_pass(cbuf_s, rows_valid);
          }
        }
        TOK_PROF(5)
      }
    }
    if (prof) {
      long long* dst = p.prof + (blockIdx.x * 2 + (threadIdx.x == 64 ? 0 : 1)) * 8;
      for (int i = 0; i < 6; ++i) dst[i] = pt[i];
      dst[6] = li;
    }
#undef TOK_PROF
    if (leader) tma_store_wait_all();
    if (kDefer && pend) stats_pass(pend_cb, pend_rows);
    if (want_stats && prev_n0 >= 0) flush_stats(prev_n0);
    if (want_stats && p.fin.counter != nullptr) {
      // last CTA standing finalizes the BatchNorm statistics (every CTA takes a ticket, also one that had no tile)
      __threadfence();   // this thread's column-sum atomics are ordered before the ticket
      epi_bar();         // ... for every epilogue thread of the CTA; s_stat is free again
      if (leader) sts_f32(s_stat_s, __int_as_float(atomicAdd(p.fin.counter, 1u) == gridDim.x - 1 ? 1 : 0));
      epi_bar();
      if (__float_as_int(lds_f32(s_stat_s)) != 0) {
        __threadfence();
        const FwdFin& f = p.fin;
        for (int c = threadIdx.x - 64; c < p.N; c += 32 * kEpiWarps) {
          const float mean = __ldcg(p.col_sum + c) / f.count;
          float var = __ldcg(p.col_sqsum + c) / f.count - mean * mean;
          var = fmaxf(var, 0.f);
          p.col_sum[c] = 0.f;   // consumed: handed back zeroed for the next step
          p.col_sqsum[c] = 0.f;
          const float invstd = rsqrtf(var + f.eps);
          const float g = f.gamma ? f.gamma[c] : 1.f;
          const float b = f.beta ? f.beta[c] : 0.f;
          f.scale[c] = g * invstd;
          f.shift[c] = b - mean * g * invstd;
          f.save_mean[c] = mean;
          f.save_invstd[c] = invstd;
          if (f.running_mean) {
            const float unbiased = f.count > 1.f ? var * f.count / (f.count - 1.f) : var;
            f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
            f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * unbiased;
          }
        }
        if (leader) *p.fin.counter = 0u;
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// grid.x = tap + taps * (n_tile + n_tiles * m_tile), grid.y = split index over 64-pixel K blocks.
template <int BN, int STAGES>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const ConvWgradParams p) {
  constexpr int kBTile = BN * kBlockK * 2;
  constexpr int kStage = kATile + kBTile;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStage);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.x.R * p.x.S;
  const int n_tiles = (p.Cin + BN - 1) / BN;
  int t = blockIdx.x;
  const int tap = t % taps;
  t /= taps;
  const int n_t = t % n_tiles;
  const int m_t = t / n_tiles;
  const int m0 = m_t * kBlockM;  // output-channel offset
  const int n0 = n_t * BN;       // input-channel offset
  const int total_chunks = (p.Mpix + kBlockK - 1) / kBlockK;
  const int kb_begin = blockIdx.y * p.chunks_per_split;
  const int kb_end = min(kb_begin + p.chunks_per_split, total_chunks);
  const int num_kb = kb_end - kb_begin;
  if (num_kb <= 0) return;  // uniform for the whole CTA; nothing allocated yet

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    if (elect_one()) {
      const int r = tap / p.x.S;
      const int s = tap - r * p.x.S;
      for (int i = 0; i < num_kb; ++i) {
        const int stage = i % STAGES;
        const uint32_t phase = (i / STAGES) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStage;
        uint8_t* sb = sa + kATile;
        mbar_arrive_expect_tx(&full_bar[stage], kStage);
        const int pix0 = (kb_begin + i) * kBlockK;
        tma_load_2d(&tmDY, &full_bar[stage], sa, m0, pix0);
        tma_load_2d(&tmDY, &full_bar[stage], sa + 8192, m0 + 64, pix0);
        if (p.x.im2col) {
          int w0, h0, img;
          pixel_coords(p.x, pix0, w0, h0, img);
#pragma unroll
          for (int j = 0; j < BN / 64; ++j)
            tma_load_im2col_4d(&tmX, &full_bar[stage], sb + j * 8192, n0 + j * 64, w0, h0, img,
                               static_cast<uint16_t>(s * p.x.dil), static_cast<uint16_t>(r * p.x.dil));
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_2d(&tmX, &full_bar[stage], sb + j * 8192, n0 + j * 64, pix0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN, true, true);
      for (int i = 0; i < num_kb; ++i) {
        const int stage = i % STAGES;
        const uint32_t phase = (i / STAGES) & 1;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * kStage);
        const uint32_t b_addr = a_addr + kATile;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(a_addr + k * p.mn_kadv, p.mn_lbo, p.mn_sbo);
          const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * p.mn_kadv, p.mn_lbo, p.mn_sbo);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int co = m0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (co < p.Cout) {
        float* dst = p.dw + static_cast<long long>(co) * p.ldw + static_cast<long long>(tap) * p.Cin + col0;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          // 16-byte vector reductions (REDG.E.ADD.F32x4): a quarter of the L2 atomic requests; Cin % 8 == 0 keeps
          // every group of four columns entirely inside or outside the tensor
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (col0 + j < p.Cin)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "r"(r[j]), "r"(r[j + 1]),
                           "r"(r[j + 2]), "r"(r[j + 3])
                           : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.Cin) atomicAdd(dst + j, __uint_as_float(r[j]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Host side
void mn_desc_geometry(int* lbo, int* sbo, int* kadv) {
  static int v[3] = {-1, -1, -1};
  if (v[0] < 0) {
    const char* e;
    v[0] = (e = getenv("TOK_MN_LBO")) ? atoi(e) : 8192;
    v[1] = (e = getenv("TOK_MN_SBO")) ? atoi(e) : 1024;
    v[2] = (e = getenv("TOK_MN_KADV")) ? atoi(e) : 2048;
  }
  *lbo = v[0];
  *sbo = v[1];
  *kadv = v[2];
}
template <int BN, int STAGES>
constexpr int conv_smem_bytes() {
  return STAGES * (kATile + BN * kBlockK * 2) + (2 * STAGES + 1) * 8 + 16 + 2 * BN * 4 + 1024;
}

template <int BN, int STAGES, int CBUFS, int ABUFS, int BRES>
constexpr int conv_persist_smem_bytes() {
  return STAGES * (kATile + (BRES > 0 ? 0 : BN * kBlockK * 2)) + BRES + (CBUFS + ABUFS) * kBlockM * BN * 2 +
         (2 * STAGES + 10) * 8 + 16 + 2 * BN * 4 + 1024;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int STAGES, bool B_MN, int CBUFS, int ABUFS, int BRES = 0, int EPI = 0>
static cudaError_t launch_persist_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                    const CUtensorMap& tmD, const ConvFwdParams& p, cudaStream_t st) {
  constexpr int smem = conv_persist_smem_bytes<BN, STAGES, CBUFS, ABUFS, BRES>();
  static_assert(smem <= 232448, "persistent conv kernel exceeds the 227 KB shared-memory limit");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_fwd_persist_kernel<BN, STAGES, B_MN, CBUFS, ABUFS, BRES, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  return launch_pdl(conv_fwd_persist_kernel<BN, STAGES, B_MN, CBUFS, ABUFS, BRES, EPI>, dim3(grid), dim3(kPersistThreads), smem,
                    st, tmA, tmB, tmC, tmD, p);
}

// Tile configurations (shared memory: operand ring + staging + addend buffers):
//   BN  64: 6 x 24 KB ring, 2 staging, 2 addend (when the launch has an addend)          = 208 KB
//   BN 128: 4 x 32 KB ring, 2 staging            | 3 x 32 KB ring, 2 staging, 2 addend   = 192 / 224 KB
//   BN 256: 3 x 48 KB ring, 1 staging, addend loaded in place                            = 208 KB
//   BN 192: 3 x 40 KB ring, 2 x 48 KB staging, no addend (plain GEMM launches, N a multiple of 96)  = 216 KB
cudaError_t launch_conv_fwd_persist(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                    const CUtensorMap& tmD, const ConvFwdParams& p, int bn, bool b_mn,
                                    cudaStream_t st) {
  const bool add = p.addend != nullptr && !p.scatter;
  if (p.addend_mode == 1) {   // GELU backward in the epilogue: one instantiation (MN-major weights, 128-wide tile)
    if (!(add && bn == 128 && b_mn)) return cudaErrorInvalidValue;
    return launch_persist_t<128, 3, true, 2, 2, 0, 1>(tmA, tmB, tmC, tmD, p, st);
  }
  // weight slab of one n-tile (all taps x channel blocks): resident when it fits next to the ring (TOK_CONV_BRES=0: off)
  static const bool bres_on = !(getenv("TOK_CONV_BRES") && atoi(getenv("TOK_CONV_BRES")) == 0);
  const long long num_kb = (long long)p.a.R * p.a.S * ((p.Cin + kBlockK - 1) / kBlockK);
  const long long slab = num_kb * bn * kBlockK * 2;
  const long long m_tiles = (p.M + kBlockM - 1) / kBlockM;
  // the slab load is a pipeline bubble once per n-tile and CTA: worth it when a CTA then re-uses it for >= 8 m-tiles
  const bool reuse = m_tiles >= 8LL * num_sms();
  if (bn == 64) {
    if (add)
      return b_mn ? launch_persist_t<64, 6, true, 2, 2>(tmA, tmB, tmC, tmD, p, st)
                  : launch_persist_t<64, 6, false, 2, 2>(tmA, tmB, tmC, tmD, p, st);
    if (bres_on && reuse && slab <= 73728)
      return b_mn ? launch_persist_t<64, 6, true, 2, 0, 73728>(tmA, tmB, tmC, tmD, p, st)
                  : launch_persist_t<64, 6, false, 2, 0, 73728>(tmA, tmB, tmC, tmD, p, st);
    return b_mn ? launch_persist_t<64, 6, true, 2, 0>(tmA, tmB, tmC, tmD, p, st)
                : launch_persist_t<64, 6, false, 2, 0>(tmA, tmB, tmC, tmD, p, st);
  }
  if (bn == 128) {
    if (add)
      return b_mn ? launch_persist_t<128, 3, true, 2, 2>(tmA, tmB, tmC, tmD, p, st)
                  : launch_persist_t<128, 3, false, 2, 2>(tmA, tmB, tmC, tmD, p, st);
    // measured r2 (profiles/r2_resnet50_step.md): no gain for the 128-wide tile (1x1 64->256: 133 vs 129 us), -10 % for the
    // 64-wide 3x3 (180 -> 161 us); so the 128-wide variant needs TOK_CONV_BRES=2
    static const bool bres128 = getenv("TOK_CONV_BRES") && atoi(getenv("TOK_CONV_BRES")) == 2;
    if (bres128 && reuse && slab <= 65536)
      return b_mn ? launch_persist_t<128, 5, true, 2, 0, 65536>(tmA, tmB, tmC, tmD, p, st)
                  : launch_persist_t<128, 5, false, 2, 0, 65536>(tmA, tmB, tmC, tmD, p, st);
    // r3 experiment (opt-in, TOK_CONV_CBUFS3=1): three staging tiles and a 3-stage ring (192 KB either way).  The short-
    // reduction layers wait ~570 clk per tile for the store two tiles back to release its staging buffer (phase profile,
    // profiles/r2_logs) — but a third buffer does not help (64->256 @56: 134 vs 133 us) and the shallower ring costs
    // the 4-k-block layers (256->1024 @14: 52 vs 44 us): that wait is HBM write back-pressure, not a missing buffer.
    static const bool cb3 = getenv("TOK_CONV_CBUFS3") && atoi(getenv("TOK_CONV_CBUFS3")) == 1;
    const long long kblocks = (long long)p.a.R * p.a.S * ((p.Cin + kBlockK - 1) / kBlockK);
    if (cb3 && kblocks <= 4)
      return b_mn ? launch_persist_t<128, 3, true, 3, 0>(tmA, tmB, tmC, tmD, p, st)
                  : launch_persist_t<128, 3, false, 3, 0>(tmA, tmB, tmC, tmD, p, st);
    return b_mn ? launch_persist_t<128, 4, true, 2, 0>(tmA, tmB, tmC, tmD, p, st)
                : launch_persist_t<128, 4, false, 2, 0>(tmA, tmB, tmC, tmD, p, st);
  }
  if (bn == 256)
    return b_mn ? launch_persist_t<256, 3, true, 1, 0>(tmA, tmB, tmC, tmD, p, st)
                : launch_persist_t<256, 3, false, 1, 0>(tmA, tmB, tmC, tmD, p, st);
  // 128x192: the Swin widths are multiples of 96, so 128- and 256-wide strips leave a ragged one (N = 288, 576, 1152) or
  // run the long reductions at the 128-wide tile's operand-fetch cap (N = 384).  3 x 40 KB ring, 2 x 48 KB staging.
  // Launches without an addend only.
  if (bn == 192 && !add)
    return b_mn ? launch_persist_t<192, 3, true, 2, 0>(tmA, tmB, tmC, tmD, p, st)
                : launch_persist_t<192, 3, false, 2, 0>(tmA, tmB, tmC, tmD, p, st);
  return cudaErrorInvalidValue;
}

template <int BN, int STAGES, bool B_MN>
static cudaError_t launch_fwd_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvFwdParams& p,
                                cudaStream_t st) {
  constexpr int smem = conv_smem_bytes<BN, STAGES>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_fwd_kernel<BN, STAGES, B_MN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int n_tiles = (p.N + BN - 1) / BN;
  conv_fwd_kernel<BN, STAGES, B_MN><<<m_tiles * n_tiles, kNumThreads, smem, st>>>(tmA, tmB, p);
  return cudaGetLastError();
}

cudaError_t launch_conv_fwd(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvFwdParams& p, int bn,
                            bool b_mn, cudaStream_t st) {
  if (bn == 64) return b_mn ? launch_fwd_t<64, 4, true>(tmA, tmB, p, st) : launch_fwd_t<64, 4, false>(tmA, tmB, p, st);
  if (bn == 128)
    return b_mn ? launch_fwd_t<128, 3, true>(tmA, tmB, p, st) : launch_fwd_t<128, 3, false>(tmA, tmB, p, st);
  return cudaErrorInvalidValue;
}

template <int BN, int STAGES>
static cudaError_t launch_wgrad_t(const CUtensorMap& tmDY, const CUtensorMap& tmX, const ConvWgradParams& p,
                                  int splits, cudaStream_t st) {
  constexpr int smem = conv_smem_bytes<BN, STAGES>();
  static bool configured = false;
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(conv_wgrad_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int taps = p.x.R * p.x.S;
  const int m_tiles = (p.Cout + kBlockM - 1) / kBlockM;
  const int n_tiles = (p.Cin + BN - 1) / BN;
  dim3 grid(taps * m_tiles * n_tiles, splits);
  return launch_pdl(conv_wgrad_kernel<BN, STAGES>, grid, dim3(kNumThreads), smem, st, tmDY, tmX, p);
}

cudaError_t launch_conv_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const ConvWgradParams& p, int bn,
                              int splits, cudaStream_t st) {
  if (bn == 64) return launch_wgrad_t<64, 4>(tmDY, tmX, p, splits, st);
  if (bn == 128) return launch_wgrad_t<128, 3>(tmDY, tmX, p, splits, st);
  return cudaErrorInvalidValue;
}

}  // namespace tok
