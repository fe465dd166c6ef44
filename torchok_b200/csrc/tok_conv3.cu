// tok_conv3.cu — "halo" formulation of the 3x3 / stride 1 / pad 1 convolution (forward and data gradient) for layers
// with few channels per tap (Cin <= 128): ResNet's 64- and 128-wide 3x3s, every BasicBlock conv of HRNet's 18/36-channel
// branches.
//
// Why: the generic persistent kernel (tok_conv.cu) fetches the pixel operand once PER FILTER TAP (nine im2col TMA walks
// over the same rows) and the weight tile with every k-block.  ncu r2 (profiles/r2_conv3x3_c64_kernel.md): DRAM traffic
// equals the algorithmic bytes, but the SM ingests 1.25 GB for 206 MB of HBM traffic at ~27 B/clk/SM, and that ingest is
// the bound: 4.2x the floor for 64->64 @56, ~20x for HRNet's 18->18 @128.
//
// Here a CTA loads the input PATCH of an output row block once — (TR + 2) image rows x (W + 2) pixels x 64 (or 32)
// channels, one tiled 4-D TMA box whose out-of-bounds rows / columns are the zero padding — and feeds filter tap (r, s)
// to the tensor core as THE SAME shared-memory tile shifted down by r*(W+2)+s pixel rows: the accumulator row of
// padded-raster position m takes its tap from position m + r*(W+2) + s.  The shift is not a multiple of eight rows, so
// the operand start address is not 1024-byte aligned; tests/gpu/halo_probe.cu established on a B200 that tcgen05.mma
// applies the 128B / 64B swizzle to absolute shared-memory address bits (base_offset = 0 is correct for every shift).
// Two of every W + 2 accumulator rows are the padding columns: computed and thrown away (3.4 % at W = 56).
// The weights of the CTA's n-tile (all taps) are written ONCE per CTA into shared memory by the threads themselves in
// the K-major swizzled layout — for the data gradient transposed and tap-flipped on the way — so there is no per-tile
// weight traffic and no MN-major descriptor.
//
//   warp 0      TMA producer: patch ring (one patch per 64-channel block of a tile)
//   warp 1      tcgen05.mma issuer: MT accumulators (128 rows x BN) per tile, double-buffered in TMEM when they fit
//   warps 2..9  epilogue: TMEM -> bf16 -> swizzled staging tile -> coalesced 16-byte global stores of the valid pixels
//               (+ optional addend, + BatchNorm sum / sum of squares of the stored values)
//
// Reference call sites: the torch.nn.Conv2d 3x3 dispatches of timm's BasicBlock / Bottleneck built by
// torchok/models/backbones/resnet.py:363-405 and of timm's HighResolutionModule (torchok/models/backbones/hrnet.py:
// 140-192), forward and autograd.
#include <stdlib.h>
#include <string.h>

#include "../../include/tokb200.h"
#include "tok_conv.cuh"
#include "tok_internal.h"
#include "tok_ptx.cuh"

namespace tok {

struct HaloParams {
  int n_img, H, W;       // spatial size (input == output: stride 1, pad 1)
  int Cin;               // reduction channels (multiple of 8), pitch of the pixel operand
  int N;                 // output channels (multiple of 8), pitch of out / addend
  int wK, wC;            // weight tensor [wK][3][3][wC] bf16
  int transposed;        // 0: B[n][tap][c] = w[n][tap][c] (fprop); 1: B[n][tap][k] = w[k][8 - tap][n] (dgrad)
  int TR;                // output image rows per tile
  int MT;                // 128-row accumulators per tile = ceil(TR * (W + 2) / 128)
  int BN;                // n-tile width handed to the tensor core (multiple of 16)
  int BNC;               // TMEM / staging columns per accumulator (32, 64 or 128)
  int n_tiles;           // N / BN rounded up
  int KBLK;              // channel blocks per tap
  int NBUF;              // accumulator sets in TMEM (1 or 2)
  int SBUF;              // staging buffers (1 or 2)
  int stages;            // patch ring depth
  int patch_bytes;       // bytes one TMA box delivers
  int patch_stride;      // bytes between ring slots (1024-aligned)
  int slab_bytes;
  int tmem_cols;
  int dbg;               // bring-up aid (TOK_HALO_DBG): bit 0 = tap shifts rounded to 8 rows (WRONG results, timing only)
  const __nv_bfloat16* w;
  __nv_bfloat16* out;
  const __nv_bfloat16* addend;
  const uint8_t* addend_bits;   // optional ReLU mask of the addend: 1 bit per element, [pixels][N / 8] (tok_conv.cuh)
  float* col_sum;
  float* col_sqsum;
};

__device__ __forceinline__ void tma_load_tile_4d(const CUtensorMap* desc, uint64_t* bar, uint32_t smem, int32_t c0,
                                                 int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// K-major operand descriptor; swz: 2 = SWIZZLE_128B (128-byte rows), 4 = SWIZZLE_64B (64-byte rows)
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t addr, uint32_t sbo_bytes, uint64_t swz) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= swz << 61;
  return d;
}

constexpr int kHaloThreads = 320;

__device__ __forceinline__ void halo_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// KB = channels per k-block = one swizzle row: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
template <int KB, bool STATS, bool ADDEND>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const HaloParams p) {
  constexpr int kRowB = KB * 2;                    // bytes per pixel row of the patch / per weight row
  constexpr uint64_t kSwz = KB == 64 ? 2 : 4;
  constexpr uint32_t kSbo = 8 * kRowB;
  constexpr uint32_t kSwzMask = KB == 64 ? 7u : 3u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_patch = smem;
  uint8_t* smem_w = smem_patch + p.stages * p.patch_stride;
  uint8_t* smem_c = smem_w + p.slab_bytes;
  const int stage_tile = 128 * p.BNC * 2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_c + p.SBUF * stage_tile);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* tmem_full_bar = empty_bar + 4;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Wp = p.W + 2;
  const int row_blocks = (p.H + p.TR - 1) / p.TR;
  const int pix_tiles = p.n_img * row_blocks;
  const int n_t = blockIdx.x % p.n_tiles;
  const int cta_first = blockIdx.x / p.n_tiles;
  const int cta_step = gridDim.x / p.n_tiles;
  const int n0 = n_t * p.BN;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  // The padded shadow of a non-direct layer is written by the kernel launched right before this one: nothing of an
  // earlier kernel is read before the wait (barrier init and the TMEM allocation above overlap the predecessor's tail).
  pdl_wait();
  pdl_launch();
  // ---- weight slab: tile (tap, kb) = BN rows (output channel of the GEMM) x KB reduction channels, K-major, swizzled
  {
    const uint32_t slab_s = smem_u32(smem_w);
    const int tile_bytes = p.BN * kRowB;
    // four independent 16-byte loads in flight per thread (a load -> store chain per item would pay the L2 latency
    // once per item: ~30 us for the 73 KB slab of a 64 -> 64 layer)
    constexpr int kFillBatch = 4;
    if (!p.transposed) {
      // item = 16-byte chunk: 8 consecutive reduction channels of one (tap, kb, n)
      const int chunks_per_row = KB / 8;
      const int total = 9 * p.KBLK * p.BN * chunks_per_row;
      for (int i0 = threadIdx.x; i0 < total; i0 += kHaloThreads * kFillBatch) {
        uint4 v[kFillBatch];
        uint32_t dst[kFillBatch];
#pragma unroll
        for (int u = 0; u < kFillBatch; ++u) {
          const int i = i0 + u * kHaloThreads;
          v[u] = make_uint4(0, 0, 0, 0);
          dst[u] = 0;
          if (i < total) {
            const int ch = i % chunks_per_row;
            int t = i / chunks_per_row;
            const int n = t % p.BN;
            t /= p.BN;
            const int kb = t % p.KBLK;
            const int tap = t / p.KBLK;
            const int c = kb * KB + ch * 8;
            if (n0 + n < p.wK && c < p.wC) {
              const __nv_bfloat16* src = p.w + (static_cast<long long>(n0 + n) * 9 + tap) * p.wC + c;
              if ((p.wC & 7) == 0) {
                v[u] = __ldg(reinterpret_cast<const uint4*>(src));
              } else {   // unpadded weights of a channel count that is not a multiple of 8: element loads, zero tail
                uint32_t w4[4] = {0, 0, 0, 0};
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (c + j < p.wC)
                    w4[j >> 1] |= static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(src) + j)) << ((j & 1) * 16);
                v[u] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
              }
            }
            const uint32_t a = slab_s + (tap * p.KBLK + kb) * tile_bytes + n * kRowB + ch * 16;
            dst[u] = a ^ (((a >> 7) & kSwzMask) << 4);
          }
        }
#pragma unroll
        for (int u = 0; u < kFillBatch; ++u)
          if (dst[u]) sts128(dst[u], v[u]);
      }
    } else {
      // item = 16-byte chunk of the weight tensor: 8 consecutive GEMM columns n of one (tap, kb, reduction channel);
      // transposed on the way: eight 2-byte stores into the rows n .. n+7 of the K-major tile
      const int nchunks = p.BN / 8;
      const int total = 9 * p.KBLK * KB * nchunks;
      for (int i0 = threadIdx.x; i0 < total; i0 += kHaloThreads * kFillBatch) {
        uint4 v[kFillBatch];
        uint32_t dst[kFillBatch];
#pragma unroll
        for (int u = 0; u < kFillBatch; ++u) {
          const int i = i0 + u * kHaloThreads;
          v[u] = make_uint4(0, 0, 0, 0);
          dst[u] = 0;
          if (i < total) {
            const int nc = i % nchunks;
            int t = i / nchunks;
            const int kk = t % KB;
            t /= KB;
            const int kb = t % p.KBLK;
            const int tap = t / p.KBLK;
            const int k = kb * KB + kk;
            const int n = nc * 8;
            if (k < p.wK && n0 + n < p.wC) {
              const __nv_bfloat16* src = p.w + (static_cast<long long>(k) * 9 + (8 - tap)) * p.wC + n0 + n;
              if ((p.wC & 7) == 0) {
                v[u] = __ldg(reinterpret_cast<const uint4*>(src));
              } else {
                uint32_t w4[4] = {0, 0, 0, 0};
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (n0 + n + j < p.wC)
                    w4[j >> 1] |= static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(src) + j)) << ((j & 1) * 16);
                v[u] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
              }
            }
            dst[u] = slab_s + (tap * p.KBLK + kb) * tile_bytes + n * kRowB + kk * 2;
          }
        }
#pragma unroll
        for (int u = 0; u < kFillBatch; ++u) {
          if (!dst[u]) continue;
          const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = dst[u] + j * kRowB;
            const uint32_t sa = a ^ (((a >> 7) & kSwzMask) << 4);
            const uint16_t h = static_cast<uint16_t>(j & 1 ? (w4[j >> 1] >> 16) : (w4[j >> 1] & 0xffffu));
            asm volatile("st.shared.b16 [%0], %1;" ::"r"(sa), "h"(h) : "memory");
          }
        }
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int last_ksteps = (p.Cin - (p.KBLK - 1) * KB + 15) / 16;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = cta_first; t < pix_tiles; t += cta_step) {
        const int img = t / row_blocks;
        const int h0 = (t - img * row_blocks) * p.TR;
        for (int kb = 0; kb < p.KBLK; ++kb, ++it) {
          const int stage = it % p.stages;
          mbar_wait(&empty_bar[stage], ((it / p.stages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], p.patch_bytes);
          tma_load_tile_4d(&tmA, &full_bar[stage], smem_u32(smem_patch + stage * p.patch_stride), kb * KB, -1, h0 - 1,
                           img);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // One thread issues every UMMA of the CTA, so the instructions BETWEEN two tcgen05.mma are the budget: the first
      // version rebuilt both descriptors from addresses inside a rolled (tap, k) loop — 75 instructions per tap, 118
      // clocks per UMMA against 32-48 of tensor time (ncu r2o: tensor pipe 31 % active, no barrier stall anywhere).
      // Now the low descriptor words advance by precomputed per-tap deltas and the 9 x 4 loop is straight-line code.
      const uint32_t idesc = make_idesc_bf16(128, p.BN, false, false);
      const uint32_t slab_s = smem_u32(smem_w);
      const int tile_bytes = p.BN * kRowB;
      constexpr uint32_t kDescHi = static_cast<uint32_t>((static_cast<uint64_t>(kSbo >> 4) | (1ull << 14) | (kSwz << 29)));
      uint32_t a_tap[9], b_tap[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        a_tap[tap] = static_cast<uint32_t>((((tap / 3) * Wp + (tap % 3)) & ((p.dbg & 1) ? ~7 : ~0)) * kRowB) >> 4;
        b_tap[tap] = static_cast<uint32_t>(tap * p.KBLK * tile_bytes) >> 4;
      }
      uint32_t it = 0;
      int li = 0;
      for (int t = cta_first; t < pix_tiles; t += cta_step, ++li) {
        const int buf = p.NBUF == 2 ? (li & 1) : 0;
        const uint32_t use = p.NBUF == 2 ? (li >> 1) : li;
        mbar_wait(&tmem_empty_bar[buf], (use & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < p.KBLK; ++kb, ++it) {
          const int stage = it % p.stages;
          mbar_wait(&full_bar[stage], (it / p.stages) & 1);
          tc_fence_after();
          const uint32_t patch_s = smem_u32(smem_patch + stage * p.patch_stride);
          const int ksteps = kb == p.KBLK - 1 ? last_ksteps : KB / 16;
          const uint32_t b_lo0 = (((slab_s + kb * tile_bytes) >> 4) & 0x3FFFu) | (1u << 16);
          for (int mt = 0; mt < p.MT; ++mt) {
            const uint32_t acc = tmem_base + (buf * p.MT + mt) * p.BNC;
            const uint32_t a_lo0 = (((patch_s + mt * 128 * kRowB) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
              for (int k = 0; k < KB / 16; ++k) {
                if (k < ksteps) {
                  const uint64_t adesc = (static_cast<uint64_t>(kDescHi) << 32) | (a_lo0 + a_tap[tap] + 2 * k);
                  const uint64_t bdesc = (static_cast<uint64_t>(kDescHi) << 32) | (b_lo0 + b_tap[tap] + 2 * k);
                  umma_bf16(acc, adesc, bdesc, idesc, (kb | tap | k) != 0 ? 1u : 0u);
                }
              }
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;                 // 0..255
    const int row = q * 32 + lane;
    const int chunks32 = p.BNC / 32;
    const int srb = p.BNC * 2;                       // staging row bytes
    const uint32_t smask = srb >= 128 ? 7u : 3u;
    const int sshift = srb >= 128 ? 0 : 1;
    // copy-out mapping: nch 16-byte chunks per pixel, thread owns chunk `cch` of rows crow, crow + rows_per_it, ...
    int nvalid = p.N - n0;
    if (nvalid > p.BN) nvalid = p.BN;
    const int nch = nvalid / 8;
    const int rows_per_it = 256 / nch;
    const bool cact = et < rows_per_it * nch;
    const int cch = et % nch;
    const int crow = et / nch;
    // (row, column) of the padded raster advance by a fixed step per item: no division inside the item loop
    const int step_row = rows_per_it / Wp;
    const int step_col = rows_per_it - step_row * Wp;
    const float inv_wp = 1.0f / static_cast<float>(Wp);
    const int col_off = n0 + cch * 8;
    float ssum[8], ssq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ssum[j] = ssq[j] = 0.f;
    int li = 0;
    uint32_t sidx = 0;   // staging buffer use counter
    for (int t = cta_first; t < pix_tiles; t += cta_step, ++li) {
      const int img = t / row_blocks;
      const int h0 = (t - img * row_blocks) * p.TR;
      const int rows_valid = p.H - h0 < p.TR ? p.H - h0 : p.TR;
      const long long tile_off = (static_cast<long long>(img) * p.H + h0) * p.W * p.N + col_off;
      __nv_bfloat16* out_t = p.out + tile_off;
      const __nv_bfloat16* add_t = ADDEND ? p.addend + tile_off : nullptr;
      const int buf = p.NBUF == 2 ? (li & 1) : 0;
      const uint32_t use = p.NBUF == 2 ? (li >> 1) : li;
      mbar_wait(&tmem_full_bar[buf], use & 1);
      tc_fence_after();
      for (int mt = 0; mt < p.MT; ++mt, ++sidx) {
        const uint32_t cbuf_s = smem_u32(smem_c + (p.SBUF == 2 ? (sidx & 1) : 0) * stage_tile);
        // Items of this accumulator owned by the thread: global offset (-1: padding column / row past the image) and,
        // for a dgrad with addend, the addend vector — fetched NOW, before the TMEM drain and the barrier, so its DRAM
        // latency hides behind them (loaded inside the copy-out loop it doubled the kernel: 59 vs 29 us, HRNet 18->18).
        int item_off[4];
        uint4 item_add[4];
        uint32_t item_msk[4];
        {
          const int pos0 = mt * 128 + crow;
          int prow = static_cast<int>((static_cast<float>(pos0) + 0.5f) * inv_wp);   // exact: pos0 < 2^15, Wp <= 256
          int pcol = pos0 - prow * Wp;
#pragma unroll
          for (int i4 = 0; i4 < 4; ++i4) {
            const int rr = crow + i4 * rows_per_it;
            item_off[i4] = -1;
            if (cact && rr < 128 && pcol < p.W && prow < rows_valid) {
              item_off[i4] = (prow * p.W + pcol) * p.N;   // inside one image: < 2^31
              if (ADDEND) {
                item_add[i4] = __ldg(reinterpret_cast<const uint4*>(add_t + item_off[i4]));
                item_msk[i4] = 0xffu;
                if (p.addend_bits != nullptr)
                  item_msk[i4] = __ldg(p.addend_bits + ((tile_off + item_off[i4]) >> 3));
              }
            }
            pcol += step_col;
            prow += step_row;
            if (pcol >= Wp) {
              pcol -= Wp;
              ++prow;
            }
          }
        }
        if (p.SBUF == 1) halo_epi_bar();   // readers of the previous accumulator are done with the staging tile
        for (int c = half; c < chunks32; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (buf * p.MT + mt) * p.BNC + c * 32, r);
          tmem_ld_wait();
          const uint32_t rbase = cbuf_s + row * srb;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t chunk = static_cast<uint32_t>(c * 4 + g) ^ ((static_cast<uint32_t>(row) >> sshift) & smask);
            sts128(rbase + chunk * 16,
                   make_uint4(pack_bf16x2(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1])),
                              pack_bf16x2(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3])),
                              pack_bf16x2(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5])),
                              pack_bf16x2(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]))));
          }
        }
        if (mt == p.MT - 1) {
          // every accumulator of this tile is in registers / staged: hand the TMEM set back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        halo_epi_bar();   // accumulator mt staged
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          if (item_off[i4] < 0) continue;
          const int rr = crow + i4 * rows_per_it;
          const uint32_t chunk = static_cast<uint32_t>(cch) ^ ((static_cast<uint32_t>(rr) >> sshift) & smask);
          uint4 v = lds128(cbuf_s + rr * srb + chunk * 16);
          if (ADDEND) {
            const uint4 a = item_add[i4];
            const uint32_t mk = item_msk[i4];
            v.x = pack_bf16x2(bf16_lo(v.x) + (mk & 1u ? bf16_lo(a.x) : 0.f), bf16_hi(v.x) + (mk & 2u ? bf16_hi(a.x) : 0.f));
            v.y = pack_bf16x2(bf16_lo(v.y) + (mk & 4u ? bf16_lo(a.y) : 0.f), bf16_hi(v.y) + (mk & 8u ? bf16_hi(a.y) : 0.f));
            v.z = pack_bf16x2(bf16_lo(v.z) + (mk & 16u ? bf16_lo(a.z) : 0.f), bf16_hi(v.z) + (mk & 32u ? bf16_hi(a.z) : 0.f));
            v.w = pack_bf16x2(bf16_lo(v.w) + (mk & 64u ? bf16_lo(a.w) : 0.f), bf16_hi(v.w) + (mk & 128u ? bf16_hi(a.w) : 0.f));
          }
          *reinterpret_cast<uint4*>(out_t + item_off[i4]) = v;
          if (STATS) {
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo = bf16_lo(w4[e]), hi = bf16_hi(w4[e]);
              ssum[2 * e] += lo;
              ssum[2 * e + 1] += hi;
              ssq[2 * e] = fmaf(lo, lo, ssq[2 * e]);
              ssq[2 * e + 1] = fmaf(hi, hi, ssq[2 * e + 1]);
            }
          }
        }
      }
    }
    if (STATS) {
      // CTA reduction through the (now idle) staging tile: part[thread][8] per statistic, then one thread per channel
      // adds up the threads that own its chunk and issues ONE global atomic.  (The first version used shared-memory
      // float atomics: 32 threads per address, a CAS loop each — ~20 us per launch.)
      const uint32_t part_s = smem_u32(smem_c);
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        halo_epi_bar();
        if (cact) {
#pragma unroll
          for (int j = 0; j < 8; ++j) sts_f32(part_s + (et * 8 + j) * 4, pass == 0 ? ssum[j] : ssq[j]);
        }
        halo_epi_bar();
        if (et < nvalid) {
          const int c = et >> 3, j = et & 7;
          float tot = 0.f;
          for (int tt = c; tt < rows_per_it * nch; tt += nch) tot += lds_f32(part_s + (tt * 8 + j) * 4);
          atomicAdd((pass == 0 ? p.col_sum : p.col_sqsum) + n0 + et, tot);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Weight gradient by the same trick:  dW[co][r][s][ci] += sum over pixels dy[pix][co] * x[pix + (r-1, s-1)][ci].
// A tile stages the dy rows of TR image rows IN THE PADDED RASTER (box of W + 2 columns: the two columns past the
// image edge are TMA zero fill, so the padding positions contribute nothing) and the x patch of TR + 2 rows.  Both are
// MN-major UMMA operands with the pixel index as the reduction dimension: A = dy tile (M = 64 output channels),
// B = the x patch read r*(W+2)+s rows further down (N = input channels).  One fp32 accumulator per filter tap lives in
// TMEM for the whole kernel (9 x N columns) and is flushed once per CTA with 16-byte vector reductions.
// The generic kernel (tok_conv.cu) walks the pixels once per tap and per 128x64 tile: 294 us for HRNet's 18 -> 18 @128.
// M = 64 accumulator rows sit in TMEM lanes (row / 16) * 32 + row % 16 (tests/gpu/halo_probe.cu, probe 2).
struct HaloWgradParams {
  int n_img, H, W;
  int Cin, Cout;         // pitches of x / dy (multiples of 8)
  int N;                 // UMMA N = Cin rounded up to 16
  int TR;                // image rows per tile
  int ksteps;            // 16-pixel reduction steps per tile = ceil(TR * (W + 2) / 16)
  int dy_bytes, x_bytes; // TMA box bytes
  int dy_stride, x_stride;   // shared-memory bytes reserved per stage for each operand (zero-initialised guard rows included)
  int stages;
  int tmem_cols;
  int wK, wC;            // dimensions of dw (<= Cout / Cin: unpadded gradient of a channel count that is not a multiple of 8)
  float* dw;             // [wK][9][wC] fp32, accumulated
};

__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                          const HaloWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = p.dy_stride + p.x_stride;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* done_bar = empty_bar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int Wp = p.W + 2;
  const int row_blocks = (p.H + p.TR - 1) / p.TR;
  const int tiles = p.n_img * row_blocks;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  // guard rows (past the TMA boxes) are read by the last reduction steps: they must be zero (dy) / finite (x)
  {
    const uint4 z = make_uint4(0, 0, 0, 0);
    const uint32_t base = smem_u32(smem);
    for (int i = threadIdx.x; i < p.stages * stage_bytes / 16; i += 192) sts128(base + i * 16, z);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // the 200 KB zero fill above ran while the previous kernel drained
  pdl_launch();

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int img = t / row_blocks;
        const int h0 = (t - img * row_blocks) * p.TR;
        const int stage = it % p.stages;
        mbar_wait(&empty_bar[stage], ((it / p.stages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], p.dy_bytes + p.x_bytes);
        const uint32_t sdy = smem_u32(smem + stage * stage_bytes);
        tma_load_tile_4d(&tmDY, &full_bar[stage], sdy, 0, 0, h0, img);
        tma_load_tile_4d(&tmX, &full_bar[stage], sdy + p.dy_stride, 0, -1, h0 - 1, img);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(64, p.N, true, true);
      // high descriptor word: SBO 1024 B, version 1, SWIZZLE_128B; low word: address >> 4 | LBO (8192 B, unused) << 16
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t kLbo = (8192u >> 4) << 16;
      uint32_t tap_rows[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) tap_rows[tap] = static_cast<uint32_t>((tap / 3) * Wp + (tap % 3)) * 8u;
      uint32_t it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int stage = it % p.stages;
        mbar_wait(&full_bar[stage], (it / p.stages) & 1);
        tc_fence_after();
        const uint32_t sdy = smem_u32(smem + stage * stage_bytes);
        const uint32_t a_lo0 = ((sdy >> 4) & 0x3FFFu) | kLbo;
        const uint32_t b_lo0 = (((sdy + p.dy_stride) >> 4) & 0x3FFFu) | kLbo;
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint32_t acc_flag = (it | ks) != 0 ? 1u : 0u;
          const uint64_t adesc = (static_cast<uint64_t>(kDescHi) << 32) | (a_lo0 + ks * 128);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint64_t bdesc = (static_cast<uint64_t>(kDescHi) << 32) | (b_lo0 + tap_rows[tap] + ks * 128);
            umma_bf16(tmem_base + tap * p.N, adesc, bdesc, idesc, acc_flag);
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(done_bar);
    }
  } else {
    // ---- flush: warp q reads TMEM lanes q*32 .. q*32+15 = output channels q*16 .. q*16+15
    const int q = warp & 3;
    const int co = q * 16 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (blockIdx.x < tiles) {
      for (int tap = 0; tap < 9; ++tap) {
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          uint32_t r[16];
          tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tap * p.N + c0, r);
          tmem_ld_wait();
          if (lane < 16 && co < p.wK) {
            float* dst = p.dw + (static_cast<long long>(co) * 9 + tap) * p.wC + c0;
            if ((p.wC & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                if (c0 + j < p.wC)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "r"(r[j]), "r"(r[j + 1]),
                               "r"(r[j + 2]), "r"(r[j + 3])
                               : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < p.wC) atomicAdd(dst + j, __uint_as_float(r[j]));
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

int num_sms();

constexpr int kHaloSmemLimit = 232448 - 1024;   // 227 KB minus the alignment slack

struct HaloPlan {
  bool ok;
  int kb, TR, MT, BN, BNC, n_tiles, KBLK, NBUF, SBUF, stages, patch_bytes, patch_stride, slab_bytes, tmem_cols, smem;
};

static int pow2_at_least(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

// Chooses the tile geometry; ok = false when the layer is not worth / not possible on this path.
static HaloPlan halo_plan(int n_img, int H, int W, int Cin, int N) {
  HaloPlan best;
  memset(&best, 0, sizeof(best));
  if (W + 2 > 256 || Cin > 128 || Cin < 8 || N < 8 || N > 128) return best;
  const int kb = Cin <= 32 ? 32 : 64;
  const int KBLK = (Cin + kb - 1) / kb;
  const int n16 = (N + 15) / 16 * 16;
  const int BN = n16 <= 64 ? n16 : 64;
  const int n_tiles = (N + BN - 1) / BN;
  const int BNC = BN <= 32 ? 32 : 64;
  const int rowB = kb * 2;
  const int slab = 9 * KBLK * BN * rowB;
  const int Wp = W + 2;
  const int ctas = num_sms() / n_tiles;
  if (ctas < 1) return best;
  double best_cost = 1e30;
  static const int forced_tr = getenv("TOK_HALO_TR") ? atoi(getenv("TOK_HALO_TR")) : 0;
  for (int TR = 1; TR <= 16; ++TR) {
    if (forced_tr && TR != forced_tr) continue;
    if (TR > H) break;
    const int MT = (TR * Wp + 127) / 128;
    const int patch_bytes = (TR + 2) * Wp * rowB;
    const int patch_stride = (patch_bytes + 1023) / 1024 * 1024;
    if (TR + 2 > 256) break;
    int NBUF = 2 * MT * BNC <= 512 ? 2 : 1;
    if (MT * BNC > 512) break;
    const int tmem_cols = pow2_at_least(NBUF * MT * BNC);
    const int stage_tile = 128 * BNC * 2;
    const int fixed = slab + 16 * 8 + 16 + 2 * BNC * 4 + 64;
    int stages = 0, SBUF = 0;
    // preference: 2 staging buffers + >= 2 patch slots; then 1 staging buffer
    for (int sb = 2; sb >= 1 && !stages; --sb) {
      int st = (kHaloSmemLimit - fixed - sb * stage_tile) / patch_stride;
      if (st > 4) st = 4;
      if (st >= 2) {
        stages = st;
        SBUF = sb;
      }
    }
    if (!stages) continue;
    // the last accumulator reads up to MT*128 + 2*Wp + 2 patch rows: the overrun must stay inside the allocation
    const int overrun = (MT * 128 + 2 * Wp + 2) * rowB - patch_stride;
    if (overrun > slab + SBUF * stage_tile) continue;
    const int smem = stages * patch_stride + slab + SBUF * stage_tile + 16 * 8 + 16 + 2 * BNC * 4 + 64 + 1024;
    // cost model (clocks per tile): patch ingest at ~32 B/clk, UMMA at BN/2 clk per K=16 step, epilogue drain
    const int ksteps_total = (KBLK - 1) * (kb / 16) + ((Cin - (KBLK - 1) * kb + 15) / 16);
    const double load = (double)patch_bytes * KBLK / 32.0;
    const double mma = (double)MT * 9 * ksteps_total * (BN / 2.0 < 16 ? 16 : BN / 2.0);
    const double epi = (double)MT * (128.0 * BNC * 4 / 64.0 + 400.0) * (NBUF == 2 ? 1.0 : 1.3);
    double tile = load > mma ? load : mma;
    if (epi > tile) tile = epi;
    tile += 0.15 * (load + mma + epi);
    const long long tiles = (long long)n_img * ((H + TR - 1) / TR);
    const long long waves = (tiles + ctas - 1) / ctas;
    const double cost = (double)waves * tile;
    if (cost < best_cost) {
      best_cost = cost;
      best.ok = true;
      best.kb = kb; best.TR = TR; best.MT = MT; best.BN = BN; best.BNC = BNC; best.n_tiles = n_tiles; best.KBLK = KBLK;
      best.NBUF = NBUF; best.SBUF = SBUF; best.stages = stages; best.patch_bytes = patch_bytes;
      best.patch_stride = patch_stride; best.slab_bytes = slab; best.tmem_cols = tmem_cols; best.smem = smem;
    }
  }
  return best;
}

bool conv3x3_halo_eligible(int n_img, int H, int W, int Cin, int N) {
  const char* e = getenv("TOK_CONV_HALO");   // read per call: the parity scripts A/B both paths in one process
  if (e && atoi(e) == 0) return false;
  return halo_plan(n_img, H, W, Cin, N).ok;
}

// x: [n_img][H][W][Cin] bf16; w: [wK][3][3][wC] bf16; out / addend: [n_img][H][W][N] bf16.
int launch_conv3x3_halo(const void* x, int n_img, int H, int W, int Cin, int N, const void* w, int wK, int wC,
                        int transposed, void* out, const void* addend, const void* addend_bits, float* col_sum,
                        float* col_sqsum, cudaStream_t st) {
  const HaloPlan pl = halo_plan(n_img, H, W, Cin, N);
  if (!pl.ok) return set_error(TOK_ERR_INVALID, "conv3x3 halo path: unsupported shape");
  static const bool debug = getenv("TOK_HALO_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "halo n%d %dx%dx%d->%d %s: kb %d TR %d MT %d BN %d x%d KBLK %d NBUF %d SBUF %d stages %d patch %d smem %d\n",
            n_img, Cin, H, W, N, transposed ? "dgrad" : "fprop", pl.kb, pl.TR, pl.MT, pl.BN, pl.n_tiles, pl.KBLK, pl.NBUF,
            pl.SBUF, pl.stages, pl.patch_bytes, pl.smem);
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return set_error(TOK_ERR_NODRIVER, "cuTensorMapEncodeTiled unavailable");
  if (reinterpret_cast<uintptr_t>(x) & 15) return set_error(TOK_ERR_INVALID, "conv3x3 halo: operand not 16-byte aligned");
  CUtensorMap tmA;
  cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
  cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
  cuuint32_t box[4] = {(cuuint32_t)pl.kb, (cuuint32_t)(W + 2), (cuuint32_t)(pl.TR + 2), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, pl.kb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(TOK_ERR_CUDA, "conv3x3 halo: cuTensorMapEncodeTiled failed (%d) C=%d W=%d H=%d N=%d box=%d,%d,%d",
                     (int)r, Cin, W, H, n_img, pl.kb, W + 2, pl.TR + 2);
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = n_img; p.H = H; p.W = W; p.Cin = Cin; p.N = N; p.wK = wK; p.wC = wC; p.transposed = transposed;
  p.TR = pl.TR; p.MT = pl.MT; p.BN = pl.BN; p.BNC = pl.BNC; p.n_tiles = pl.n_tiles; p.KBLK = pl.KBLK; p.NBUF = pl.NBUF;
  p.SBUF = pl.SBUF; p.stages = pl.stages; p.patch_bytes = pl.patch_bytes; p.patch_stride = pl.patch_stride;
  p.slab_bytes = pl.slab_bytes; p.tmem_cols = pl.tmem_cols;
  p.dbg = getenv("TOK_HALO_DBG") ? atoi(getenv("TOK_HALO_DBG")) : 0;
  p.w = static_cast<const __nv_bfloat16*>(w);
  p.out = static_cast<__nv_bfloat16*>(out);
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.addend_bits = static_cast<const uint8_t*>(addend_bits);
  p.col_sum = col_sum;
  p.col_sqsum = col_sqsum;
  const long long pix_tiles = (long long)n_img * ((H + pl.TR - 1) / pl.TR);
  long long per = num_sms() / pl.n_tiles;
  if (per > pix_tiles) per = pix_tiles;
  const int grid = (int)per * pl.n_tiles;
  const bool stats = col_sum != nullptr;
  if (stats && addend) return set_error(TOK_ERR_INVALID, "conv3x3 halo: statistics and addend are not combined");
  cudaError_t e0 = cudaSuccess;
#define TOK_HALO_LAUNCH(KBV, ST, AD)                                                                             \
  {                                                                                                              \
    static bool configured = false;                                                                              \
    if (!configured) {                                                                                           \
      e0 = cudaFuncSetAttribute(conv3x3_halo_kernel<KBV, ST, AD>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                232448);                                                                         \
      configured = e0 == cudaSuccess;                                                                            \
    }                                                                                                            \
    if (e0 == cudaSuccess)                                                                                      \
      e0 = launch_pdl(conv3x3_halo_kernel<KBV, ST, AD>, dim3(grid), dim3(kHaloThreads), pl.smem, st, tmA, p);   \
  }
  if (pl.kb == 64) {
    if (stats) TOK_HALO_LAUNCH(64, true, false)
    else if (addend) TOK_HALO_LAUNCH(64, false, true)
    else TOK_HALO_LAUNCH(64, false, false)
  } else {
    if (stats) TOK_HALO_LAUNCH(32, true, false)
    else if (addend) TOK_HALO_LAUNCH(32, false, true)
    else TOK_HALO_LAUNCH(32, false, false)
  }
#undef TOK_HALO_LAUNCH
  if (e0 != cudaSuccess) return set_error(TOK_ERR_CUDA, "conv3x3 halo: %s", cudaGetErrorString(e0));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "conv3x3 halo launch: %s", cudaGetErrorString(e));
  return TOK_OK;
}

// ---- weight gradient
struct HaloWgradPlan {
  bool ok;
  int N, TR, ksteps, dy_bytes, x_bytes, dy_stride, x_stride, stages, tmem_cols, smem;
};

static HaloWgradPlan halo_wgrad_plan(int n_img, int H, int W, int Cin, int Cout) {
  HaloWgradPlan best;
  memset(&best, 0, sizeof(best));
  const int N = (Cin + 15) / 16 * 16;
  if (W + 2 > 256 || Cout > 64 || Cout < 8 || Cin < 8 || 9 * N > 512) return best;
  const int Wp = W + 2;
  double best_cost = 1e30;
  static const int forced_tr = getenv("TOK_HALO_WGRAD_TR") ? atoi(getenv("TOK_HALO_WGRAD_TR")) : 0;
  for (int TR = 1; TR <= 16 && TR <= H; ++TR) {
    if (forced_tr && TR != forced_tr) continue;
    const int ksteps = (TR * Wp + 15) / 16;
    const int drows = ksteps * 16;
    const int xrows = drows + 2 * Wp + 2;
    const int dy_stride = (drows * 128 + 1023) / 1024 * 1024;
    const int x_stride = (xrows * 128 + 1023) / 1024 * 1024;
    int stages = (kHaloSmemLimit - 256) / (dy_stride + x_stride);
    if (stages > 4) stages = 4;
    if (stages < 2) break;
    // per tile: reduction steps x 9 taps (shared-memory operand bound), amortised over the valid pixels; a tile count
    // that does not fill the last wave costs a whole tile
    const long long tiles = (long long)n_img * ((H + TR - 1) / TR);
    const long long waves = (tiles + num_sms() - 1) / num_sms();
    const double cost = (double)waves * (ksteps * 9.0 + 40.0);
    if (cost < best_cost) {
      best_cost = cost;
      best.ok = true;
      best.N = N; best.TR = TR; best.ksteps = ksteps; best.dy_bytes = TR * Wp * 128; best.x_bytes = (TR + 2) * Wp * 128;
      best.dy_stride = dy_stride; best.x_stride = x_stride; best.stages = stages;
      best.tmem_cols = pow2_at_least(9 * N);
      best.smem = stages * (dy_stride + x_stride) + 256 + 1024;
    }
  }
  return best;
}

bool conv3x3_wgrad_halo_eligible(int n_img, int H, int W, int Cin, int Cout) {
  const char* e = getenv("TOK_CONV_HALO");
  if (e && atoi(e) == 0) return false;
  return halo_wgrad_plan(n_img, H, W, Cin, Cout).ok;
}

static int make_halo_tmap(CUtensorMap* tm, const void* base, int n_img, int H, int W, int C, int box_w, int box_h) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return set_error(TOK_ERR_NODRIVER, "cuTensorMapEncodeTiled unavailable");
  if (reinterpret_cast<uintptr_t>(base) & 15) return set_error(TOK_ERR_INVALID, "conv3x3 halo: operand not 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(TOK_ERR_CUDA, "conv3x3 halo: cuTensorMapEncodeTiled failed (%d) C=%d W=%d H=%d N=%d box=%d,%d", (int)r,
                     C, W, H, n_img, box_w, box_h);
  return TOK_OK;
}

// x: [n_img][H][W][Cin], dy: [n_img][H][W][Cout] bf16; dw: [wK][3][3][wC] fp32 (wK <= Cout, wC <= Cin), accumulated.
int launch_conv3x3_wgrad_halo(const void* x, const void* dy, int n_img, int H, int W, int Cin, int Cout, int wK, int wC,
                              float* dw, cudaStream_t st) {
  const HaloWgradPlan pl = halo_wgrad_plan(n_img, H, W, Cin, Cout);
  if (!pl.ok) return set_error(TOK_ERR_INVALID, "conv3x3 halo wgrad: unsupported shape");
  CUtensorMap tmDY, tmX;
  int rc = make_halo_tmap(&tmDY, dy, n_img, H, W, Cout, W + 2, pl.TR);
  if (rc) return rc;
  rc = make_halo_tmap(&tmX, x, n_img, H, W, Cin, W + 2, pl.TR + 2);
  if (rc) return rc;
  static const bool debug = getenv("TOK_HALO_DEBUG") != nullptr;
  if (debug)
    fprintf(stderr, "halo wgrad n%d %dx%dx%d->%d: N %d TR %d ksteps %d stages %d smem %d tmem %d\n", n_img, Cin, H, W, Cout,
            pl.N, pl.TR, pl.ksteps, pl.stages, pl.smem, pl.tmem_cols);
  HaloWgradParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = n_img; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.N = pl.N; p.TR = pl.TR; p.ksteps = pl.ksteps;
  p.dy_bytes = pl.dy_bytes; p.x_bytes = pl.x_bytes; p.dy_stride = pl.dy_stride; p.x_stride = pl.x_stride;
  p.stages = pl.stages; p.tmem_cols = pl.tmem_cols; p.dw = dw; p.wK = wK; p.wC = wC;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "conv3x3 halo wgrad: %s", cudaGetErrorString(e));
    configured = true;
  }
  const long long tiles = (long long)n_img * ((H + pl.TR - 1) / pl.TR);
  const int grid = tiles < num_sms() ? (int)tiles : num_sms();
  cudaError_t e = launch_pdl(conv3x3_wgrad_halo_kernel, dim3(grid), dim3(192), pl.smem, st, tmDY, tmX, p);
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "conv3x3 halo wgrad launch: %s", cudaGetErrorString(e));
  return TOK_OK;
}

}  // namespace tok
