// tok_conv2.cu — CTA-pair (cta_group::2) variant of the persistent implicit-GEMM convolution, OPT-IN (TOK_CONV_2CTA=1).
//
// Status: first hardware run correct — `TOK_CONV_2CTA=1 TOK_CONV_BN=256 tests/gpu/tok_selftest pair1` (fprop 3x3
// 256->256 on 4x16x16, output and both BatchNorm column statistics against the CPU reference,
// profiles/r1_selftest_pair1_2cta.log).  The other shapes of the `pair` group (dgrad / MN-major B, addend, stride 2,
// N = 512, many tiles per pair) and its timing are still to be run, so it stays off the default path.
//
// Why: with one SM per 128x256 tile a k-block moves 96 KB through that SM's shared memory per 512 tensor
// clocks (48 KB in by TMA, 48 KB read by the four UMMAs) — ~187 B/clk against ~128 B/clk, a ~68 % ceiling on the tensor
// pipe (measured: 62 % while active on the 3x3 256->256 @14^2 fprop).  Here a CTA PAIR computes a 256x256 tile: each CTA
// stages its own 128 rows of A and only HALF of B (32 KB in, 32 KB read per k-block), the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256), and each CTA drains its own 128 accumulator rows.  Everything else — im2col-free
// TMA operand ring, double-buffered TMEM accumulators, 8 epilogue warps, swizzled staging + TMA store, BatchNorm
// statistics read back from the staged tile, fused finalize ticket — is the persistent kernel of tok_conv.cu unchanged:
// this file is that kernel with the pair protocol applied (the differences are the lines mentioning rank / lead_cta /
// *_2cta / mapa_u32 / cluster_sync_all).
//
// Pair protocol: full[s] lives in the leader CTA and is armed with the bytes of BOTH CTAs; every TMA load credits it
// through its shared::cluster address; empty[s] and tmem_full[b] are raised in both CTAs by multicast commits;
// tmem_empty[b] lives in the leader and counts the epilogue warps of both CTAs (the peer arrives remotely).
#include <stdlib.h>

#include "tok_conv.cuh"
#include "tok_pair.cuh"

namespace tok {

constexpr int kBlockM = 128;                   // rows per CTA (the pair's tile has 256)
constexpr int kBlockK = 64;
constexpr int kATile = kBlockM * kBlockK * 2;  // 16 KiB: this CTA's half of the A tile
constexpr int kEpiWarps = 8;
constexpr int kPersistThreads = 64 + 32 * kEpiWarps;  // 320

int num_sms();   // tok_conv.cu

namespace {

__device__ __forceinline__ void pixel_coords(const PixelSrc& s, int m, int& w, int& h, int& n) {
  const int pq = s.P * s.Q;
  n = m / pq;
  const int rem = m - n * pq;
  const int p = rem / s.Q;
  const int q = rem - p * s.Q;
  w = q * s.stride - s.pad;
  h = p * s.stride - s.pad;
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <int STAGES, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPersistThreads, 1)
conv_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD,
                        const ConvFwdParams p) {
  constexpr int BN = 256, CBUFS = 1, ABUFS = 0;   // 256 x 256 tile per CTA pair; staging single, addend in place
  constexpr int kBTile = (BN / 2) * kBlockK * 2;  // THIS CTA's half of the B tile
  constexpr int kStage = kATile + kBTile;
  constexpr int kCTile = kBlockM * BN * 2;
  constexpr int kBlocks = BN / 64;              // 64-column staging blocks per tile
  constexpr int kChunks = BN / 32;              // 32-column TMEM chunks per tile
  constexpr int kChunksPerWarp = kChunks / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_c = smem + STAGES * kStage;
  uint8_t* smem_d = smem_c + CBUFS * kCTile;    // prefetched addend tiles (ABUFS of them)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_d + ABUFS * kCTile);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* addend_full_bar = tmem_empty_bar + 2; // [2]
  uint64_t* addend_empty_bar = addend_full_bar + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(addend_empty_bar + 2);
  float* s_stat = reinterpret_cast<float*>(tmem_slot + 2);  // [sum | sqsum][BN], one owner lane per slot

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int m_tiles = (p.M + 2 * kBlockM - 1) / (2 * kBlockM);   // 256-row tiles, 128 rows per CTA of the pair
  const int total_tiles = n_tiles * m_tiles;
  const uint32_t rank = cluster_ctarank();
  const bool lead_cta = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
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
      mbar_init(&tmem_empty_bar[i], 2 * kEpiWarps);  // one arrive per epilogue warp of BOTH CTAs (leader's copy)
      mbar_init(&addend_full_bar[i], 1);
      mbar_init(&addend_empty_bar[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 2 * BN);   // warp 1 of both CTAs, same slot offset
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < 2 * BN; i += 32 * kEpiWarps) s_stat[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();         // the .aligned cluster barrier wants converged warps
  cluster_sync_all();   // barriers initialised and TMEM allocated in BOTH CTAs before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      int li = 0;
      for (int t = pair; t < total_tiles; t += num_pairs, ++li) {
        const int n0 = (t / m_tiles) * BN;
        const int m0 = (t % m_tiles) * 2 * kBlockM + rank * kBlockM;   // this CTA's rows
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
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStage;
          uint8_t* sb = sa + kATile;
          const uint32_t full_lead = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (lead_cta) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStage);
          const int tap = kb / cin_chunks;
          const int kc = (kb - tap * cin_chunks) * kBlockK;
          if (p.a.im2col) {
            const int r = tap / p.a.S;
            const int s = tap - r * p.a.S;
            tma_load_im2col_4d_2cta(&tmA, full_lead, smem_u32(sa), kc, w0, h0, img, static_cast<uint16_t>(s * p.a.dil),
                                    static_cast<uint16_t>(r * p.a.dil));
          } else {
            tma_load_2d_2cta(&tmA, full_lead, smem_u32(sa), kc, m0);
          }
          const int wtap = p.use_tapmap ? p.tapmap[tap] : (p.flip_taps ? (taps - 1 - tap) : tap);
          if (!B_MN) {
            tma_load_2d_2cta(&tmB, full_lead, smem_u32(sb), wtap * p.Cin + kc, n0 + rank * (BN / 2));
          } else {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j)   // this CTA's half of the N columns: two 64-column boxes
              tma_load_2d_2cta(&tmB, full_lead, smem_u32(sb + j * 8192), wtap * p.N + n0 + rank * (BN / 2) + j * 64, kc);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lead_cta && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BN, false, B_MN);
      uint32_t it = 0;
      int li = 0;
      for (int t = pair; t < total_tiles; t += num_pairs, ++li) {
        const int buf = li & 1;
        mbar_wait(&tmem_empty_bar[buf], ((li >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStage);
          const uint32_t b_addr = a_addr + kATile;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_addr + k * p.mn_kadv, p.mn_lbo, p.mn_sbo)
                                        : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16_2cta(acc, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2cta(&empty_bar[stage], 0b11);
        }
        umma_commit_2cta(&tmem_full_bar[buf], 0b11);
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue warps
    const int q = warp & 3;                 // TMEM lane quarter
    const int half = (warp - 2) >> 2;       // which interleaved set of 32-column chunks this warp drains
    const int ew = warp - 2;                // 0..7: 16-byte chunk owned in the statistics pass
    const int row = q * 32 + lane;
    const bool want_stats = p.col_sum != nullptr;
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
    const bool prof = p.prof != nullptr && (threadIdx.x == 64 || threadIdx.x == 96 + 128);
    long long pt[6] = {0, 0, 0, 0, 0, 0};
    long long tp = prof ? clock64() : 0;
#define TOK_PROF(i)                      \
  if (prof) {                            \
    const long long now = clock64();     \
    pt[i] += now - tp;                   \
    tp = now;                            \
  }
    for (int t = pair; t < total_tiles; t += num_pairs, ++li) {
      const int n0 = (t / m_tiles) * BN;
      const int m0 = (t % m_tiles) * 2 * kBlockM + rank * kBlockM;
      const int buf = li & 1;
      const int m = m0 + row;
      const bool row_ok = m < p.M;
      uint8_t* cbuf = smem_c + (CBUFS == 2 ? (li & 1) * kCTile : 0);
      // (A) single staging buffer: every epilogue thread must have finished reading it (statistics pass of the
      //     previous tile).  With two buffers the readers of this buffer (two tiles ago) are behind barrier (B) of
      //     the previous tile already.
      if (CBUFS == 1) epi_bar();
      if (leader) {
        // the TMA store that last read this staging buffer must have drained it
        if (CBUFS == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
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
      if (want_stats && prev_n0 >= 0 && prev_n0 != n0) flush_stats(prev_n0);  // rare: a finished n_tile
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
        out_row = (static_cast<long long>(img) * p.sc_H + static_cast<long long>(pp) * p.sc_sh) * p.sc_W +
                  static_cast<long long>(qq) * p.sc_sw;
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
      constexpr int kInFlight = kChunksPerWarp < 2 ? kChunksPerWarp : 2;
#pragma unroll 1
      for (int c0 = 0; c0 < kChunksPerWarp; c0 += kInFlight) {
      uint32_t r[kInFlight][32];
#pragma unroll
      for (int cc = 0; cc < kInFlight; ++cc)
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + ((c0 + cc) * 2 + half) * 32,
                           r[cc]);
      tmem_ld_wait();
      if (c0 + kInFlight >= kChunksPerWarp) {
        // accumulator is in registers: hand the TMEM buffer back to the MMA warp right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[buf]), 0));
      }
#pragma unroll
      for (int cc = 0; cc < kInFlight; ++cc) {
        const int c = (c0 + cc) * 2 + half;
        const int col0 = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[cc][j]);
        if (p.bias != nullptr) {
          if (col0 + 32 <= p.N && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
            // 8 x LDG.128 instead of 32 scalar loads per thread and chunk (the linear layers of the Swin blocks spent
            // as many instructions fetching the bias as converting the tile)
            const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = __ldg(bp + j);
              v[4 * j] += b4.x;
              v[4 * j + 1] += b4.y;
              v[4 * j + 2] += b4.z;
              v[4 * j + 3] += b4.w;
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
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[g * 8 + 2 * e] += bf16_lo(aw[e]);
              v[g * 8 + 2 * e + 1] += bf16_hi(aw[e]);
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
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int r2 = lane + rr * 32;
            uint4 vv[kBlocks];
#pragma unroll
            for (int h = 0; h < kBlocks; ++h)
              vv[h] = r2 < rows_valid ? lds128(cbuf_s + h * (kBlockM * 128) + r2 * 128 + ((ew ^ (r2 & 7)) * 16))
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
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();   // both CTAs are done with TMEM and with each other's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 2 * BN);
  }
}

template <int STAGES>
constexpr int conv_pair_smem_bytes() {
  // operand ring (A half + B half per stage) + one 128 x 256 staging tile + barriers + statistics + alignment slack
  return STAGES * (kATile + 128 * kBlockK * 2) + kBlockM * 256 * 2 + (2 * STAGES + 8) * 8 + 16 + 2 * 256 * 4 + 1024;
}

template <int STAGES, bool B_MN>
cudaError_t launch_pair_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                          const CUtensorMap& tmD, const ConvFwdParams& p, cudaStream_t st) {
  constexpr int smem = conv_pair_smem_bytes<STAGES>();
  static_assert(smem <= 232448, "pair conv kernel exceeds the 227 KB shared-memory limit");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_fwd_pair_kernel<STAGES, B_MN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int m_tiles = (p.M + 2 * kBlockM - 1) / (2 * kBlockM);
  const int n_tiles = (p.N + 255) / 256;
  const int tiles = m_tiles * n_tiles;
  const int pairs = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
  conv_fwd_pair_kernel<STAGES, B_MN><<<2 * pairs, kPersistThreads, smem, st>>>(tmA, tmB, tmC, tmD, p);
  return cudaGetLastError();
}

}  // namespace

// tmB must be built with 128-row boxes for the K-major weight matrix (each CTA fetches half of the 256 columns);
// the MN-major (dgrad) map keeps its 64 x 64 boxes.  Preconditions checked by the caller: M % 256 == 0,
// N % 256 == 0, no row scatter.
cudaError_t launch_conv_fwd_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                 const CUtensorMap& tmD, const ConvFwdParams& p, bool b_mn, cudaStream_t st) {
  return b_mn ? launch_pair_t<4, true>(tmA, tmB, tmC, tmD, p, st) : launch_pair_t<4, false>(tmA, tmB, tmC, tmD, p, st);
}

}  // namespace tok
