// tok_retrieval2.cu — CTA-pair (cta_group::2) variant of cosine_topk_kernel, OPT-IN (TOK_TOPK_2CTA=1).
//
// Status: compiles for sm_100a; the pair mechanism is verified on a B200 by tests/gpu/gemm2cta_probe.cu (this kernel is
// that probe plus the resident query panel and the running top-KP epilogue of tok_retrieval.cu), the kernel itself has
// NOT been run yet and is not on the default path.
//
// Why: the 1-SM kernel reaches 453 TFLOP/s (33 % of the sustained bf16 peak) on the 1 M x 512 search
// because every SM streams the whole gallery through L2 -> shared memory at ~64 B/clk where ~40 B/clk is available.
// Here two SMs share each gallery tile: the pair holds 256 query rows (128 per CTA, resident), a gallery tile is 256
// rows of which each CTA fetches 128, the leader issues tcgen05.mma.cta_group::2 with M = 256, N = 256, and each CTA
// folds its own 128 x 256 score tile into its running top-KP.  Per SM the gallery stream halves for the same MMA work.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "tok_pair.cuh"
#include "tok_topk.cuh"

namespace tok {
namespace {

constexpr int kQRows = 128;   // query rows per CTA (the pair's UMMA M is 256)
constexpr int kGTile = 128;   // gallery rows fetched per CTA and k-block
constexpr int kGPair = 256;   // gallery rows per tile of the pair (= UMMA N)
constexpr int kKB = 64;
constexpr int kStages = 4;
constexpr int kTileBytes = 128 * kKB * 2;
constexpr int kMaxKBlocks = 8;
constexpr int kThreads = 192;

template <int KP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
cosine_topk_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG, int nq, int ng,
                   int d, const float* __restrict__ g_sqnorm, float* __restrict__ cand_score,
                   int* __restrict__ cand_idx) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int num_kb = (d + kKB - 1) / kKB;
  uint8_t* smem_q = smem;                              // [num_kb][128 rows][64] resident query panel
  uint8_t* smem_g = smem + kMaxKBlocks * kTileBytes;   // [kStages][128 rows][64] gallery ring
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_g + kStages * kTileBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* q_bar = empty_bar + kStages;
  uint64_t* tmem_full_bar = q_bar + 1;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQRows;                    // CTAs 2p and 2p+1 hold query rows 256p .. 256p+255
  const int g_tiles = (ng + kGPair - 1) / kGPair;        // 256 gallery rows per tile, 128 fetched by each CTA
  const uint32_t rank = cluster_ctarank();
  const bool lead_cta = rank == 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmG);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(q_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 8);   // four epilogue warps of BOTH CTAs (leader's copy is the one waited on)
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 2 * kGPair);   // warp 1 of both CTAs: 2 x 256 accumulator columns
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();   // barriers initialised and TMEM allocated in both CTAs before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // both query panels must be resident before the leader issues the first MMA: both credit the leader's q_bar
      const uint32_t q_lead = mapa_u32(smem_u32(q_bar), 0);
      if (lead_cta) mbar_arrive_expect_tx(q_bar, 2 * num_kb * kTileBytes);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d_2cta(&tmQ, q_lead, smem_u32(smem_q + kb * kTileBytes), kb * kKB, q0);
      uint32_t it = 0;
      for (int t = 0; t < g_tiles; ++t) {
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % kStages;
          mbar_wait(&empty_bar[stage], ((it / kStages) & 1) ^ 1);
          const uint32_t full_lead = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (lead_cta) mbar_arrive_expect_tx(&full_bar[stage], 2 * kTileBytes);
          tma_load_2d_2cta(&tmG, full_lead, smem_u32(smem_g + stage * kTileBytes), kb * kKB,
                           t * kGPair + static_cast<int>(rank) * kGTile);   // this CTA's half of the gallery tile
        }
      }
    }
  } else if (warp == 1) {
    if (lead_cta && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kQRows, kGPair, false, false);
      mbar_wait(q_bar, 0);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(smem_q);
      uint32_t it = 0;
      for (int t = 0; t < g_tiles; ++t) {
        const int buf = t & 1;
        mbar_wait(&tmem_empty_bar[buf], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * kGPair;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % kStages;
          mbar_wait(&full_bar[stage], (it / kStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = q_addr + kb * kTileBytes;
          const uint32_t b_addr = smem_u32(smem_g + stage * kTileBytes);
#pragma unroll
          for (int k = 0; k < kKB / 16; ++k) {
            umma_bf16_2cta(acc, make_smem_desc_sw128(a_addr + k * 32, 16, 1024),
                      make_smem_desc_sw128(b_addr + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2cta(&empty_bar[stage], 0b11);
        }
        umma_commit_2cta(&tmem_full_bar[buf], 0b11);
      }
    }
  } else {
    // ------------------------------------------------------------- epilogue: running top-KP per query row
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    float val[KP];
    int id[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) {
      val[i] = -INFINITY;
      id[i] = -1;
    }
    for (int t = 0; t < g_tiles; ++t) {
      const int buf = t & 1;
      mbar_wait(&tmem_full_bar[buf], (t >> 1) & 1);
      tc_fence_after();
      // two 32-column TMEM loads in flight per wait; each chunk is folded by a max tree + one compare (tok_topk.cuh)
#pragma unroll 1
      for (int c = 0; c < kGPair / 32; c += 2) {
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + buf * kGPair + c * 32;
        tmem_ld_32x32b_x32(taddr, r0);
        tmem_ld_32x32b_x32(taddr + 32, r1);
        tmem_ld_wait();
        const int col0 = t * kGPair + c * 32;
        if (col0 < ng) topk_fold_chunk<KP>(val, id, r0, col0, ng, g_sqnorm);
        if (col0 + 32 < ng) topk_fold_chunk<KP>(val, id, r1, col0 + 32, ng, g_sqnorm);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[buf]), 0));
    }
    const int q = q0 + row;
    if (q < nq) {
#pragma unroll
      for (int i = 0; i < KP; ++i) {
        cand_score[static_cast<long long>(q) * KP + i] = val[i];
        cand_idx[static_cast<long long>(q) * KP + i] = id[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();   // both CTAs are done with TMEM and with each other's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 2 * kGPair);
  }
}

template <int KP>
cudaError_t launch_topk_pair_t(const CUtensorMap& tmQ, const CUtensorMap& tmG, int nq, int ng, int d,
                               const float* g_sqnorm, float* cand_score, int* cand_idx, cudaStream_t st) {
  constexpr int smem = (kMaxKBlocks + kStages) * kTileBytes + (2 * kStages + 5) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(cosine_topk_pair_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int pairs = (nq + 2 * kQRows - 1) / (2 * kQRows);
  cosine_topk_pair_kernel<KP><<<2 * pairs, kThreads, smem, st>>>(tmQ, tmG, nq, ng, d, g_sqnorm, cand_score, cand_idx);
  return cudaGetLastError();
}

}  // namespace

// Same tensor maps as the 1-SM kernel (128-row boxes over the bf16 query / gallery matrices).
cudaError_t launch_topk_pair(int kp, const CUtensorMap& tmQ, const CUtensorMap& tmG, int nq, int ng, int d,
                             const float* g_sqnorm, float* cand_score, int* cand_idx, cudaStream_t st) {
  if (kp == 8) return launch_topk_pair_t<8>(tmQ, tmG, nq, ng, d, g_sqnorm, cand_score, cand_idx, st);
  if (kp == 16) return launch_topk_pair_t<16>(tmQ, tmG, nq, ng, d, g_sqnorm, cand_score, cand_idx, st);
  return launch_topk_pair_t<32>(tmQ, tmG, nq, ng, d, g_sqnorm, cand_score, cand_idx, st);
}

}  // namespace tok
