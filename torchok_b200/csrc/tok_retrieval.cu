// tok_retrieval.cu — the N x N retrieval search of IndexBasedMeter on sm_100a.
//
// Reference path replaced (torchok/metrics/index_base_metric.py:170-270, 444-521, 523-545): embeddings are copied to
// the CPU, a faiss IndexFlatIP / IndexFlatL2 is built over them and `index.search(queries, k + 1)` is called in
// batches — an N^2 * D brute force on host BLAS, repeated on every rank.  Here:
//
//   cosine_topk_kernel   S = Q * G^T on tcgen05 (bf16 operands, fp32 accumulate in TMEM), never materialised: a CTA
//                        keeps its 128 query rows resident in shared memory, streams gallery tiles through a TMA ring,
//                        and the epilogue threads (one per query row) fold every 128 x 128 score tile into a running
//                        top-KP list held in registers while the tensor core computes the next tile (TMEM double
//                        buffer).  KP = k + slack candidates per query.
//   topk_rerank_kernel   the KP candidates of each query are re-scored EXACTLY in fp32 from the fp32 vectors (what
//                        faiss computes) and the best k are emitted in faiss order (descending inner product /
//                        ascending squared L2, ties -> lower index), so the returned indices do not depend on the bf16
//                        rounding of the search pass.
//   l2_normalize_rows    row-wise L2 normalisation (the "cosine" of the reference's golden vectors) + bf16
//                        copy + squared norms.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/tokb200.h"
#include "tok_internal.h"
#include "tok_ptx.cuh"
#include "tok_topk.cuh"

namespace tok {
namespace {

constexpr int kQRows = 128;   // query rows per CTA (= UMMA M)
constexpr int kGTile = 128;   // gallery rows per tile (= UMMA N)
constexpr int kKB = 64;       // reduction elements per k-block (one 128-byte swizzle row)
constexpr int kStages = 4;
constexpr int kTileBytes = 128 * kKB * 2;  // 16 KiB: one k-block of either operand
constexpr int kMaxKBlocks = 8;             // resident query panel: d <= 512
constexpr int kThreads = 192;              // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (one query row per thread)

template <int KP>
__global__ void __launch_bounds__(kThreads, 1)
cosine_topk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG, int nq, int ng,
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
  const int q0 = blockIdx.x * kQRows;
  const int g_tiles = (ng + kGTile - 1) / kGTile;

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
      mbar_init(&tmem_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * kGTile);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_bar, num_kb * kTileBytes);
      for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(&tmQ, q_bar, smem_q + kb * kTileBytes, kb * kKB, q0);
      uint32_t it = 0;
      for (int t = 0; t < g_tiles; ++t) {
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % kStages;
          mbar_wait(&empty_bar[stage], ((it / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kTileBytes);
          tma_load_2d(&tmG, &full_bar[stage], smem_g + stage * kTileBytes, kb * kKB, t * kGTile);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kQRows, kGTile, false, false);
      mbar_wait(q_bar, 0);
      tc_fence_after();
      const uint32_t q_addr = smem_u32(smem_q);
      uint32_t it = 0;
      for (int t = 0; t < g_tiles; ++t) {
        const int buf = t & 1;
        mbar_wait(&tmem_empty_bar[buf], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * kGTile;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % kStages;
          mbar_wait(&full_bar[stage], (it / kStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = q_addr + kb * kTileBytes;
          const uint32_t b_addr = smem_u32(smem_g + stage * kTileBytes);
#pragma unroll
          for (int k = 0; k < kKB / 16; ++k) {
            umma_bf16(acc, make_smem_desc_sw128(a_addr + k * 32, 16, 1024),
                      make_smem_desc_sw128(b_addr + k * 32, 16, 1024), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
        }
        umma_commit(&tmem_full_bar[buf]);
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
      for (int c = 0; c < kGTile / 32; c += 2) {
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + buf * kGTile + c * 32;
        tmem_ld_32x32b_x32(taddr, r0);
        tmem_ld_32x32b_x32(taddr + 32, r1);
        tmem_ld_wait();
        const int col0 = t * kGTile + c * 32;
        if (col0 < ng) topk_fold_chunk<KP>(val, id, r0, col0, ng, g_sqnorm);
        if (col0 + 32 < ng) topk_fold_chunk<KP>(val, id, r1, col0 + 32, ng, g_sqnorm);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
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
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kGTile);
  }
}

// One warp per query: exact fp32 score of each candidate, then rank (score order, ties -> lower index) and emit k.
__global__ void __launch_bounds__(128)
topk_rerank_kernel(int nq, int d, int kp, int k, int metric, const float* __restrict__ q, const float* __restrict__ g,
                   const int* __restrict__ cand_idx, float* __restrict__ out_score, long long* __restrict__ out_idx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nq) return;
  const float* qv = q + static_cast<long long>(warp) * d;
  float my_score = 0.f;
  int my_id = -1;
  for (int c = 0; c < kp; ++c) {
    const int gi = cand_idx[static_cast<long long>(warp) * kp + c];
    float acc = 0.f;
    if (gi >= 0) {
      const float* gv = g + static_cast<long long>(gi) * d;
      for (int e = lane; e < d; e += 32) {
        const float a = qv[e], b = gv[e];
        acc += metric == 0 ? a * b : (a - b) * (a - b);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == c) {
      my_score = acc;
      my_id = gi;
    }
  }
  // rank among the kp candidates: "better" = larger IP / smaller L2, ties by lower index; missing (-1) last
  const bool valid = lane < kp && my_id >= 0;
  int rank = 0;
  for (int c = 0; c < kp; ++c) {
    const float os = __shfl_sync(0xffffffffu, my_score, c);
    const int oi = __shfl_sync(0xffffffffu, my_id, c);
    if (oi < 0 || c == lane) continue;
    const bool better = metric == 0 ? (os > my_score) : (os < my_score);
    if (better || (os == my_score && oi < my_id)) ++rank;
  }
  if (valid && rank < k) {
    out_score[static_cast<long long>(warp) * k + rank] = my_score;
    out_idx[static_cast<long long>(warp) * k + rank] = my_id;
  }
  // faiss pads missing results with -1 / -inf (+inf for L2)
  int n_valid = __popc(__ballot_sync(0xffffffffu, valid));
  for (int r = n_valid + lane; r < k; r += 32) {
    out_score[static_cast<long long>(warp) * k + r] = metric == 0 ? -INFINITY : INFINITY;
    out_idx[static_cast<long long>(warp) * k + r] = -1;
  }
}

// One warp per row: xn = x / max(|x|, eps) (or a plain copy when normalize == 0), bf16 copy, squared norm of the output.
__global__ void __launch_bounds__(128)
l2_normalize_rows_kernel(int n, int d, int normalize, const float* __restrict__ x, float* __restrict__ xn,
                         __nv_bfloat16* __restrict__ xb, float* __restrict__ sqnorm, int ldb) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* xr = x + static_cast<long long>(warp) * d;
  float ss = 0.f;
  for (int e = lane; e < d; e += 32) ss = fmaf(xr[e], xr[e], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = normalize ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
  float ss2 = 0.f;
  for (int e = lane; e < ldb; e += 32) {
    const float v = e < d ? xr[e] * inv : 0.f;
    if (e < d && xn) xn[static_cast<long long>(warp) * d + e] = v;
    if (xb) xb[static_cast<long long>(warp) * ldb + e] = __float2bfloat16(v);
    ss2 = fmaf(v, v, ss2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss2 += __shfl_xor_sync(0xffffffffu, ss2, o);
  if (lane == 0 && sqnorm) sqnorm[warp] = ss2;
}

template <int KP>
cudaError_t launch_topk(const CUtensorMap& tmQ, const CUtensorMap& tmG, int nq, int ng, int d, const float* g_sqnorm,
                        float* cand_score, int* cand_idx, cudaStream_t st) {
  constexpr int smem = (kMaxKBlocks + kStages) * kTileBytes + (2 * kStages + 5) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(cosine_topk_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  cosine_topk_kernel<KP><<<(nq + kQRows - 1) / kQRows, kThreads, smem, st>>>(tmQ, tmG, nq, ng, d, g_sqnorm,
                                                                            cand_score, cand_idx);
  return cudaGetLastError();
}

}  // namespace

// tok_retrieval2.cu: CTA-pair (cta_group::2) variant, opt-in through TOK_TOPK_2CTA=1 (not yet run on hardware)
cudaError_t launch_topk_pair(int kp, const CUtensorMap& tmQ, const CUtensorMap& tmG, int nq, int ng, int d,
                             const float* g_sqnorm, float* cand_score, int* cand_idx, cudaStream_t st);
}  // namespace tok

using namespace tok;

extern "C" {

int tok_l2_normalize_rows(int n, int d, int normalize, const float* x, float* xn, void* xn_bf16, int ld_bf16,
                          float* sqnorm, void* stream) {
  if (n <= 0 || d <= 0) return set_error(TOK_ERR_INVALID, "l2_normalize_rows: empty input");
  if (xn_bf16 && (ld_bf16 < d || (ld_bf16 % 8)))
    return set_error(TOK_ERR_INVALID, "l2_normalize_rows: bf16 pitch must be >= d and a multiple of 8");
  const long long threads = (long long)n * 32;
  l2_normalize_rows_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      n, d, normalize, x, xn, (__nv_bfloat16*)xn_bf16, sqnorm, xn_bf16 ? ld_bf16 : d);
  TOK_CHECK_LAUNCH("l2_normalize_rows");
  return TOK_OK;
}

int tok_topk_candidates(int nq, int ng, int d, int kp, const void* q_bf16, const void* g_bf16, const float* g_sqnorm,
                        float* cand_score, int* cand_idx, void* stream) {
  if (nq <= 0 || ng <= 0) return set_error(TOK_ERR_INVALID, "topk_candidates: empty query or gallery set");
  if (d <= 0 || (d % 8) || d > kMaxKBlocks * kKB)
    return set_error(TOK_ERR_INVALID, "topk_candidates: d must be a multiple of 8 in [8, %d] (got %d)", kMaxKBlocks * kKB, d);
  if (kp != 8 && kp != 16 && kp != 32) return set_error(TOK_ERR_INVALID, "topk_candidates: kp must be 8, 16 or 32");
  CUtensorMap tmQ, tmG;
  int rc = make_tmap_2d(&tmQ, q_bf16, nq, d, d, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmG, g_bf16, ng, d, d, 128);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  // CTA-pair kernel (tok_retrieval2.cu): 256 query rows per pair, each SM streams HALF of every gallery tile.  Default since
  // r2 (bit-exact neighbour tests green, 842 vs 736 TFLOP/s at N = 262144); TOK_TOPK_2CTA=0 selects the 1-SM kernel.
  static const bool pair = !(getenv("TOK_TOPK_2CTA") && atoi(getenv("TOK_TOPK_2CTA")) == 0);
  if (pair && nq >= 256) e = launch_topk_pair(kp, tmQ, tmG, nq, ng, d, g_sqnorm, cand_score, cand_idx, st);
  else if (kp == 8) e = launch_topk<8>(tmQ, tmG, nq, ng, d, g_sqnorm, cand_score, cand_idx, st);
  else if (kp == 16) e = launch_topk<16>(tmQ, tmG, nq, ng, d, g_sqnorm, cand_score, cand_idx, st);
  else e = launch_topk<32>(tmQ, tmG, nq, ng, d, g_sqnorm, cand_score, cand_idx, st);
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "topk_candidates launch: %s", cudaGetErrorString(e));
  return TOK_OK;
}

int tok_topk_rerank(int nq, int d, int kp, int k, int metric, const float* q_f32, const float* g_f32,
                    const int* cand_idx, float* out_score, long long* out_idx, void* stream) {
  if (nq <= 0 || d <= 0) return set_error(TOK_ERR_INVALID, "topk_rerank: empty input");
  if (kp <= 0 || kp > 32 || k <= 0 || k > kp) return set_error(TOK_ERR_INVALID, "topk_rerank: need 0 < k <= kp <= 32");
  if (metric != 0 && metric != 1) return set_error(TOK_ERR_INVALID, "topk_rerank: metric must be 0 (IP) or 1 (L2)");
  const long long threads = (long long)nq * 32;
  topk_rerank_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      nq, d, kp, k, metric, q_f32, g_f32, cand_idx, out_score, out_idx);
  TOK_CHECK_LAUNCH("topk_rerank");
  return TOK_OK;
}

}  // extern "C"
