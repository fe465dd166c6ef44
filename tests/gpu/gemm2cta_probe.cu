// Bring-up probe for the CTA-pair (cta_group::2) UMMA path — NOT part of the library.  First run on a B200:
// 2048x1024x1024 correct against the CPU reference (profiles/r1_gemm2cta_probe.log).
//
// Why it exists: the 3x3 convolutions with C >= 128 sit at ~62 % of the tensor peak while active because a
// 1-SM 128x256x64 k-block moves 96 KB through one SM's shared memory per 512 tensor clocks.  With cta_group::2 a CTA
// pair computes a 256 x BN tile: each SM stages its own 128 rows of A and only HALF of B, so the shared-memory traffic
// per MMA clock drops by a third and the tile count halves.  Before the conv kernel is converted, this file checks the
// mechanism in isolation on a plain GEMM:  D[M][N] = A[M][K] . B[N][K]^T  (bf16 in, fp32 accumulate, bf16 out).
//
// Protocol (one 256 x BN tile per 2-CTA cluster, S-stage TMA ring):
//   * warp 0 (both CTAs): TMA producer.  Each CTA loads ITS 128 rows of A and ITS BN/2 rows of B into its own shared
//     memory with cp.async.bulk.tensor...cta_group::2, signalling the LEADER's (cluster rank 0) full[s] barrier;
//     the leader arms that barrier with the bytes of both CTAs.
//   * warp 1 (leader only): tcgen05.mma.cta_group::2 (M = 256, N = BN, K = 16) x 4 per 64-wide k-block; the smem
//     descriptors name the leader's offsets, the hardware reads the same offsets in the peer.  tcgen05.commit with
//     .multicast::cluster mask 0b11 releases empty[s] in BOTH CTAs and finally raises tmem_full in both.
//   * warps 2-5 (both CTAs): epilogue.  Each CTA's TMEM holds its own 128 accumulator rows x BN columns.
//   * TMEM is allocated with tcgen05.alloc.cta_group::2 by warp 1 of both CTAs (same shared-memory slot offset).
// Every wait is bounded (tok_ptx.cuh: mbar_wait traps after ~4 s), so a protocol bug ends as a launch failure.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I torchok_b200/csrc \
//              tests/gpu/gemm2cta_probe.cu -o tests/gpu/gemm2cta_probe -lcuda
// Run :   timeout 60 tests/gpu/gemm2cta_probe [M N K]          (prints max error vs a CPU fp32 reference and TFLOP/s)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tok_ptx.cuh"

using namespace tok;

constexpr int kStages = 4;
constexpr int kBN = 256;                       // tile N (UMMA N), also the TMEM column count
constexpr int kBK = 64;                        // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kThreads = 192;                  // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5 epilogue
constexpr uint32_t kABytes = 128 * kBK * 2;          // this CTA's half of the 256-row A tile
constexpr uint32_t kBBytes = (kBN / 2) * kBK * 2;    // this CTA's half of the B tile
constexpr uint32_t kSmemBytes = kStages * (kABytes + kBBytes) + 1024 /*alignment*/ + 256 /*barriers*/;

// ---------------------------------------------------------------------------------------------- cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
  return remote;
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA tile load into THIS CTA's shared memory; the transaction bytes are credited to `mbar_cluster_addr`, a
// shared::cluster address (the leader's barrier).
__device__ __forceinline__ void tma_load_2d_2cta(const CUtensorMap* desc, uint32_t mbar_cluster_addr, uint32_t dst_smem,
                                                 int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(desc)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in every CTA of `cta_mask` once the MMAs issued so far are done
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}

// ---------------------------------------------------------------------------------------------- the kernel
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm2cta_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                __nv_bfloat16* __restrict__ d, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B tiles want 1024-byte alignment
  uint8_t* aligned = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_smem = base;                                    // [stage][128 rows][128 B]
  const uint32_t b_smem = base + kStages * kABytes;                // [stage][BN/2 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(aligned + kStages * (kABytes + kBBytes));
  uint64_t* full = bars;                   // [kStages]  (only the leader's copies are waited on)
  uint64_t* empty = bars + kStages;        // [kStages]  (each CTA waits on its own copy)
  uint64_t* tmem_full = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int m_tiles = M / 256;
  const int mt = cluster_id % m_tiles, nt = cluster_id / m_tiles;
  const int kblocks = K / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);              // one arrive.expect_tx by the leader's producer; bytes from both CTAs
      mbar_init(&empty[s], 1);             // one multicast commit per use
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();                  // make the inits visible cluster-wide before any remote signal
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, kBN);          // warp 1 of BOTH CTAs, same slot offset
  __syncwarp();                                            // the .aligned cluster barrier wants converged warps
  tc_fence_before();
  cluster_sync_all();                                      // barriers initialised + TMEM allocated in both CTAs
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ producer
    if (elect_one()) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t use = kb / kStages;
        mbar_wait(&empty[s], (use & 1) ^ 1);                               // first pass: passes immediately
        const uint32_t full_leader = mapa_u32(smem_u32(&full[s]), 0);
        if (leader) mbar_arrive_expect_tx(&full[s], 2 * (kABytes + kBBytes));
        tma_load_2d_2cta(&tm_a, full_leader, a_smem + s * kABytes, kb * kBK, mt * 256 + rank * 128);
        tma_load_2d_2cta(&tm_b, full_leader, b_smem + s * kBBytes, kb * kBK, nt * kBN + rank * (kBN / 2));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA (leader)
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(256, kBN, false, false);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % kStages;
        mbar_wait(&full[s], (kb / kStages) & 1);
        tc_fence_after();
        const uint32_t a_addr = a_smem + s * kABytes, b_addr = b_smem + s * kBBytes;
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
          umma_bf16_2cta(tmem_base, adesc, bdesc, idesc, (kb | k) != 0);
        }
        umma_commit_2cta(&empty[s], 0b11);                                 // frees the stage in both CTAs
      }
      umma_commit_2cta(tmem_full, 0b11);                                   // accumulators complete, both CTAs
    }
  } else {
    // ------------------------------------------------------------------------------------------ epilogue
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;                                          // TMEM lanes a warp may touch
    const int row = quarter * 32 + lane;
    const long long grow = (long long)mt * 256 + rank * 128 + row;
    __nv_bfloat16* out = d + grow * N + (long long)nt * kBN;
    for (int col = 0; col < kBN; col += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + col, v);
      tmem_ld_wait();
      uint4* dst = reinterpret_cast<uint4*>(out + col);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        dst[i] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * i]), __uint_as_float(v[8 * i + 1])),
                            pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3])),
                            pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5])),
                            pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7])));
    }
  }
  __syncwarp();                                            // role branches diverge inside warps 0 and 1
  tc_fence_before();
  cluster_sync_all();                                      // both CTAs done with TMEM and with each other's barriers
  if (warp == 1) tmem_dealloc_2cta(tmem_base, kBN);
}

// ---------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x)                                                                                        \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess) {                                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);                \
      return 2;                                                                                      \
    }                                                                                                \
  } while (0)

static int make_map(EncodeTiledFn enc, CUtensorMap* tm, void* base, long long rows, long long cols, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

int main(int argc, char** argv) {
  const int M = argc > 3 ? atoi(argv[1]) : 4096, N = argc > 3 ? atoi(argv[2]) : 2048, K = argc > 3 ? atoi(argv[3]) : 2304;
  if (M % 256 || N % kBN || K % kBK) {
    printf("M %% 256, N %% %d, K %% %d must be 0\n", kBN, kBK);
    return 2;
  }
  std::vector<__nv_bfloat16> ha((size_t)M * K), hb((size_t)N * K);
  uint32_t seed = 12345u;
  auto rnd = [&]() {
    seed = seed * 1664525u + 1013904223u;
    return ((seed >> 9) & 0xFFFF) / 65536.0f - 0.5f;
  };
  for (auto& x : ha) x = __float2bfloat16(rnd());
  for (auto& x : hb) x = __float2bfloat16(rnd());
  __nv_bfloat16 *da, *db, *dd;
  CK(cudaMalloc(&da, ha.size() * 2));
  CK(cudaMalloc(&db, hb.size() * 2));
  CK(cudaMalloc(&dd, (size_t)M * N * 2));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xFF, (size_t)M * N * 2));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  CUtensorMap tm_a, tm_b;
  if (make_map((EncodeTiledFn)fn, &tm_a, da, M, K, 128) || make_map((EncodeTiledFn)fn, &tm_b, db, N, K, kBN / 2)) {
    printf("cuTensorMapEncodeTiled failed\n");
    return 2;
  }
  CK(cudaFuncSetAttribute(gemm2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  const int clusters = (M / 256) * (N / kBN);
  gemm2cta_kernel<<<2 * clusters, kThreads, kSmemBytes>>>(tm_a, tm_b, dd, M, N, K);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<__nv_bfloat16> hd((size_t)M * N);
  CK(cudaMemcpy(hd.data(), dd, hd.size() * 2, cudaMemcpyDeviceToHost));
  // reference on a sample of rows (every 37th) to keep the CPU part short
  double max_err = 0, max_ref = 0;
  for (int i = 0; i < M; i += 37)
    for (int j = 0; j < N; ++j) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += __bfloat162float(ha[(size_t)i * K + k]) * __bfloat162float(hb[(size_t)j * K + k]);
      max_err = fmax(max_err, fabs(acc - __bfloat162float(hd[(size_t)i * N + j])));
      max_ref = fmax(max_ref, fabs(acc));
    }
  printf("gemm2cta %dx%dx%d: max |err| %.4g of max |ref| %.4g -> %s\n", M, N, K, max_err, max_ref,
         max_err <= 1e-2 * max_ref ? "OK" : "MISMATCH");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int reps = 20;
  cudaEventRecord(e0);
  for (int r = 0; r < reps; ++r) gemm2cta_kernel<<<2 * clusters, kThreads, kSmemBytes>>>(tm_a, tm_b, dd, M, N, K);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%.3f ms per launch, %.1f TFLOP/s (non-persistent, plain-store epilogue)\n", ms / reps,
         2.0 * M * N * K / (ms / reps * 1e-3) / 1e12);
  return max_err <= 1e-2 * max_ref ? 0 : 1;
}
