// Bring-up probe for the halo formulation of the 3x3 convolution — NOT part of the library.
//
// Question it answers on hardware: may the start address of a K-major SWIZZLE_128B UMMA operand be shifted by a whole
// number of 128-byte rows that is NOT a multiple of 8 (i.e. not 1024-byte aligned), and if so, which value does the
// descriptor's base_offset field (bits 49-51) need?  The halo conv keeps an input patch of (rows + 2) x (W + 2) pixels
// x 64 channels in shared memory exactly as TMA writes it (16-byte chunk index XOR (row & 7), rows counted from a
// 1024-byte aligned base) and feeds tap (r, s) to the tensor core as "the same tile, r*(W+2)+s rows further down".
//
// Test: A = 384 rows x 64 bf16 written in that layout, B = 64x64 identity; D = A[shift .. shift+128) . B must equal
// the shifted rows exactly.  Also probes SWIZZLE_64B (32 channels per row) for the narrow HRNet branches.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I torchok_b200/csrc tests/gpu/halo_probe.cu \
//             -o tests/gpu/halo_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tok_ptx.cuh"

using namespace tok;

constexpr int kRows = 384;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off,
                                              uint32_t swz) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(swz) << 61;   // 2 = 128B, 4 = 64B, 6 = 32B
  return d;
}

// mode 0: SWIZZLE_128B, 64 bf16 per row (128 B rows);  mode 1: SWIZZLE_64B, 32 bf16 per row (64 B rows)
__global__ void __launch_bounds__(128, 1) halo_probe_kernel(float* out, int shift, int use_base_off, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int row_bytes = mode == 0 ? 128 : 64;
  const int kcols = mode == 0 ? 64 : 32;
  uint8_t* sa = smem;                         // kRows rows
  uint8_t* sb = smem + kRows * 128;           // 64 rows (n) of B, K-major, same swizzle
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // fill A: logical (p, c) -> value; physical = swizzle on ABSOLUTE address bits like TMA does
  for (int i = tid; i < kRows * kcols; i += 128) {
    const int p = i / kcols, c = i % kcols;
    const float v = static_cast<float>((p * 3 + c * 5) % 251) - 125.f;
    uint32_t off = p * row_bytes + c * 2;
    const uint32_t abs = smem_u32(sa) + off;
    uint32_t sw;
    if (mode == 0) sw = abs ^ (((abs >> 7) & 7) << 4);
    else sw = abs ^ (((abs >> 7) & 3) << 4);
    *reinterpret_cast<__nv_bfloat16*>(sa + (sw - smem_u32(sa))) = __float2bfloat16(v);
  }
  for (int i = tid; i < 64 * kcols; i += 128) {
    const int n = i / kcols, k = i % kcols;
    uint32_t off = n * row_bytes + k * 2;
    const uint32_t abs = smem_u32(sb) + off;
    uint32_t sw;
    if (mode == 0) sw = abs ^ (((abs >> 7) & 7) << 4);
    else sw = abs ^ (((abs >> 7) & 3) << 4);
    *reinterpret_cast<__nv_bfloat16*>(sb + (sw - smem_u32(sb))) = __float2bfloat16(n == k ? 1.f : 0.f);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, 64, false, false);
    const uint32_t a0 = smem_u32(sa) + shift * row_bytes;
    const uint32_t b0 = smem_u32(sb);
    const uint32_t sbo = mode == 0 ? 1024 : 512;
    const uint32_t swz = mode == 0 ? 2 : 4;
    for (int k = 0; k < kcols / 16; ++k) {
      const uint32_t aa = a0 + k * 32;
      const uint32_t bo = use_base_off ? ((aa >> 7) & 7) : 0;
      umma_bf16(tmem, make_desc(aa, 16, sbo, bo, swz), make_desc(b0 + k * 32, 16, sbo, 0, swz), idesc, k != 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}


// ---- probe 2 (halo weight gradient): MN-major operands.  A = dy tile [K rows = pixels][64 co] (M = 64), B = x patch
// [rows][64 ci] read `shift` rows further down (N = 32); D[m][n] = sum_k A[k][m] * B[k + shift][n] over K = 32.
// Dumps all 128 TMEM lanes so the host can see where the 64 accumulator rows of an M = 64 UMMA live.
__global__ void __launch_bounds__(128, 1) mn_probe_kernel(float* out, int shift, int m_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                  // 64 K-rows x 128 B, second 64-element chunk 8192 B further (M = 128 only)
  uint8_t* sb = smem + 16384;          // kRows rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + kRows * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 16384 / 2; i += 128) reinterpret_cast<__nv_bfloat16*>(sa)[i] = __float2bfloat16(0.f);
  __syncthreads();
  for (int i = tid; i < 64 * 128; i += 128) {
    const int k = i / 128, m = i % 128;
    const float v = static_cast<float>(((k * 128 + m) * 2654435761u) >> 28) - 8.f;
    const uint32_t abs = smem_u32(sa) + (m / 64) * 8192 + k * 128 + (m % 64) * 2;
    const uint32_t sw = abs ^ (((abs >> 7) & 7) << 4);
    *reinterpret_cast<__nv_bfloat16*>(sa + (sw - smem_u32(sa))) = __float2bfloat16(v);
  }
  for (int i = tid; i < kRows * 64; i += 128) {
    const int k = i / 64, n = i % 64;
    const float v = static_cast<float>(((k * 64 + n) * 2246822519u) >> 28) - 8.f;
    const uint32_t abs = smem_u32(sb) + k * 128 + n * 2;
    const uint32_t sw = abs ^ (((abs >> 7) & 7) << 4);
    *reinterpret_cast<__nv_bfloat16*>(sb + (sw - smem_u32(sb))) = __float2bfloat16(v);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(m_rows, 32, true, true);
    for (int k = 0; k < 2; ++k) {
      const uint32_t aa = smem_u32(sa) + k * 2048;
      const uint32_t bb = smem_u32(sb) + (shift + k * 16) * 128;
      umma_bf16(tmem, make_desc(aa, 8192, 1024, 0, 2), make_desc(bb, 8192, 1024, 0, 2), idesc, k != 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  float* d_out;
  cudaMalloc(&d_out, 128 * 64 * 4);
  const int smem = kRows * 128 + 64 * 128 + 64 + 1024;
  cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> h(128 * 64);
  const int shifts[] = {0, 8, 1, 3, 7, 58, 59, 117, 130, 131, 255};
  int all_ok[2][2] = {{1, 1}, {1, 1}};
  for (int mode = 0; mode < 2; ++mode) {
    const int kcols = mode == 0 ? 64 : 32;
    for (int ubo = 0; ubo < 2; ++ubo) {
      for (int shift : shifts) {
        cudaMemset(d_out, 0xff, 128 * 64 * 4);
        halo_probe_kernel<<<1, 128, smem>>>(d_out, shift, ubo, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d base_off %d shift %d: CUDA error %s\n", mode, ubo, shift, cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(h.data(), d_out, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int i = 0; i < 128; ++i)
          for (int j = 0; j < 64; ++j) {
            const float want = j < kcols ? static_cast<float>(((i + shift) * 3 + j * 5) % 251) - 125.f : 0.f;
            if (h[i * 64 + j] != want) {
              if (first < 0) first = i * 64 + j;
              ++bad;
            }
          }
        printf("mode %s base_off %s shift %3d: %s (%d mismatches%s)\n", mode == 0 ? "SW128" : "SW64 ",
               ubo ? "(addr>>7)&7" : "0          ", shift, bad ? "WRONG" : "ok", bad,
               bad ? "" : "");
        if (bad && first >= 0)
          printf("    first mismatch at row %d col %d: got %.1f\n", first / 64, first % 64, h[first]);
        if (bad) all_ok[mode][ubo] = 0;
      }
    }
  }
  printf("summary: SW128 base_off=0 %s | SW128 base_off=addr %s | SW64 base_off=0 %s | SW64 base_off=addr %s\n",
         all_ok[0][0] ? "OK" : "FAIL", all_ok[0][1] ? "OK" : "FAIL", all_ok[1][0] ? "OK" : "FAIL",
         all_ok[1][1] ? "OK" : "FAIL");
  // ---- probe 2
  {
    const int smem2 = 16384 + kRows * 128 + 64 + 1024;
    cudaFuncSetAttribute(mn_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    std::vector<float> h2(128 * 32);
    for (int m_rows : {128, 64}) {
      for (int shift : {0, 8, 1, 3, 59, 131, 262}) {
        cudaMemset(d_out, 0xff, 128 * 32 * 4);
        mn_probe_kernel<<<1, 128, smem2>>>(d_out, shift, m_rows);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mn probe M %d shift %d: CUDA error %s\n", m_rows, shift, cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(h2.data(), d_out, 128 * 32 * 4, cudaMemcpyDeviceToHost);
        // expected rows
        std::vector<float> want(128 * 32, 0.f);
        for (int m = 0; m < m_rows; ++m)
          for (int n = 0; n < 32; ++n) {
            float acc = 0.f;
            for (int k = 0; k < 32; ++k)
              acc += (static_cast<float>(((k * 128 + m) * 2654435761u) >> 28) - 8.f) * (static_cast<float>((((k + shift) * 64 + n) * 2246822519u) >> 28) - 8.f);
            want[m * 32 + n] = acc;
          }
        // where did row m land?
        int identity_ok = 1, found_all = 1;
        int lane_of[128];
        for (int m = 0; m < m_rows; ++m) {
          lane_of[m] = -1;
          for (int l = 0; l < 128; ++l) {
            bool same = true;
            for (int n = 0; n < 32 && same; ++n) same = h2[l * 32 + n] == want[m * 32 + n];
            if (same) { lane_of[m] = l; break; }
          }
          if (lane_of[m] != m) identity_ok = 0;
          if (lane_of[m] < 0) found_all = 0;
        }
        printf("mn probe M %3d shift %3d: rows found %s, identity lanes %s; lane of row 0/1/15/16/17/31/32/33/63: %d %d %d %d %d %d %d %d %d\n",
               m_rows, shift, found_all ? "ALL" : "NO", identity_ok ? "yes" : "no", lane_of[0], lane_of[1], lane_of[15],
               lane_of[16], lane_of[17], lane_of[31], lane_of[32], lane_of[33], lane_of[63]);
      }
    }
  }
  return 0;
}
