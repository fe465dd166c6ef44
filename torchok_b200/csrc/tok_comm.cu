// tok_comm.cu — data-parallel gradient exchange over NVLink/NVSwitch peer memory, fused with the optimizer step.
//
// Replaces the DDP all-reduce + optimizer.step pair that Lightning runs for `trainer.strategy: ddp`
// (torchok/constructor/config_structure.py:137-140, examples/configs/classification_imagenet.yaml:121-122) and the
// optimizer construction of torchok/constructor/constructor.py:86-160.  See include/tokb200.h (tok_peer_step).
//
// One kernel per (rank, gradient bucket):
//   phase A   every rank publishes "my gradients of bucket b, epoch e, are complete" into every peer's flag block
//             (st.release.sys over NVLink) and waits until all ranks have done so
//   reduce    rank r owns slice r of the bucket: 16-byte loads of that slice from all `world` gradient arenas
//             (world - 1 of them remote), summed in rank order on every rank -> identical results on all ranks
//   update    torch.optim.SGD / Adam(W) arithmetic on the slice (tok_optim.cuh), state local to the owner (ZeRO-1)
//   scatter   new fp32 master + bf16 shadow values stored into all `world` arenas (world - 1 remote)
//   phase B   the last CTA of the rank to finish publishes "my slice is written everywhere"; when all ranks have, the
//             local gradients of the bucket are cleared and the epoch advances
// HBM / NVLink bytes per arena element and rank: 4 (gradients in, (world-1)/world remote) + 6 (weights out) + state.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/tokb200.h"
#include "tok_internal.h"
#include "tok_optim.cuh"
#include "tok_ptx.cuh"

namespace tok {

constexpr int kPeerThreads = 512;
constexpr int kPeerBlocks = 48;   // all CTAs must be co-resident (they wait for each other through flags)
// flag block layout (32-bit words): [phase 2][bucket][rank] epochs, then per-bucket epoch and ticket words
constexpr int kFlagWords = 2 * TOK_PEER_MAX_BUCKETS * TOK_PEER_MAX_RANKS;
constexpr int kEpochOff = kFlagWords;
constexpr int kTicketOff = kFlagWords + TOK_PEER_MAX_BUCKETS;
constexpr int kFlagTotalWords = kFlagWords + 2 * TOK_PEER_MAX_BUCKETS;

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {   // streaming 16-byte load (no reuse)
  float4 v;
  asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// Bounded spin (about 10 s at 2 GHz): a peer that never arrives traps this context instead of hanging the GPU.
__device__ __forceinline__ void wait_epoch(const unsigned* flag, unsigned e) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - e) < 0) {
    __nanosleep(64);
    if (clock64() - t0 > 20000000000LL) __trap();
  }
}

struct PeerStepParams {
  tokPeerArenas ar;
  int bucket, kind;
  long long begin, end;   // elements
  float* state0;
  float* state1;
  const float* lr_dev;
  const int* step_dev;
  float h0, h1, h2, h3;
  int i0;
  float gscale;
  ParamSegs segs;
};

__global__ void __launch_bounds__(kPeerThreads, 1) peer_step_kernel(const PeerStepParams p) {
  const int world = p.ar.world, rank = p.ar.rank, b = p.bucket;
  unsigned* my_flags = p.ar.flags[rank];
  __shared__ unsigned s_last;
  const unsigned e = ld_acquire_sys(my_flags + kEpochOff + b) + 1u;   // advanced only after every CTA has passed phase B
  // ---- phase A
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(p.ar.flags[threadIdx.x] + (0 * TOK_PEER_MAX_BUCKETS + b) * TOK_PEER_MAX_RANKS + rank, e);
  }
  if (threadIdx.x < world) wait_epoch(my_flags + (0 * TOK_PEER_MAX_BUCKETS + b) * TOK_PEER_MAX_RANKS + threadIdx.x, e);
  __syncthreads();
  // ---- my slice of the bucket, in float4 units
  const long long len4 = (p.end - p.begin) >> 2;
  const long long per = (len4 + world - 1) / world;
  const long long lo = (p.begin >> 2) + per * rank;
  long long hi = lo + per;
  if (hi > (p.end >> 2)) hi = p.end >> 2;
  const float lr0 = __ldg(p.lr_dev);
  const int step = __ldg(p.step_dev);
  SgdArgs sa;
  AdamArgs aa;
  float bc1 = 1.f;
  if (p.kind == 0) {
    sa.lr = lr0; sa.mu = p.h0; sa.wd = p.h1; sa.damp = p.h2; sa.nesterov = p.i0; sa.gscale = p.gscale;
    sa.first = step <= 1; sa.zero_grad = 0;
  } else {
    const float t = (float)step;
    bc1 = 1.f - powf(p.h0, t);
    aa.lr = lr0; aa.b1 = p.h0; aa.b2 = p.h1; aa.eps = p.h2; aa.wd = p.h3; aa.decoupled = p.i0; aa.gscale = p.gscale;
    aa.step = lr0 / bc1;
    aa.rbc2 = rsqrtf(1.f - powf(p.h1, t));
  }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi; i += stride) {
    float4 g[TOK_PEER_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < TOK_PEER_MAX_RANKS; ++r)
      if (r < world) g[r] = ld_peer_f4(p.ar.grad[r] + (i << 2));
    float4 s = g[0];
#pragma unroll
    for (int r = 1; r < TOK_PEER_MAX_RANKS; ++r)
      if (r < world) { s.x += g[r].x; s.y += g[r].y; s.z += g[r].z; s.w += g[r].w; }
    bool skip = false;
    if (p.segs.n) {
      float lm, wm;
      const int sg = seg_lookup(p.segs, i << 2, lm, wm);
      skip = lm == 0.f && wm == 0.f;
      if (p.kind == 0) {
        sa.lr = lr0 * lm;
        sa.wd = p.h1 * wm;
      } else {
        aa.lr = lr0 * lm;
        aa.wd = p.h3 * wm;
        float c1 = bc1;
        if (p.segs.steps) {
          const float ts = (float)__ldg(p.segs.steps + sg);
          c1 = 1.f - powf(p.h0, ts);
          aa.rbc2 = rsqrtf(1.f - powf(p.h1, ts));
        }
        aa.step = aa.lr / c1;
      }
    }
    if (skip) continue;   // frozen parameter: identical on every rank already
    float4 w = reinterpret_cast<const float4*>(p.ar.master[rank])[i];
    if (p.kind == 0) {
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
      if (sa.mu != 0.f && !sa.first) m = reinterpret_cast<const float4*>(p.state0)[i];
      w.x = sgd_one(w.x, s.x, &m.x, sa);
      w.y = sgd_one(w.y, s.y, &m.y, sa);
      w.z = sgd_one(w.z, s.z, &m.z, sa);
      w.w = sgd_one(w.w, s.w, &m.w, sa);
      if (sa.mu != 0.f) reinterpret_cast<float4*>(p.state0)[i] = m;
    } else {
      float4 m = reinterpret_cast<const float4*>(p.state0)[i];
      float4 v = reinterpret_cast<const float4*>(p.state1)[i];
      w.x = adam_one(w.x, s.x, &m.x, &v.x, aa);
      w.y = adam_one(w.y, s.y, &m.y, &v.y, aa);
      w.z = adam_one(w.z, s.z, &m.z, &v.z, aa);
      w.w = adam_one(w.w, s.w, &m.w, &v.w, aa);
      reinterpret_cast<float4*>(p.state0)[i] = m;
      reinterpret_cast<float4*>(p.state1)[i] = v;
    }
    const uint2 wb = make_uint2(pack_bf16x2(w.x, w.y), pack_bf16x2(w.z, w.w));
#pragma unroll
    for (int r = 0; r < TOK_PEER_MAX_RANKS; ++r)
      if (r < world) {
        reinterpret_cast<float4*>(p.ar.master[r])[i] = w;
        reinterpret_cast<uint2*>(p.ar.shadow[r])[i] = wb;
      }
  }
  // ---- phase B: my slice is stored everywhere
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(my_flags + kTicketOff + b, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (s_last && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(p.ar.flags[threadIdx.x] + (1 * TOK_PEER_MAX_BUCKETS + b) * TOK_PEER_MAX_RANKS + rank, e);
  }
  if (threadIdx.x < world) wait_epoch(my_flags + (1 * TOK_PEER_MAX_BUCKETS + b) * TOK_PEER_MAX_RANKS + threadIdx.x, e);
  __syncthreads();
  // every rank has read my gradients and written my weights: clear the bucket's local gradients
  float4* gl = reinterpret_cast<float4*>(p.ar.grad[rank]);
  for (long long i = (p.begin >> 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (p.end >> 2); i += stride)
    gl[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s_last && threadIdx.x == 0) {   // every CTA of this rank has read the epoch (they all took a ticket)
    my_flags[kTicketOff + b] = 0u;
    st_release_sys(my_flags + kEpochOff + b, e);
  }
}

__global__ void peer_advance_kernel(int* step, int* seg_steps, const float* lr_mult, const float* wd_mult, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *step += 1;
  if (seg_steps && i < n && !(lr_mult[i] == 0.f && wd_mult[i] == 0.f)) seg_steps[i] += 1;
}

}  // namespace tok

using namespace tok;

extern "C" {

size_t tok_peer_flag_bytes(void) { return (size_t)kFlagTotalWords * 4; }

int tok_ipc_alloc(size_t bytes, void** dev_ptr, void* handle64) {
  if (!dev_ptr || !handle64 || bytes == 0) return set_error(TOK_ERR_INVALID, "ipc_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "ipc_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    cudaGetLastError();
    return set_error(TOK_ERR_CUDA, "ipc_alloc: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return TOK_OK;
}

int tok_ipc_free(void* dev_ptr) {
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "ipc_free: %s", cudaGetErrorString(e));
  return TOK_OK;
}

int tok_ipc_open(const void* handle64, void** dev_ptr) {
  if (!dev_ptr || !handle64) return set_error(TOK_ERR_INVALID, "ipc_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(TOK_ERR_CUDA, "ipc_open: %s", cudaGetErrorString(e));
  }
  *dev_ptr = p;
  return TOK_OK;
}

int tok_ipc_close(void* dev_ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "ipc_close: %s", cudaGetErrorString(e));
  return TOK_OK;
}

int tok_peer_step(const tokPeerArenas* arenas, int bucket, long long begin, long long end, int kind, float* state0,
                  float* state1, const float* lr_dev, int* step_dev, float h0, float h1, float h2, float h3, int i0,
                  float grad_scale, const int* seg_begin, const float* seg_lr_mult, const float* seg_wd_mult,
                  int* seg_steps, int n_segs, int advance_step, void* stream) {
  if (!arenas || arenas->world < 1 || arenas->world > TOK_PEER_MAX_RANKS || arenas->rank < 0 ||
      arenas->rank >= arenas->world)
    return set_error(TOK_ERR_INVALID, "peer_step: bad arena table");
  if (bucket < 0 || bucket >= TOK_PEER_MAX_BUCKETS) return set_error(TOK_ERR_INVALID, "peer_step: bucket index out of range");
  if (begin < 0 || end <= begin || (begin & 3) || (end & 3))
    return set_error(TOK_ERR_INVALID, "peer_step: [begin, end) must be a non-empty range of multiples of 4 elements");
  if (kind != 0 && kind != 1) return set_error(TOK_ERR_INVALID, "peer_step: kind must be 0 (SGD) or 1 (Adam)");
  if (!lr_dev || !step_dev) return set_error(TOK_ERR_INVALID, "peer_step: lr_dev and step_dev are required");
  if (kind == 1 && (!state0 || !state1)) return set_error(TOK_ERR_INVALID, "peer_step: Adam needs exp_avg and exp_avg_sq");
  if (kind == 0 && h0 != 0.f && !state0) return set_error(TOK_ERR_INVALID, "peer_step: SGD momentum needs its buffer");
  for (int r = 0; r < arenas->world; ++r)
    if (!arenas->master[r] || !arenas->grad[r] || !arenas->shadow[r] || !arenas->flags[r])
      return set_error(TOK_ERR_INVALID, "peer_step: missing arena pointer of rank %d", r);
  if (n_segs < 0 || (n_segs > 0 && (!seg_begin || !seg_lr_mult || !seg_wd_mult)))
    return set_error(TOK_ERR_INVALID, "peer_step: a segment table needs begin / lr_mult / wd_mult arrays");
  cudaStream_t st = (cudaStream_t)stream;
  if (advance_step) {
    const int n = n_segs > 0 ? n_segs : 1;
    peer_advance_kernel<<<(n + 255) / 256, 256, 0, st>>>(step_dev, n_segs > 0 ? seg_steps : nullptr, seg_lr_mult,
                                                       seg_wd_mult, n_segs);
  }
  PeerStepParams p;
  memset(&p, 0, sizeof(p));
  p.ar = *arenas;
  p.bucket = bucket;
  p.kind = kind;
  p.begin = begin;
  p.end = end;
  p.state0 = state0;
  p.state1 = state1;
  p.lr_dev = lr_dev;
  p.step_dev = step_dev;
  p.h0 = h0; p.h1 = h1; p.h2 = h2; p.h3 = h3; p.i0 = i0;
  p.gscale = grad_scale;
  p.segs.begin = seg_begin;
  p.segs.lr_mult = seg_lr_mult;
  p.segs.wd_mult = seg_wd_mult;
  p.segs.n = n_segs;
  p.segs.steps = (kind == 1 && n_segs > 0) ? seg_steps : nullptr;
  peer_step_kernel<<<kPeerBlocks, kPeerThreads, 0, st>>>(p);
  TOK_CHECK_LAUNCH("peer_step");
  return TOK_OK;
}

}  // extern "C"
