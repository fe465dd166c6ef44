// tok_api.cu — C ABI: argument checking, TMA tensor-map construction, kernel dispatch for the tensor-core ops.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/tokb200.h"
#include "tok_conv.cuh"
#include "tok_internal.h"

namespace tok {
cudaError_t launch_conv_fwd(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvFwdParams& p, int bn, bool b_mn,
                            cudaStream_t st);
cudaError_t launch_conv_fwd_persist(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                    const CUtensorMap& tmD, const ConvFwdParams& p, int bn, bool b_mn,
                                    cudaStream_t st);
// tok_conv2.cu: CTA-pair (cta_group::2) variant of the 128x256 persistent kernel, opt-in through TOK_CONV_2CTA=1
cudaError_t launch_conv_fwd_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                                 const CUtensorMap& tmD, const ConvFwdParams& p, bool b_mn, cudaStream_t st);
cudaError_t launch_conv_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const ConvWgradParams& p, int bn,
                              int splits, cudaStream_t st);
// tok_conv3.cu: halo formulation of the 3x3 / stride 1 / pad 1 convolution for Cin <= 128
bool conv3x3_halo_eligible(int n_img, int H, int W, int Cin, int N);
int launch_conv3x3_halo(const void* x, int n_img, int H, int W, int Cin, int N, const void* w, int wK, int wC,
                        int transposed, void* out, const void* addend, const void* addend_bits, float* col_sum,
                        float* col_sqsum, cudaStream_t st);
bool conv3x3_wgrad_halo_eligible(int n_img, int H, int W, int Cin, int Cout);
int launch_conv3x3_wgrad_halo(const void* x, const void* dy, int n_img, int H, int W, int Cin, int Cout, int wK, int wC,
                              float* dw, cudaStream_t st);
void launch_dilate_rows(const __nv_bfloat16* src, __nv_bfloat16* dst, int n, int p, int q, int c, int H, int W,
                        int sh, int sw, cudaStream_t st);

bool pdl_enabled() {
  static const bool on = !(getenv("TOK_PDL") && atoi(getenv("TOK_PDL")) == 0);
  return on;
}

static thread_local char g_err[512] = "";
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---- driver entry points (resolved lazily so the library loads on hosts without libcuda) -------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static long long* g_prof_buf = nullptr;
static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;

static int resolve_driver() {
  if (g_encode_tiled && g_encode_im2col) return TOK_OK;
  void* f1 = nullptr;
  void* f2 = nullptr;
  cudaDriverEntryPointQueryResult q1, q2;
  cudaError_t e1 = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f1, cudaEnableDefault, &q1);
  cudaError_t e2 = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f2, cudaEnableDefault, &q2);
  if (e1 != cudaSuccess || e2 != cudaSuccess || !f1 || !f2) {
    cudaGetLastError();
    return set_error(TOK_ERR_NODRIVER, "cuTensorMapEncode* entry points unavailable (%s)",
                     cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  }
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(f1);
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(f2);
  return TOK_OK;
}

// 2-D bf16 matrix [rows][cols] with row pitch ld (elements); box = (64 cols, box_rows), 128-byte swizzle.
int make_tmap_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  int rc = resolve_driver();
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15))
    return set_error(TOK_ERR_INVALID, "TMA operand must be 16-byte aligned with a 16-byte multiple pitch (ld=%lld)", ld);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(TOK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
  return TOK_OK;
}

// NHWC bf16 tensor walked in im2col mode: `pixels` consecutive conv-output positions x 64 channels per load.
// dims = (C, W, H, N) extents, strides = byte strides of W, H, N (explicit so that overlapping views are possible).
int make_tmap_im2col_ex(CUtensorMap* tm, const void* base, const cuuint64_t dims[4], const cuuint64_t strides[3],
                        const PixelSrc& s, int pixels) {
  int rc = resolve_driver();
  if (rc) return rc;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15))
    return set_error(TOK_ERR_INVALID, "im2col TMA operand needs 16-byte aligned base and strides");
  const int w = (int)dims[1], h = (int)dims[2];
  int lower[2] = {-s.pad, -s.pad};
  int upper[2] = {(-s.pad + (s.Q - 1) * s.stride) - (w - 1), (-s.pad + (s.P - 1) * s.stride) - (h - 1)};
  for (int i = 0; i < 2; ++i)
    if (lower[i] < -128 || lower[i] > 127 || upper[i] < -128 || upper[i] > 127)
      return set_error(TOK_ERR_INVALID, "im2col corner out of range (lower %d upper %d)", lower[i], upper[i]);
  cuuint32_t estr[4] = {1, (cuuint32_t)s.stride, (cuuint32_t)s.stride, 1};
  CUresult r = g_encode_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower,
                               upper, 64, (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(TOK_ERR_CUDA,
                     "cuTensorMapEncodeIm2col failed (%d) dims=%llu,%llu,%llu,%llu strides=%llu,%llu,%llu lower=%d,%d "
                     "upper=%d,%d stride=%d",
                     (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                     (unsigned long long)dims[3], (unsigned long long)strides[0], (unsigned long long)strides[1],
                     (unsigned long long)strides[2], lower[0], lower[1], upper[0], upper[1], s.stride);
  return TOK_OK;
}
int make_tmap_im2col(CUtensorMap* tm, const void* base, int n, int h, int w, int c, const PixelSrc& s, int pixels) {
  if ((c * 2) & 15) return set_error(TOK_ERR_INVALID, "im2col TMA operand needs C %% 8 == 0 (C=%d)", c);
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  return make_tmap_im2col_ex(tm, base, dims, strides, s, pixels);
}

static int check_desc(const tokConvDesc* d) {
  if (!d) return set_error(TOK_ERR_INVALID, "null conv descriptor");
  if (d->n <= 0 || d->h <= 0 || d->w <= 0 || d->c <= 0 || d->k <= 0 || d->r <= 0 || d->s <= 0 || d->stride <= 0 ||
      d->dil <= 0 || d->pad < 0)
    return set_error(TOK_ERR_INVALID, "bad conv descriptor");
  if ((d->c % 8) || (d->k % 8)) return set_error(TOK_ERR_INVALID, "channel counts must be multiples of 8 (c=%d k=%d)", d->c, d->k);
  if (d->wk < 0 || d->wc < 0 || d->wk > d->k || d->wc > d->c)
    return set_error(TOK_ERR_INVALID, "weight dimensions wk=%d wc=%d exceed the activation pitches k=%d c=%d", d->wk, d->wc,
                     d->k, d->c);
  return TOK_OK;
}
static inline int desc_wk(const tokConvDesc* d) { return d->wk ? d->wk : d->k; }
static inline int desc_wc(const tokConvDesc* d) { return d->wc ? d->wc : d->c; }
static inline bool desc_unpadded(const tokConvDesc* d) { return desc_wk(d) != d->k || desc_wc(d) != d->c; }
static inline bool desc_halo_shape(const tokConvDesc* d) {
  return d->r == 3 && d->s == 3 && d->stride == 1 && d->pad == 1 && d->dil == 1 && !getenv("TOK_CONV_V1");
}

static int pick_bn(int n) { return n <= 64 ? 64 : 128; }

// Persistent kernel: a 128x256 tile halves the L2->SM operand traffic per FLOP (the 128x128 tile is L2-bandwidth
// bound at ~1/3 of the tensor peak); it is chosen when it does not cost wave-quantisation efficiency on 148 SMs.
static int pick_bn_persist(long long M, int N, long long k_total, bool gather_a) {
  if (N <= 64) return 64;
  const int forced = getenv("TOK_CONV_BN") ? atoi(getenv("TOK_CONV_BN")) : 0;
  if (forced == 128 || (forced == 256 && N > 128)) return forced;
  if (N <= 128) return 128;
  // Short reductions are bound by the epilogue / HBM, not by the tensor pipe: the 128x128 tile double-buffers its
  // staging (and addend) tiles so stores drain behind the next tile, which the 128x256 tile has no room for.  Exception:
  // a strided (im2col-gathered) A operand is expensive to fetch, and the wider tile fetches it half as often
  // (measured r2, 1x1 s2 256->512 @56: 91 us vs 109 us).
  // r5 (scripts/linear_shapes.py, Swin stage 3: M = 50176, K = 384): with N a multiple of 256 the wide tile already wins at
  // K = 384 (fc1 forward 95 -> 89 us, fc2 dgrad 100 -> 78 us), so such layers go on to the wave model below; ragged N at
  // that depth does not (qkv N = 1152: 72 -> 76 us).
  static const long long small_k = getenv("TOK_CONV_SMALLK") ? atoll(getenv("TOK_CONV_SMALLK")) : 512;
  if (k_total < small_k && !gather_a && !(k_total >= 384 && N % 256 == 0)) return 128;
  // Waves of the persistent grid x relative tile time.  A 128x256 tile moves 1.5x the operand bytes of a 128x128 tile for
  // 2x the MACs and measures ~1.25x its time on the long reductions (r2 selftest: 3x3 512->512 @7, 196 tiles in 2 waves
  // = 69 us against 392 tiles in 3 waves = 95 us).
  const long long m_tiles = (M + 127) / 128;
  auto cost = [&](int bn) {
    const long long tiles = m_tiles * ((N + bn - 1) / bn);
    const long long waves = (tiles + 147) / 148;
    return (double)waves * (bn == 256 ? 1.25 : 1.0);
  };
  if (N % 256 == 0) return cost(256) <= cost(128) ? 256 : 128;
  // ragged N: the padded half tile of the wide configuration is pure waste — keep the r1 efficiency rule
  auto eff = [&](int bn) {
    const long long tiles = m_tiles * ((N + bn - 1) / bn);
    const long long waves = (tiles + 147) / 148;
    const double useful = (double)N / (((N + bn - 1) / bn) * bn);
    return useful * (double)tiles / (double)(waves * 148);
  };
  return eff(256) >= 0.85 * eff(128) ? 256 : 128;
}

// Generic "pixels x weights" launch used by fprop, dgrad and linear.
static int run_fwd_tm(const CUtensorMap& tmA, int ac, const PixelSrc& src, long long M, const void* wmat,
                      long long w_rows, long long w_cols, bool b_mn, int N, int flip, ConvFwdParams p,
                      cudaStream_t st) {
  CUtensorMap tmB;
  static const bool v1 = getenv("TOK_CONV_V1") != nullptr;  // bring-up aid: the one-tile-per-CTA kernel
  int bn = v1 ? pick_bn(N) : pick_bn_persist(M, N, (long long)src.R * src.S * ac, src.im2col && src.R * src.S == 1);
  {
    // 192-wide strips when they tile N with no more padding than 128-wide ones and N is not a multiple of 256: plain
    // GEMM launches only (no BatchNorm sums, no addend, no scatter) — the transformer linear layers.  TOK_CONV_BN192=0: off.
    static const bool bn192 = !(getenv("TOK_CONV_BN192") && atoi(getenv("TOK_CONV_BN192")) == 0);
    static const bool forced = getenv("TOK_CONV_BN") != nullptr;
    if (bn192 && !forced && !v1 && N > 128 && (N % 256) != 0 && p.col_sum == nullptr && p.addend == nullptr && !p.scatter &&
        ((N + 191) / 192) * 192 <= ((N + 127) / 128) * 128)
      bn = 192;
  }
  if (p.addend_mode == 1 && !v1) bn = 128;   // the GELU-backward epilogue exists for the 128-wide tile only
  // Opt-in CTA-pair kernel (unverified on hardware as a conv; the default path is untouched unless the variable is
  // set): a 256x256 tile per pair of SMs, each CTA fetches half of the weight tile, hence the 128-row boxes.
  static const bool pair_env = getenv("TOK_CONV_2CTA") != nullptr;
  const bool pair = pair_env && !v1 && bn == 256 && !p.scatter && (M % 256) == 0 && (N % 256) == 0;
  int rc = make_tmap_2d(&tmB, wmat, w_rows, w_cols, w_cols, b_mn ? 64 : (pair ? 128 : bn));
  if (rc) return rc;
  p.M = (int)M;
  p.N = N;
  p.Cin = ac;
  p.a = src;
  p.flip_taps = flip;
  mn_desc_geometry(&p.mn_lbo, &p.mn_sbo, &p.mn_kadv);
  static const int defer = getenv("TOK_CONV_DEFER_STATS") ? atoi(getenv("TOK_CONV_DEFER_STATS")) : 1;
  p.defer_stats = defer;
  // Tile order of the persistent kernel (tok_conv.cu: decode_tile).  Several column strips over an A operand that does
  // not stay in L2 between strips: walk groups of m-tiles whose A rows total <= 16 MB.  Measured r5 (one B200, A/B in one
  // box): Swin-T linear fwd 4.64 -> 4.46 ms, linear dgrad 4.40 -> 4.19 ms, step 27.36 -> 26.62 ms; ResNet-50 conv dgrad
  // 4.08 -> 3.95 ms.  With BatchNorm sums a group change costs a flush of the register partials and the forward convs
  // measured no gain (3.78 vs 3.80 ms), so launches with sums keep whole strips.  TOK_CONV_MGROUP=0: off, =n: n x 148
  // m-tiles per group for every launch.
  {
    static const int mg_env = getenv("TOK_CONV_MGROUP") ? atoi(getenv("TOK_CONV_MGROUP")) : -1;
    const long long a_bytes = M * (long long)ac * 2;
    const int strips = (N + bn - 1) / bn;
    p.m_group = 0;
    if (mg_env != 0 && strips > 1 && a_bytes > (24LL << 20)) {
      long long k = mg_env > 0 ? mg_env : (16LL << 20) / (148LL * 128 * ac * 2);
      if (k < 1) k = 1;
      // (the fused GELU-backward dgrad also carries sums, but its A operand — 154 MB at Swin stage 1, read once per
      // 128-column strip of the 4x wider output — is exactly the case the grouping is for)
      if (p.col_sum == nullptr || mg_env > 0 || (p.addend_mode == 1 && k >= 2)) p.m_group = (int)(148 * k);
    }
  }
  // TOK_CONV_PROFILE=1: the epilogue phase counters of every launch land in a static device buffer which
  // tok_debug_conv_profile() copies out (bring-up aid; not part of the production path)
  static const bool want_prof = getenv("TOK_CONV_PROFILE") != nullptr;
  if (want_prof && !v1) {
    if (!g_prof_buf) cudaMalloc(&g_prof_buf, 148 * 16 * sizeof(long long));
    cudaMemsetAsync(g_prof_buf, 0, 148 * 16 * sizeof(long long), st);
    p.prof = g_prof_buf;
  }
  cudaError_t e;
  if (v1) {
    e = launch_conv_fwd(tmA, tmB, p, bn, b_mn, st);
  } else {
    // output tile store: [M][N] matrix with pitch ldo, 64-column x 128-row boxes (unused by the scatter path)
    // the optional addend tile is fetched through a map of the same geometry (unused by the scatter path)
    CUtensorMap tmC = tmA, tmD = tmA;
    if (!p.scatter) {
      rc = make_tmap_2d(&tmC, p.out, M, N, p.ldo, 128);
      if (rc) return rc;
      if (p.addend) {
        rc = make_tmap_2d(&tmD, p.addend, M, N, p.ldo, 128);
        if (rc) return rc;
      }
    }
    e = pair ? launch_conv_fwd_pair(tmA, tmB, tmC, tmD, p, b_mn, st)
             : launch_conv_fwd_persist(tmA, tmB, tmC, tmD, p, bn, b_mn, st);
  }
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "conv_fwd launch: %s", cudaGetErrorString(e));
  return TOK_OK;
}
static int run_fwd(const void* a, int an, int ah, int aw, int ac, const PixelSrc& src, long long M, const void* wmat,
                   long long w_rows, long long w_cols, bool b_mn, int N, int flip, ConvFwdParams p, cudaStream_t st) {
  CUtensorMap tmA;
  int rc;
  if (src.im2col)
    rc = make_tmap_im2col(&tmA, a, an, ah, aw, ac, src, 128);
  else
    rc = make_tmap_2d(&tmA, a, M, ac, ac, 128);
  if (rc) return rc;
  return run_fwd_tm(tmA, ac, src, M, wmat, w_rows, w_cols, b_mn, N, flip, p, st);
}

static int pick_splits(long long tiles, int total_chunks, int* chunks_per_split) {
  // ONE wave: tiles x splits must not exceed the resident CTAs (2 per SM; 1 per SM when the tiles alone fill half the
  // chip, since every extra split costs a full tile of fp32 reductions), and at least 8 K-blocks per CTA.
  static const int forced = getenv("TOK_WGRAD_CTAS_PER_SM") ? atoi(getenv("TOK_WGRAD_CTAS_PER_SM")) : 0;
  const int per_sm = forced > 0 ? forced : (tiles >= 74 ? 1 : 2);
  const long long capacity = 148LL * per_sm;
  int splits = (int)(capacity / tiles);  // floor: never spill into a second wave
  int max_splits = (total_chunks + 7) / 8;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int cps = (total_chunks + splits - 1) / splits;
  splits = (total_chunks + cps - 1) / cps;
  *chunks_per_split = cps;
  return splits;
}

static int run_wgrad_tm(const CUtensorMap& tmX, int xc, const PixelSrc& src, const void* dy, long long Mpix, int Cout,
                        float* dw, cudaStream_t st) {
  CUtensorMap tmDY;
  int rc = make_tmap_2d(&tmDY, dy, Mpix, Cout, Cout, 64);
  if (rc) return rc;
  ConvWgradParams p;
  memset(&p, 0, sizeof(p));
  p.Mpix = (int)Mpix;
  p.Cout = Cout;
  p.Cin = xc;
  p.x = src;
  p.dw = dw;
  p.ldw = (long long)src.R * src.S * xc;
  mn_desc_geometry(&p.mn_lbo, &p.mn_sbo, &p.mn_kadv);
  const int bn = pick_bn(xc);
  const long long tiles = (long long)src.R * src.S * ((Cout + 127) / 128) * ((xc + bn - 1) / bn);
  const int total_chunks = (int)((Mpix + 63) / 64);
  const int splits = pick_splits(tiles, total_chunks, &p.chunks_per_split);
  cudaError_t e = launch_conv_wgrad(tmDY, tmX, p, bn, splits, st);
  if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "conv_wgrad launch: %s", cudaGetErrorString(e));
  return TOK_OK;
}
static int run_wgrad(const void* x, int xn, int xh, int xw, int xc, const PixelSrc& src, const void* dy, long long Mpix,
                     int Cout, float* dw, cudaStream_t st) {
  CUtensorMap tmX;
  int rc;
  if (src.im2col)
    rc = make_tmap_im2col(&tmX, x, xn, xh, xw, xc, src, 64);
  else
    rc = make_tmap_2d(&tmX, x, Mpix, xc, xc, 64);
  if (rc) return rc;
  return run_wgrad_tm(tmX, xc, src, dy, Mpix, Cout, dw, st);
}

// ---- ResNet stem (7x7 stride 2 pad 3, <=4 input channels) on the 2x2 space-to-depth packed input ---------------
// xs2d is [N][H2][W2][16]; four horizontally adjacent packed pixels are one 64-element im2col row, so the tensor map
// is an OVERLAPPING view: dims (64, Q, H2, N) with a W stride of one packed pixel (32 bytes).
static void stem_geom(int h, int w, int* P, int* Q, int* H2, int* W2) {
  *P = (h + 6 - 7) / 2 + 1;
  *Q = (w + 6 - 7) / 2 + 1;
  *H2 = *P + 3;
  *W2 = *Q + 3;
}
static int stem_tmap(CUtensorMap* tm, const void* xs2d, int n, int h, int w, PixelSrc* src, int pixels) {
  int P, Q, H2, W2;
  stem_geom(h, w, &P, &Q, &H2, &W2);
  memset(src, 0, sizeof(*src));
  src->im2col = 1;
  src->P = P;
  src->Q = Q;
  src->stride = 1;
  src->pad = 0;
  src->dil = 1;
  src->R = 4;
  src->S = 1;
  cuuint64_t dims[4] = {64, (cuuint64_t)Q, (cuuint64_t)H2, (cuuint64_t)n};
  cuuint64_t strides[3] = {32, (cuuint64_t)W2 * 32, (cuuint64_t)H2 * W2 * 32};
  return make_tmap_im2col_ex(tm, xs2d, dims, strides, *src, pixels);
}

static PixelSrc conv_src(const tokConvDesc* d, int P, int Q) {
  PixelSrc s;
  s.im2col = !(d->r == 1 && d->s == 1 && d->stride == 1 && d->pad == 0);
  s.P = P;
  s.Q = Q;
  s.stride = d->stride;
  s.pad = d->pad;
  s.dil = d->dil;
  s.R = d->r;
  s.S = d->s;
  return s;
}

}  // namespace tok

using namespace tok;

extern "C" {

int tok_version(void) { return 1; }
const char* tok_last_error(void) { return g_err; }

int tok_device_ok(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_error(TOK_ERR_CUDA, "no CUDA device (%s)", cudaGetErrorString(e));
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return set_error(TOK_ERR_INVALID, "device compute capability %d.x is not sm_100", major);
  return resolve_driver();
}

int tok_debug_conv_profile(long long* host_out, int max_entries) {
  if (!g_prof_buf) return 0;
  const int n = max_entries < 148 * 16 ? max_entries : 148 * 16;
  cudaDeviceSynchronize();
  cudaMemcpy(host_out, g_prof_buf, n * sizeof(long long), cudaMemcpyDeviceToHost);
  return n;
}

int tok_conv_halo_caps(const tokConvDesc* d) {
  if (!d || check_desc(d) != TOK_OK || !desc_halo_shape(d)) return 0;
  int caps = 0;
  if (conv3x3_halo_eligible(d->n, d->h, d->w, d->c, d->k)) caps |= 1;
  if (conv3x3_halo_eligible(d->n, d->h, d->w, d->k, d->c)) caps |= 2;
  if (conv3x3_wgrad_halo_eligible(d->n, d->h, d->w, d->c, d->k)) caps |= 4;
  return caps;
}

void tok_conv_out_hw(const tokConvDesc* d, int* p, int* q) {
  *p = (d->h + 2 * d->pad - d->dil * (d->r - 1) - 1) / d->stride + 1;
  *q = (d->w + 2 * d->pad - d->dil * (d->s - 1) - 1) / d->stride + 1;
}

static int conv_fprop_impl(const tokConvDesc* d, const void* x, const void* w, void* y, float* sum, float* sqsum,
                           const void* addend, const float* bias, int relu, const FwdFin* fin, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  int P, Q;
  tok_conv_out_hw(d, &P, &Q);
  if (P <= 0 || Q <= 0) return set_error(TOK_ERR_INVALID, "empty conv output");
  if ((sum == nullptr) != (sqsum == nullptr)) return set_error(TOK_ERR_INVALID, "sum and sqsum must be given together");
  PixelSrc src = conv_src(d, P, Q);
  ConvFwdParams p;
  memset(&p, 0, sizeof(p));
  p.out = static_cast<__nv_bfloat16*>(y);
  p.ldo = d->k;
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.col_sum = sum;
  p.col_sqsum = sqsum;
  p.bias = bias;
  p.relu = relu;
  const long long M = (long long)d->n * P * Q;
  if (desc_halo_shape(d) && !addend && !bias && !relu && !fin && conv3x3_halo_eligible(d->n, d->h, d->w, d->c, d->k))
    return launch_conv3x3_halo(x, d->n, d->h, d->w, d->c, d->k, w, desc_wk(d), desc_wc(d), 0, y, nullptr, nullptr, sum, sqsum,
                               static_cast<cudaStream_t>(stream));
  if (desc_unpadded(d)) return set_error(TOK_ERR_INVALID, "conv_fprop: unpadded weights (wk / wc) need the halo 3x3 path");
  if (fin) {
    p.fin = *fin;
    p.fin.count = (float)M;
  }
  return run_fwd(x, d->n, d->h, d->w, d->c, src, M, w, d->k, (long long)d->r * d->s * d->c, false, d->k, 0, p,
                 static_cast<cudaStream_t>(stream));
}

int tok_conv_fprop(const tokConvDesc* d, const void* x, const void* w, void* y, float* sum, float* sqsum,
                   const void* addend, const float* bias, int relu, void* stream) {
  return conv_fprop_impl(d, x, w, y, sum, sqsum, addend, bias, relu, nullptr, stream);
}

int tok_conv_fprop_bn(const tokConvDesc* d, const void* x, const void* w, void* y, float* sum, float* sqsum,
                      const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                      float* running_var, float* scale, float* shift, float* save_mean, float* save_invstd,
                      unsigned* counter, void* stream) {
  if (!sum || !sqsum || !scale || !shift || !save_mean || !save_invstd || !counter)
    return set_error(TOK_ERR_INVALID, "conv_fprop_bn: accumulators, outputs and the ticket counter are required");
  static const bool v1 = getenv("TOK_CONV_V1") != nullptr;
  if (v1) {   // the one-tile-per-CTA bring-up kernel has no fused finalize
    int rc = conv_fprop_impl(d, x, w, y, sum, sqsum, nullptr, nullptr, 0, nullptr, stream);
    if (rc) return rc;
    int P, Q;
    tok_conv_out_hw(d, &P, &Q);
    return tok_bn_finalize_train(d->k, (double)d->n * P * Q, sum, sqsum, gamma, beta, eps, momentum, running_mean,
                                 running_var, scale, shift, save_mean, save_invstd, stream);
  }
  FwdFin fin;
  fin.counter = counter;
  fin.count = 0.f;
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
  return conv_fprop_impl(d, x, w, y, sum, sqsum, nullptr, nullptr, 0, &fin, stream);
}

// Strided RxS data gradient by OUTPUT PARITY CLASS: dx[s*p'+a, s*q'+b] only receives filter taps r with
// (a + pad - r) % stride == 0 (same along w), read from dy rows p' + (a + pad - r) / stride — consecutive offsets, so
// each of the stride^2 classes is a small stride-1 correlation over the COMPACT gradient (TMA im2col walk, explicit tap
// table into the weight matrix) whose rows are scattered to the class's positions of dx.  For 3x3 / stride 2 / pad 1 the
// classes have 1, 2, 2 and 4 taps: 9 tap-GEMMs over M/4 rows each instead of 9 over M rows of a zero-dilated copy
// (r1: dilate + memset + 4x the MMA work; 374 us for 128 -> 128 @56 against 94 us for the forward conv).
struct ParityAxis {
  int taps, omin;      // number of taps of the class, dy offset of local tap 0
  int r_of_tap[8];     // filter index of local tap t (dy offset omin + t)
};
static bool parity_axis(int a, int R, int stride, int pad, ParityAxis* ax) {
  ax->taps = 0;
  int omax = -(1 << 30), omin = 1 << 30;
  for (int r = 0; r < R; ++r) {
    const int v = a + pad - r;
    if (v % stride != 0) continue;
    const int o = v / stride;
    if (o > omax) omax = o;
    if (o < omin) omin = o;
  }
  if (omax < omin) return false;   // class without taps: that part of dx is zero (not handled here)
  ax->omin = omin;
  ax->taps = omax - omin + 1;
  if (ax->taps > 4) return false;
  for (int t = 0; t < ax->taps; ++t) ax->r_of_tap[t] = a + pad - stride * (omin + t);
  return true;
}
static bool strided_dgrad_by_parity(const tokConvDesc* d) {
  static const bool off = getenv("TOK_DGRAD_DILATE") != nullptr;   // A/B aid: the r1 zero-dilation path
  if (off || getenv("TOK_CONV_V1") || d->dil != 1 || d->stride > 4) return false;
  ParityAxis ax;
  for (int a = 0; a < d->stride; ++a)
    if (!parity_axis(a, d->r, d->stride, d->pad, &ax) || !parity_axis(a, d->s, d->stride, d->pad, &ax)) return false;
  return true;
}

size_t tok_conv_dgrad_workspace_bytes(const tokConvDesc* d) {
  if (!d) return 0;
  if (d->stride > 1 && !(d->r == 1 && d->s == 1) && !strided_dgrad_by_parity(d)) return (size_t)d->n * d->h * d->w * d->k * 2;
  return 0;
}

// The masked addend needs 32-channel words of the bit mask per (row, 32-column chunk): N % 32 == 0 on the generic kernel,
// any multiple of 8 on the halo kernel; the scatter forms (strided 1x1 / parity classes) have no addend path for it.
int tok_conv_dgrad_masked_supported(const tokConvDesc* d) {
  if (!d || check_desc(d) != TOK_OK || d->stride != 1 || getenv("TOK_CONV_V1")) return 0;
  if (desc_halo_shape(d) && conv3x3_halo_eligible(d->n, d->h, d->w, d->k, d->c)) return 1;
  if (desc_unpadded(d)) return 0;
  if (d->r == 1 && d->s == 1) return d->pad == 0 && (d->c % 32) == 0;
  return (d->c % 32) == 0 && ((d->r - 1) * d->dil - d->pad) == ((d->s - 1) * d->dil - d->pad) &&
         ((d->r - 1) * d->dil - d->pad) >= 0;
}

static int conv_dgrad_impl(const tokConvDesc* d, const void* dy, const void* w, void* dx, const void* addend,
                           const void* addend_bits, void* ws, void* stream);

int tok_conv_dgrad(const tokConvDesc* d, const void* dy, const void* w, void* dx, const void* addend, void* ws,
                   void* stream) {
  return conv_dgrad_impl(d, dy, w, dx, addend, nullptr, ws, stream);
}

int tok_conv_dgrad_masked(const tokConvDesc* d, const void* dy, const void* w, void* dx, const void* addend,
                          const void* addend_bits, void* ws, void* stream) {
  if (!addend || !addend_bits) return set_error(TOK_ERR_INVALID, "conv_dgrad_masked: addend and its bit mask are required");
  if (!tok_conv_dgrad_masked_supported(d))
    return set_error(TOK_ERR_INVALID, "conv_dgrad_masked: unsupported convolution (see tok_conv_dgrad_masked_supported)");
  return conv_dgrad_impl(d, dy, w, dx, addend, addend_bits, ws, stream);
}

static int conv_dgrad_impl(const tokConvDesc* d, const void* dy, const void* w, void* dx, const void* addend,
                           const void* addend_bits, void* ws, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  int P, Q;
  tok_conv_out_hw(d, &P, &Q);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ConvFwdParams p;
  memset(&p, 0, sizeof(p));
  p.out = static_cast<__nv_bfloat16*>(dx);
  p.ldo = d->c;
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  p.addend_bits = static_cast<const uint8_t*>(addend_bits);
  const long long wcols = (long long)d->r * d->s * d->c;
  if (d->r == 1 && d->s == 1) {
    if (d->pad != 0) return set_error(TOK_ERR_INVALID, "1x1 conv with padding is not supported");
    PixelSrc src;
    memset(&src, 0, sizeof(src));
    src.R = src.S = 1;
    src.stride = 1;
    src.dil = 1;
    if (d->stride > 1) {
      // GEMM over the P*Q output pixels, rows scattered to the (stride*p, stride*q) positions of dx.
      if (addend != dx) {
        cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)d->n * d->h * d->w * d->c * 2, st);
        if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
        if (addend != nullptr) return set_error(TOK_ERR_INVALID, "strided 1x1 dgrad: addend must be NULL or alias dx");
      }
      p.scatter = 1;
      p.sc_P = P;
      p.sc_Q = Q;
      p.sc_H = d->h;
      p.sc_W = d->w;
      p.sc_sh = d->stride;
      p.sc_sw = d->stride;
    }
    return run_fwd(dy, d->n, P, Q, d->k, src, (long long)d->n * P * Q, w, d->k, wcols, true, d->c, 0, p, st);
  }
  if (desc_halo_shape(d) && conv3x3_halo_eligible(d->n, d->h, d->w, d->k, d->c))
    return launch_conv3x3_halo(dy, d->n, d->h, d->w, d->k, d->c, w, desc_wk(d), desc_wc(d), 1, dx, addend, addend_bits,
                               nullptr, nullptr, st);
  if (desc_unpadded(d)) return set_error(TOK_ERR_INVALID, "conv_dgrad: unpadded weights (wk / wc) need the halo 3x3 path");
  // RxS filter: stride-1 correlation of (zero-dilated) dy with the flipped filter.
  const int pad_h = (d->r - 1) * d->dil - d->pad;
  const int pad_w = (d->s - 1) * d->dil - d->pad;
  if (pad_h != pad_w || pad_h < 0) return set_error(TOK_ERR_INVALID, "unsupported padding for dgrad");
  const void* src_ptr = dy;
  int sh = P, sw = Q;
  if (d->stride > 1 && strided_dgrad_by_parity(d)) {
    if (addend != nullptr && addend != dx)
      return set_error(TOK_ERR_INVALID, "strided dgrad: addend must be NULL or alias dx");
    for (int a = 0; a < d->stride; ++a) {
      for (int b = 0; b < d->stride; ++b) {
        ParityAxis ah, aw;
        parity_axis(a, d->r, d->stride, d->pad, &ah);
        parity_axis(b, d->s, d->stride, d->pad, &aw);
        const int Pc = (d->h - a + d->stride - 1) / d->stride, Qc = (d->w - b + d->stride - 1) / d->stride;
        if (Pc <= 0 || Qc <= 0) continue;
        PixelSrc src;
        src.im2col = 1;
        src.P = Pc;
        src.Q = Qc;
        src.stride = 1;
        src.dil = 1;
        src.R = ah.taps;
        src.S = aw.taps;
        // one lower corner for both axes: PixelSrc has a single pad, so shift the taller axis' tap window instead
        if (ah.omin != aw.omin)
          return set_error(TOK_ERR_INVALID, "strided dgrad: asymmetric tap offsets are not supported");
        src.pad = -ah.omin;
        ConvFwdParams q = p;
        q.scatter = 1;
        q.sc_P = Pc;
        q.sc_Q = Qc;
        q.sc_H = d->h;
        q.sc_W = d->w;
        q.sc_sh = d->stride;
        q.sc_sw = d->stride;
        q.sc_oh = a;
        q.sc_ow = b;
        q.use_tapmap = 1;
        for (int t = 0; t < ah.taps; ++t)
          for (int u = 0; u < aw.taps; ++u)
            q.tapmap[t * aw.taps + u] = (signed char)(ah.r_of_tap[t] * d->s + aw.r_of_tap[u]);
        rc = run_fwd(dy, d->n, P, Q, d->k, src, (long long)d->n * Pc * Qc, w, d->k, wcols, true, d->c, 0, q, st);
        if (rc) return rc;
      }
    }
    return TOK_OK;
  }
  if (d->stride > 1) {
    if (!ws) return set_error(TOK_ERR_INVALID, "dgrad workspace required");
    cudaError_t e = cudaMemsetAsync(ws, 0, tok_conv_dgrad_workspace_bytes(d), st);
    if (e != cudaSuccess) return set_error(TOK_ERR_CUDA, "memset: %s", cudaGetErrorString(e));
    launch_dilate_rows(static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(ws), d->n, P, Q, d->k, d->h,
                       d->w, d->stride, d->stride, st);
    src_ptr = ws;
    sh = d->h;
    sw = d->w;
  }
  PixelSrc src;
  src.im2col = 1;
  src.P = d->h;
  src.Q = d->w;
  src.stride = 1;
  src.pad = pad_h;
  src.dil = d->dil;
  src.R = d->r;
  src.S = d->s;
  return run_fwd(src_ptr, d->n, sh, sw, d->k, src, (long long)d->n * d->h * d->w, w, d->k, wcols, true, d->c, 1, p, st);
}

int tok_conv_wgrad(const tokConvDesc* d, const void* x, const void* dy, float* dw, void* stream) {
  int rc = check_desc(d);
  if (rc) return rc;
  int P, Q;
  tok_conv_out_hw(d, &P, &Q);
  if (desc_halo_shape(d) && conv3x3_wgrad_halo_eligible(d->n, d->h, d->w, d->c, d->k))
    return launch_conv3x3_wgrad_halo(x, dy, d->n, d->h, d->w, d->c, d->k, desc_wk(d), desc_wc(d), dw,
                                     static_cast<cudaStream_t>(stream));
  if (desc_unpadded(d)) return set_error(TOK_ERR_INVALID, "conv_wgrad: unpadded weights (wk / wc) need the halo 3x3 path");
  PixelSrc src = conv_src(d, P, Q);
  return run_wgrad(x, d->n, d->h, d->w, d->c, src, dy, (long long)d->n * P * Q, d->k, dw,
                   static_cast<cudaStream_t>(stream));
}

static PixelSrc flat_src() {
  PixelSrc s;
  memset(&s, 0, sizeof(s));
  s.R = s.S = 1;
  s.stride = 1;
  s.dil = 1;
  return s;
}

void tok_stem_geometry(int h, int w, int* p, int* q, int* h2, int* w2) { stem_geom(h, w, p, q, h2, w2); }

int tok_stem_conv_fprop(int n, int h, int w, int k, const void* xs2d, const void* wp, void* y, float* sum,
                        float* sqsum, void* stream) {
  if (n <= 0 || h <= 0 || w <= 0 || k <= 0 || (k % 8)) return set_error(TOK_ERR_INVALID, "stem: bad shape");
  CUtensorMap tmA;
  PixelSrc src;
  int rc = stem_tmap(&tmA, xs2d, n, h, w, &src, 128);
  if (rc) return rc;
  ConvFwdParams p;
  memset(&p, 0, sizeof(p));
  p.out = static_cast<__nv_bfloat16*>(y);
  p.ldo = k;
  p.col_sum = sum;
  p.col_sqsum = sqsum;
  return run_fwd_tm(tmA, 64, src, (long long)n * src.P * src.Q, wp, k, 256, false, k, 0, p,
                    static_cast<cudaStream_t>(stream));
}

int tok_stem_conv_wgrad(int n, int h, int w, int k, const void* xs2d, const void* dy, float* dwp, void* stream) {
  if (n <= 0 || h <= 0 || w <= 0 || k <= 0 || (k % 8)) return set_error(TOK_ERR_INVALID, "stem: bad shape");
  CUtensorMap tmX;
  PixelSrc src;
  int rc = stem_tmap(&tmX, xs2d, n, h, w, &src, 64);
  if (rc) return rc;
  return run_wgrad_tm(tmX, 64, src, dy, (long long)n * src.P * src.Q, k, dwp, static_cast<cudaStream_t>(stream));
}

int tok_linear_fwd(int m, int n, int k, const void* x, const void* w, const float* bias, void* y, void* stream) {
  if (m <= 0 || n <= 0 || k <= 0 || (n % 8) || (k % 8)) return set_error(TOK_ERR_INVALID, "linear: n and k must be positive multiples of 8");
  ConvFwdParams p;
  memset(&p, 0, sizeof(p));
  p.out = static_cast<__nv_bfloat16*>(y);
  p.ldo = n;
  p.bias = bias;
  return run_fwd(x, 1, 1, m, k, flat_src(), m, w, n, k, false, n, 0, p, static_cast<cudaStream_t>(stream));
}

int tok_linear_dgrad(int m, int n, int k, const void* dy, const void* w, void* dx, void* stream) {
  if (m <= 0 || n <= 0 || k <= 0 || (n % 8) || (k % 8)) return set_error(TOK_ERR_INVALID, "linear: n and k must be positive multiples of 8");
  ConvFwdParams p;
  memset(&p, 0, sizeof(p));
  p.out = static_cast<__nv_bfloat16*>(dx);
  p.ldo = k;
  return run_fwd(dy, 1, 1, m, n, flat_src(), m, w, n, k, true, k, 0, p, static_cast<cudaStream_t>(stream));
}

int tok_linear_dgrad_add(int m, int n, int k, const void* dy, const void* w, const void* addend, void* dx, void* stream) {
  if (m <= 0 || n <= 0 || k <= 0 || (n % 8) || (k % 8)) return set_error(TOK_ERR_INVALID, "linear: n and k must be positive multiples of 8");
  ConvFwdParams p;
  memset(&p, 0, sizeof(p));
  p.out = static_cast<__nv_bfloat16*>(dx);
  p.ldo = k;
  p.addend = static_cast<const __nv_bfloat16*>(addend);
  return run_fwd(dy, 1, 1, m, n, flat_src(), m, w, n, k, true, k, 0, p, static_cast<cudaStream_t>(stream));
}

int tok_linear_dgrad_gelu(int m, int n, int k, const void* dy, const void* w, const void* h, void* dx, float* colsum,
                          void* stream) {
  if (m <= 0 || n <= 0 || k <= 0 || (n % 8) || (k % 8)) return set_error(TOK_ERR_INVALID, "linear: n and k must be positive multiples of 8");
  if (h == nullptr) return set_error(TOK_ERR_INVALID, "linear_dgrad_gelu: the GELU input is required");
  static const bool v1 = getenv("TOK_CONV_V1") != nullptr;
  if (v1) return set_error(TOK_ERR_INVALID, "linear_dgrad_gelu: persistent kernel only (TOK_CONV_V1 is set)");
  ConvFwdParams p;
  memset(&p, 0, sizeof(p));
  p.out = static_cast<__nv_bfloat16*>(dx);
  p.ldo = k;
  p.addend = static_cast<const __nv_bfloat16*>(h);
  p.addend_mode = 1;
  if (colsum != nullptr) {
    // the kernel's column statistics come as (sum, sum of squares); the squares land in a scratch row nobody reads
    static float* sq_scratch = nullptr;
    static int sq_cap = 0;
    if (sq_cap < k) {
      if (sq_scratch) cudaFree(sq_scratch);
      sq_cap = k < 8192 ? 8192 : k;
      if (cudaMalloc(&sq_scratch, sq_cap * sizeof(float)) != cudaSuccess) {
        sq_scratch = nullptr;
        sq_cap = 0;
        return set_error(TOK_ERR_CUDA, "linear_dgrad_gelu: scratch allocation failed");
      }
      cudaMemsetAsync(sq_scratch, 0, sq_cap * sizeof(float), static_cast<cudaStream_t>(stream));
    }
    p.col_sum = colsum;
    p.col_sqsum = sq_scratch;
  }
  return run_fwd(dy, 1, 1, m, n, flat_src(), m, w, n, k, true, k, 0, p, static_cast<cudaStream_t>(stream));
}

int tok_linear_wgrad(int m, int n, int k, const void* x, const void* dy, float* dw, void* stream) {
  if (m <= 0 || n <= 0 || k <= 0 || (n % 8) || (k % 8)) return set_error(TOK_ERR_INVALID, "linear: n and k must be positive multiples of 8");
  return run_wgrad(x, 1, 1, m, k, flat_src(), dy, m, n, dw, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
