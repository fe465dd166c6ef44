// tok_conv.cuh — parameter blocks shared by the tcgen05 implicit-GEMM kernels and their host launchers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tok {

// How the pixel operand (activations / output gradients, NHWC bf16) is fetched by TMA.
struct PixelSrc {
  int im2col;   // 0: plain 2-D tile of a [rows][C] matrix, 1: TMA im2col walk over an NHWC tensor
  int P, Q;     // conv output spatial size (rows of the GEMM are (n, p, q) flattened)
  int stride;   // traversal stride
  int pad;      // lower padding (the lower corner is -pad)
  int dil;      // dilation
  int R, S;     // filter taps
};

// BatchNorm "finalize" fused into the conv launch: the last CTA of the grid to finish (ticket in *counter, a zeroed 32-bit
// word owned by the BatchNorm layer and handed back zeroed) turns the completed column sums into the per-channel affine
// and the running statistics, exactly as bn_finalize_train_kernel does, and zeroes the accumulators.
struct FwdFin {
  unsigned* counter;   // nullptr: no fused finalize
  float count, eps, momentum;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* scale;
  float* shift;
  float* save_mean;
  float* save_invstd;
};

// BatchNorm training-mode "finalize" fused into the APPLY pass (tok_bn_apply_train / tok_bn_apply_bits_train): every thread
// derives scale / shift of its 8 channels from the completed column sums, CTA 0 also writes them (with the saved mean /
// invstd for the backward) and updates the running statistics, and the last CTA to have read the sums (ticket) zeroes
// them for the next step.  Replaces the single-CTA bn_finalize_train launch between the conv and the apply pass.
struct ApplyFin {
  unsigned* counter;   // nullptr: scale / shift are read from memory (eval mode, or the separate finalize launch)
  float* sum;
  float* sqsum;
  float count, eps, momentum;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  float* scale;
  float* shift;
  float* save_mean;
  float* save_invstd;
  int C;
  // r2 "chain" mode (counter == nullptr, fused != 0): no ticket — the sums stay as they are and are zeroed by the NEXT
  // fused apply launch of the stream (zero_ptr / zero_n name the accumulators the previous launch left behind).
  int fused;       // != 0: scale / shift are derived from the sums inside the kernel
  int Cv;          // channels present in gamma / beta / running statistics (<= C; pad lanes get scale = shift = 0)
  float* zero_ptr;
  int zero_n;
};

// Forward / data-gradient implicit GEMM:  out[m, n] = sum_{tap, c} A_tap[m, c] * Wmat[n, wtap*Cin + c]
struct ConvFwdParams {
  int M;        // GEMM rows = N_img * P * Q
  int N;        // GEMM cols = output channels
  int Cin;      // channels per tap
  PixelSrc a;
  int flip_taps;  // dgrad: weight tap index = R*S-1-tap
  // epilogue
  __nv_bfloat16* out;
  long long ldo;              // elements between consecutive output rows
  const __nv_bfloat16* addend;  // optional, read at the same (row, col) as out before the store
  // optional ReLU bit mask of the addend (1 bit per element, 8 channels per byte in vector order, [rows][ldo / 8]): the
  // gradient that reaches a residual block's input through the identity shortcut is dout * [block output > 0]; reading
  // dout + the forward's bits here means the BatchNorm backward of the block tail need not materialise that product
  const uint8_t* addend_bits;
  float* col_sum;             // optional per-column sum of the *stored* (bf16-rounded) values
  float* col_sqsum;           // optional per-column sum of squares
  const float* bias;          // optional per-column bias added before rounding
  int relu;                   // optional ReLU before rounding
  // optional row scatter: GEMM row (n,p,q) is written to row (n*sc_H + p*sc_sh + sc_oh)*sc_W + q*sc_sw + sc_ow
  int scatter, sc_P, sc_Q, sc_H, sc_W, sc_sh, sc_sw, sc_oh, sc_ow;
  // optional explicit filter-tap table (strided dgrad by output parity class): weight tap of local tap i
  int use_tapmap;
  signed char tapmap[16];
  // MN-major operand descriptor geometry (bytes): chunk distance, 8-row group distance, advance per 16-row MMA step.
  int mn_lbo, mn_sbo, mn_kadv;
  // bring-up aid (TOK_CONV_PROFILE=1): per-CTA cycle counts of the epilogue phases, 8 slots per CTA
  long long* prof;
  FwdFin fin;   // persistent kernel only
  // 0: out += addend.  1: out = bf16(acc) * gelu'(addend) — the GELU backward folded into the dgrad of the layer behind it
  // (timm Mlp: fc1 -> GELU -> fc2; persistent kernel, 128-wide tile with addend buffers only)
  int addend_mode;
  int m_group;       // persistent kernel tile order: m-tiles per group (0 = whole strips; tok_conv.cu: decode_tile)
  int defer_stats;   // persistent kernel, two staging buffers: statistics pass of tile t inside iteration t + 1
};

// Weight-gradient implicit GEMM:  dW[co, tap*Cin + ci] += sum_{pix in split} dy[pix, co] * x_tap[pix, ci]
struct ConvWgradParams {
  int Mpix;     // number of output pixels (rows of dy)
  int Cout;
  int Cin;
  PixelSrc x;   // how x_tap rows are fetched (pixelsPerColumn = 64)
  float* dw;    // [Cout][R*S*Cin] fp32, accumulated with atomics (caller zero-fills)
  long long ldw;
  int chunks_per_split;  // number of 64-pixel K blocks handled by one CTA
  int mn_lbo, mn_sbo, mn_kadv;
};

// Defaults 8192 / 1024 / 2048; TOK_MN_LBO / TOK_MN_SBO / TOK_MN_KADV override them (bring-up aid).
void mn_desc_geometry(int* lbo, int* sbo, int* kadv);

}  // namespace tok
