/* tokb200.h — C ABI of libtokb200.so: the sm_100a kernels behind torchok_b200.
 *
 * The reference (eora-ai/torchok) has no FFI of its own: its hot path is torch.nn modules assembled by the
 * registry/constructor (torchok/constructor/registry.py:45-99, torchok/tasks/classification.py:47-73) and every
 * FLOP is dispatched by torch to cuDNN/cuBLAS/ATen.  Each entry point below names the reference call site whose
 * torch dispatch it replaces.  The Python host (torchok_b200/_lib.py) binds these with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; nothing here allocates, frees or synchronises (the one
 *     exception: tok_ipc_alloc / tok_ipc_open own the peer-mapped arenas of the multi-GPU gradient exchange);
 *   - activations are NHWC bf16 ("pixel-major"), conv weights are [Cout][R][S][Cin] bf16 (= a torch OIHW tensor in
 *     channels_last memory format), weight gradients are fp32 in the same layout;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value 0 = success, negative = tokStatus; tok_last_error() returns a thread-local message;
 *   - channel counts must be multiples of 8 (16-byte TMA rows).
 */
#ifndef TOKB200_H_
#define TOKB200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  TOK_OK = 0,
  TOK_ERR_INVALID = -1,   /* bad argument / unsupported shape */
  TOK_ERR_CUDA = -2,      /* CUDA runtime / driver error */
  TOK_ERR_NODRIVER = -3   /* libcuda entry points unavailable (no GPU driver on this host) */
} tokStatus;

int tok_version(void);
const char* tok_last_error(void);
/* 0 if a CUDA device with compute capability 10.x is usable, else negative. */
int tok_device_ok(void);

/* Bring-up aid: with TOK_CONV_PROFILE=1 in the environment the persistent conv kernel records per-CTA cycle counts
 * of its epilogue phases (16 int64 per CTA: two observer threads x 8 slots); this copies the last launch's counters
 * to the host and returns the number of entries written (0 when profiling is off). */
int tok_debug_conv_profile(long long* host_out, int max_entries);
int tok_debug_attn_profile(long long* host_out, int max_entries);   /* TOK_ATTN_PROFILE=1: window-attention forward */

/* ---- convolution (torch.nn.Conv2d inside ConvBnAct, torchok/models/modules/bricks/convbnact.py:38-53; timm
 *      BasicBlock/Bottleneck convs built by torchok/models/backbones/resnet.py:363-405) ------------------------- */
typedef struct {
  int n, h, w, c;      /* input NHWC */
  int k;               /* output channels */
  int r, s;            /* filter */
  int stride, pad, dil;
  /* Optional (0 = k / c): dimensions of the WEIGHT tensor [wk][r][s][wc] when they differ from the activation pitches
   * k / c — a channel count that is not a multiple of 8 (HRNet's 18 / 36-channel branches,
   * torchok/models/backbones/hrnet.py:140-192) keeps its weights and weight gradient unpadded; activations still carry
   * the pitch rounded up to 8 with zero pad lanes.  Supported where tok_conv_halo_caps() says so. */
  int wk, wc;
} tokConvDesc;

/* Bit mask of the operations of this convolution that run on the halo 3x3 kernels (tok_conv3.cu) and therefore accept
 * unpadded weights (wk / wc): 1 = fprop, 2 = dgrad, 4 = wgrad. */
int tok_conv_halo_caps(const tokConvDesc* d);

/* Output spatial size of a convolution (same arithmetic as torch.nn.Conv2d). */
void tok_conv_out_hw(const tokConvDesc* d, int* p, int* q);

/* y[n,p,q,k] = sum x[n, p*stride-pad+r*dil, q*stride-pad+s*dil, c] * w[k,r,s,c]  (+bias[k]) (+addend) (relu)
 * If sum/sqsum are non-NULL, per-channel sums of the stored bf16 values and of their squares are ATOMICALLY ADDED
 * to them (caller zero-fills): these are the BatchNorm batch statistics (torch.nn.BatchNorm2d training mode). */
int tok_conv_fprop(const tokConvDesc* d, const void* x, const void* w, void* y, float* sum, float* sqsum,
                   const void* addend, const float* bias, int relu, void* stream);
/* tok_conv_fprop (with statistics) + tok_bn_finalize_train in ONE launch — the Conv2d -> BatchNorm2d(training) pair of
 * torchok/models/modules/bricks/convbnact.py:48-53: the last CTA of the persistent grid to finish (ticket in *counter, a
 * zero-initialised 32-bit word owned by the BatchNorm layer, handed back zeroed) computes scale / shift / save_mean /
 * save_invstd, updates the running statistics (torch.nn.BatchNorm2d semantics) and zeroes sum / sqsum. */
int tok_conv_fprop_bn(const tokConvDesc* d, const void* x, const void* w, void* y, float* sum, float* sqsum,
                      const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                      float* running_var, float* scale, float* shift, float* save_mean, float* save_invstd,
                      unsigned* counter, void* stream);

/* dx = conv_transpose(dy, w).  `addend` (nullable, NHWC like dx, may alias dx) is added before the store.
 * `ws` is scratch of tok_conv_dgrad_workspace_bytes(d) bytes (needed for strided RxS>1 filters). */
size_t tok_conv_dgrad_workspace_bytes(const tokConvDesc* d);
int tok_conv_dgrad(const tokConvDesc* d, const void* dy, const void* w, void* dx, const void* addend, void* ws,
                   void* stream);

/* tok_conv_dgrad whose addend is masked on the fly: dx = dgrad(dy) + addend * [bit set], addend_bits = the 1-bit-per-
 * element ReLU mask tok_bn_apply_bits wrote for the block output ([pixels][c / 8], 8 channels per byte).  The gradient that
 * enters a residual block's input through the identity shortcut is dout * [block output > 0] (timm BasicBlock /
 * Bottleneck `x += shortcut; x = act(x)`, built by torchok/models/backbones/resnet.py:363-405): with this form the
 * BatchNorm backward of the block tail does not write that product (8 -> 6 bytes per element on 75 % of a ResNet's
 * BatchNorm-backward traffic).  Stride 1 only; see tok_conv_dgrad_masked_supported. */
int tok_conv_dgrad_masked_supported(const tokConvDesc* d);
int tok_conv_dgrad_masked(const tokConvDesc* d, const void* dy, const void* w, void* dx, const void* addend,
                          const void* addend_bits, void* ws, void* stream);
/* dw[k,r,s,c] += sum_pixels dy[n,p,q,k] * x[n, ..., c]   (fp32, atomically accumulated; caller zero-fills). */
int tok_conv_wgrad(const tokConvDesc* d, const void* x, const void* dy, float* dw, void* stream);

/* ---- dense layers (torch.nn.Linear in LinearHead, torchok/models/heads/representation/linear_head.py:25-36;
 *      PoolingLinear, torchok/models/poolings/classification/linear.py:12-19) -------------------------------- */
/* y[m,n] = sum_k x[m,k] * w[n,k] + bias[n] ; x,w,y bf16, bias fp32 (nullable). */
int tok_linear_fwd(int m, int n, int k, const void* x, const void* w, const float* bias, void* y, void* stream);
/* dx[m,k] = sum_n dy[m,n] * w[n,k] */
int tok_linear_dgrad(int m, int n, int k, const void* dy, const void* w, void* dx, void* stream);
/* dx[m,k] = sum_n dy[m,n] * w[n,k] + addend[m,k]: the residual-branch gradient of a transformer block
 * (x + f(x), timm SwinTransformerBlock) joins the branch gradient in the GEMM epilogue instead of a separate add pass. */
int tok_linear_dgrad_add(int m, int n, int k, const void* dy, const void* w, const void* addend, void* dx, void* stream);
/* dx[m,k] = gelu'(h[m,k]) * bf16(sum_n dy[m,n] * w[n,k]); colsum[k] += sum_m dx[m,k] (optional): the backward of
 * timm Mlp's `fc2(act(h))` with respect to h in one launch — the exact-erf GELU derivative is applied in the dgrad epilogue,
 * and the column sums are the bias gradient of the layer that produced h (fc1).  Same rounding points as
 * tok_linear_dgrad followed by tok_gelu_bwd. */
int tok_linear_dgrad_gelu(int m, int n, int k, const void* dy, const void* w, const void* h, void* dx, float* colsum,
                          void* stream);
/* dw[n,k] += sum_m dy[m,n] * x[m,k]  (fp32, atomically accumulated) */
int tok_linear_wgrad(int m, int n, int k, const void* x, const void* dy, float* dw, void* stream);


/* ---- ResNet stem: 7x7 stride-2 pad-3 conv on <=4 input channels (torchok/models/backbones/resnet.py:488-490,
 *      542-545).  The NCHW image is packed once into a zero-padded 2x2 space-to-depth NHWC16 tensor
 *      [N][H2][W2][16]; the conv then runs as a 4-tap im2col GEMM with K = 4*64 through the common kernel. -------- */
void tok_stem_geometry(int h, int w, int* p, int* q, int* h2, int* w2);
int tok_stem_pack_input(int n, int c, int h, int w, int src_is_bf16, const void* src_nchw, void* xs2d, void* stream);
/* w: fp32 [K][7][7][C] (OIHW in channels_last memory format) -> wp: bf16 [K][256] */
int tok_stem_pack_weight(int k, int c, const float* w, void* wp, void* stream);
int tok_stem_unpack_wgrad(int k, int c, const float* dwp, float* dw, int accumulate, void* stream);
int tok_stem_conv_fprop(int n, int h, int w, int k, const void* xs2d, const void* wp, void* y, float* sum,
                        float* sqsum, void* stream);
/* dwp: fp32 [K][256], atomically accumulated (caller zero-fills) */
int tok_stem_conv_wgrad(int n, int h, int w, int k, const void* xs2d, const void* dy, float* dwp, void* stream);

/* ---- BatchNorm2d (+ReLU, +residual add) around the convs (torch.nn.BatchNorm2d in ConvBnAct,
 *      torchok/models/modules/bricks/convbnact.py:44-53; timm block tails `x += shortcut; act(x)`) --------------- */
/* batch statistics (sums from tok_conv_fprop) -> scale/shift, saved mean/invstd, running-stat update.
 * sum/sqsum are CONSUMED: they are reset to zero so the same accumulators can be reused by the next step. */
int tok_bn_finalize_train(int C, double count, float* sum, float* sqsum, const float* gamma,
                          const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                          float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);
int tok_bn_finalize_eval(int C, const float* running_mean, const float* running_var, const float* gamma,
                         const float* beta, float eps, float* scale, float* shift, void* stream);
/* The same with gamma / beta / running statistics that hold only c_valid <= C entries: channels c_valid .. C-1 are the
 * pad lanes of a channel count that is not a multiple of 8 (nn.BatchNorm2d(18) of HRNet); they get scale = shift = 0
 * and the parameter buffers are never touched past c_valid. */
int tok_bn_finalize_train_cv(int C, int c_valid, double count, float* sum, float* sqsum, const float* gamma,
                             const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                             float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);
int tok_bn_finalize_eval_cv(int C, int c_valid, const float* running_mean, const float* running_var,
                            const float* gamma, const float* beta, float eps, float* scale, float* shift,
                            void* stream);
/* out = act(y*scale + shift (+residual)) over a [rows][C] bf16 matrix */
int tok_bn_apply(long long rows, int C, const void* y, const float* scale, const float* shift, const void* residual,
                 int relu, void* out, void* stream);
/* g = (dout (+dout2)) * [out > 0] (mask skipped when out == NULL); sum_g += sum g, sum_gy += sum g*y (zero-filled
 * by the caller) */
int tok_bn_bwd_reduce(long long rows, int C, const void* dout, const void* dout2, const void* out, const void* y,
                      float* sum_g, float* sum_gy, void* stream);
/* sum_g/sum_gy are CONSUMED (reset to zero) like the forward accumulators. */
int tok_bn_bwd_finalize(int C, double count, float* sum_g, float* sum_gy, const float* save_mean,
                        const float* save_invstd, const float* gamma, float* coef_a, float* coef_c1, float* coef_c0,
                        float* dgamma, float* dbeta, int accumulate, void* stream);
int tok_bn_bwd_finalize_cv(int C, int c_valid, double count, float* sum_g, float* sum_gy, const float* save_mean,
                           const float* save_invstd, const float* gamma, float* coef_a, float* coef_c1,
                           float* coef_c0, float* dgamma, float* dbeta, int accumulate, void* stream);
/* dy = coef_a*g + coef_c1*y + coef_c0 ; dres (nullable) receives g */
int tok_bn_bwd_apply(long long rows, int C, const void* dout, const void* dout2, const void* out, const void* y,
                     const float* coef_a, const float* coef_c1, const float* coef_c0, void* dy, void* dres,
                     void* stream);

/* tok_bn_finalize_train fused INTO the apply pass (the BatchNorm2d(training) + ReLU (+ residual) tail of ConvBnAct,
 * torchok/models/modules/bricks/convbnact.py:48-53): every CTA derives scale / shift from the completed column sums, CTA 0
 * also stores them (+ saved mean / invstd) and updates the running statistics, the last CTA through the ticket (*counter,
 * a zeroed 32-bit word owned by the layer, handed back zeroed) clears the sums.  One launch less per BatchNorm layer.
 * tok_bn_apply_train needs C / 8 to divide its grid stride: ask tok_bn_apply_train_supported first. */
int tok_bn_apply_train_supported(long long rows, int C);
int tok_bn_apply_train(long long rows, int C, const void* y, float* sum, float* sqsum, const float* gamma,
                       const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                       float* scale, float* shift, float* save_mean, float* save_invstd, unsigned* counter,
                       const void* residual, int relu, void* out, void* stream);
int tok_bn_apply_bits_train(long long rows, int C, const void* y, float* sum, float* sqsum, const float* gamma,
                            const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                            float* scale, float* shift, float* save_mean, float* save_invstd, unsigned* counter,
                            const void* residual, void* out, void* bits, void* stream);
/* Second-generation passes (tok_bn2.cu): the backward no longer re-reads the activation to rebuild the ReLU mask.
 * mask_mode 0: no activation; 1: plain conv->BN->ReLU unit, mask = (y*scale + shift > 0) recomputed from the forward's
 * scale/shift; 2: residual tail, mask = bits (1 bit per element, one byte per 8-channel vector, written by
 * tok_bn_apply_bits which computes out = relu(y*scale + shift + residual)). */
int tok_bn_apply_bits(long long rows, int C, const void* y, const float* scale, const float* shift,
                      const void* residual, void* out, void* bits, void* stream);
int tok_bn_bwd_reduce2(long long rows, int C, const void* dout, const void* dout2, const void* y, int mask_mode,
                       const void* bits, const float* scale, const float* shift, float* sum_g, float* sum_gy,
                       void* stream);
/* tok_bn_bwd_reduce2 + tok_bn_bwd_finalize in one launch: the last CTA to finish (ticket in *counter, a zero-initialised
 * 32-bit word owned by the BatchNorm layer, handed back zeroed) computes coef_a / coef_c1 / coef_c0 and the gamma / beta
 * gradients from the completed sums and zeroes the accumulators. */
int tok_bn_bwd_reduce2_finalize(long long rows, int C, const void* dout, const void* dout2, const void* y, int mask_mode,
                                const void* bits, const float* scale, const float* shift, float* sum_g, float* sum_gy,
                                const float* save_mean, const float* save_invstd, const float* gamma, float* coef_a,
                                float* coef_c1, float* coef_c0, float* dgamma, float* dbeta, int accumulate,
                                unsigned* counter, void* stream);
/* "Chain" form of the fused finalize + apply (r2): every CTA derives scale / shift of its channels from the batch sums,
 * CTA 0 publishes scale / shift / saved mean / invstd and updates the running statistics — and, instead of a ticket to
 * zero the sums, zeroes the zero_n floats at zero_ptr: the accumulators the PREVIOUS chain launch of the same stream
 * left behind (NULL / 0 for the first).  The caller keeps that pointer; sum / sqsum of this launch stay non-zero until
 * the next chain launch (or an explicit memset).  Replaces the single-CTA tok_bn_finalize_train launch per unit. */
int tok_bn_apply_chain(long long rows, int C, int c_valid, const void* y, float* sum, float* sqsum, const float* gamma,
                       const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                       float* scale, float* shift, float* save_mean, float* save_invstd, float* zero_ptr, int zero_n,
                       const void* residual, int relu, void* out, void* stream);
int tok_bn_apply_bits_chain(long long rows, int C, int c_valid, const void* y, float* sum, float* sqsum,
                            const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                            float* running_var, float* scale, float* shift, float* save_mean, float* save_invstd,
                            float* zero_ptr, int zero_n, const void* residual, void* out, void* bits, void* stream);
int tok_bn_bwd_reduce2_finalize_cv(long long rows, int C, int c_valid, const void* dout, const void* dout2, const void* y,
                                   int mask_mode, const void* bits, const float* scale, const float* shift, float* sum_g,
                                   float* sum_gy, const float* save_mean, const float* save_invstd, const float* gamma,
                                   float* coef_a, float* coef_c1, float* coef_c0, float* dgamma, float* dbeta,
                                   int accumulate, unsigned* counter, void* stream);
/* tok_bn_bwd_reduce2_finalize_cv + tok_bn_bwd_apply2 in ONE launch, for tensors small enough to stay in L2 between the
 * two passes: every CTA reduces its rows, the last through the ticket finalizes and flips *release, the others wait for
 * the flip and apply dy = coef_a g + coef_c1 y + coef_c0 to the rows they already read.  `counter` and `release` are two
 * zero-initialised 32-bit words owned by the layer (release keeps toggling; no reset needed).  The grid is one resident
 * wave; the wait is bounded (~4 s, then the context traps). */
int tok_bn_bwd_fused_cv(long long rows, int C, int c_valid, const void* dout, const void* dout2, const void* y,
                        int mask_mode, const void* bits, const float* scale, const float* shift, float* sum_g,
                        float* sum_gy, const float* save_mean, const float* save_invstd, const float* gamma,
                        float* coef_a, float* coef_c1, float* coef_c0, float* dgamma, float* dbeta, int accumulate,
                        unsigned* counter, unsigned* release, void* dy, void* dres, void* stream);
int tok_bn_bwd_apply2(long long rows, int C, const void* dout, const void* dout2, const void* y, int mask_mode,
                      const void* bits, const float* scale, const float* shift, const float* coef_a,
                      const float* coef_c1, const float* coef_c0, void* dy, void* dres, void* stream);

/* Fused ResNet stem tail (resnet.py:488-490,510): BN -> ReLU -> maxpool 3x3 stride 2 pad 1 in one pass over the conv
 * output y (act is written only when non-NULL: the act1 feature of forward_features), and its backward with the pooled
 * gradient gathered on the fly (dact nullable = gradient of the act1 feature): pass 1 reduces sum_g / sum_gy, pass 2
 * writes dy = a*g + c1*y + c0.  argmax: one byte per pooled element (window slot of the first maximum). */
int tok_stem_bn_relu_pool_fwd(int n, int h, int w, int c, const void* y, const float* scale, const float* shift,
                              void* act, void* pooled, void* argmax, void* stream);
int tok_stem_bwd_reduce(int n, int h, int w, int c, const void* dpooled, const void* argmax, const void* dact,
                        const void* y, const float* scale, const float* shift, float* sum_g, float* sum_gy,
                        void* stream);
int tok_stem_bwd_apply(int n, int h, int w, int c, const void* dpooled, const void* argmax, const void* dact,
                       const void* y, const float* scale, const float* shift, const float* coef_a,
                       const float* coef_c1, const float* coef_c0, void* dy, void* stream);

/* dst[n, p*stride, q*stride, :] += src_compact[n, p, q, :] (NHWC bf16, P = (h-1)/stride+1): merges the compact data
 * gradient of a strided 1x1 downsample conv (timm downsample_conv, resnet.py:14) into the block-input gradient. */
int tok_strided_add(int n, int h, int w, int c, int stride, const void* src_compact, void* dst, void* stream);

/* ---- pooling (torch.nn.MaxPool2d, resnet.py:510; timm SelectAdaptivePool2d, poolings/classification/pooling.py:8-12)
 * argmax: one byte per output element (window slot of the first maximum). */
int tok_maxpool_fwd(int n, int h, int w, int c, int k, int s, int pad, const void* x, void* out, void* argmax,
                    void* stream);
int tok_maxpool_bwd(int n, int h, int w, int c, int k, int s, int pad, const void* dout, const void* argmax, void* dx,
                    void* stream);
/* mode 0 avg, 1 max, 2 avgmax */
int tok_gap_fwd(int n, int hw, int c, int mode, const void* x, void* out, void* stream);
int tok_gap_bwd(int n, int hw, int c, const void* dout, void* dx, void* stream);
/* backward of mode 1 / 2: the first maximal position (F.adaptive_max_pool2d's choice) takes the max-path gradient;
 * x is the pooled activation [n][hw][c] (timm SelectAdaptivePool2d, pooling.py:8-12) */
int tok_gap_bwd_max(int n, int hw, int c, int mode, const void* dout, const void* x, void* dx, void* stream);

/* ---- loss (torch.nn.CrossEntropyLoss, torchok/losses/__init__.py:26): *loss_sum += inv_norm * sum NLL (nullable);
 * dlogits (nullable) = (softmax - onehot) * gscale * (*gscale_dev if non-NULL: the upstream scalar gradient, read on
 * the device so no host sync is needed); `correct` (nullable) counts rows whose argmax equals the target. */
int tok_softmax_xent(int rows, int C, long long ld, const void* logits, const long long* target, float* loss_sum,
                     void* dlogits, float inv_norm, float gscale, const float* gscale_dev, long long ignore_index,
                     int* correct, void* stream);

/* ---- Swin-V2 passes (tok_swin.cu; timm swin_transformer_v2 semantics used by torchok/models/backbones/swin.py:71-81) ------
 * LayerNorm over the last dimension of a (rows, C) bf16 matrix, C <= 1024; with `residual` the output is
 * residual + rowscale[row / rows_per_sample] * LN(x) (res-post-norm block tail with stochastic depth; rowscale nullable).
 * Backward: dx for the LN input (the residual gradient is dout itself); dgamma / dbeta are atomically ACCUMULATED.
 * `dxsum` (nullable, C floats, ACCUMULATED): column sums of dx = the bias gradient of the torch.nn.Linear whose output
 * this LayerNorm consumed (timm Mlp.fc2 / WindowAttention.proj), saving that layer's own pass over dx; only for widths
 * with tok_layernorm_has_dxsum(C) == 1 (C = 8 * {1,2,3,4} * {4,8,16,32}, i.e. every Swin width). */
int tok_layernorm_fwd(long long rows, int C, const void* x, const float* gamma, const float* beta, float eps,
                      const void* residual, const float* rowscale, int rows_per_sample, void* out, float* mean,
                      float* rstd, void* stream);
int tok_layernorm_bwd(long long rows, int C, const void* x, const float* gamma, const float* mean, const float* rstd,
                      const void* dout, const float* rowscale, int rows_per_sample, void* dx, float* dgamma,
                      float* dbeta, float* dxsum, void* stream);
int tok_layernorm_has_dxsum(int C);
/* timm PatchEmbed.proj = Conv2d(3, E, kernel 4, stride 4) followed by flatten(2).transpose(1, 2)
 * (torchok/models/backbones/swin.py:156-171 constructs it): `image` is the (B, 3, H, W) fp32 NCHW batch exactly as the
 * task hands it over, `tokens` the (B * H/4 * W/4, E) bf16 matrix.  `weight` / `dweight` are fp32 (E, 3, 4, 4) tensors
 * addressed through `wstride` = their four element strides (the masters live in channels_last memory).  Backward
 * ACCUMULATES dweight / dbias (no image gradient: the image is the network input).  E % 32 == 0, E <= 256. */
int tok_patch_embed_supported(int Cin, int patch, int H, int W, int E);
/* im2col of the non-overlapping 4x4 patches of an NCHW image (fp32 or bf16) into a bf16 [B*H/4*W/4][48] matrix with the
 * columns in (dy, dx, c) order = the memory order of the [E][4][4][3] conv weight, so PatchEmbed.proj is tok_linear_fwd /
 * tok_linear_wgrad on it (tensor cores) instead of the CUDA-core tok_patch_embed_* kernels. */
int tok_patchify(int B, int C, int H, int W, int patch, int src_is_bf16, const void* image, void* dst, void* stream);
int tok_patch_embed_fwd(int B, int H, int W, int E, const float* image, const float* weight, const float* bias,
                        const int* wstride, void* tokens, void* stream);
int tok_patch_embed_bwd(int B, int H, int W, int E, const float* image, const void* dtokens, const int* wstride,
                        float* dweight, float* dbias, void* stream);
/* timm PatchMerging gather on a (B, H, W, C) bf16 tensor: dst[b,i,j,q*C+c] = src[b, 2i+(q&1), 2j+(q>>1), c] (q = 0..3, the
 * order of torch.cat([x[:,0::2,0::2], x[:,1::2,0::2], x[:,0::2,1::2], x[:,1::2,1::2]], -1)); inverse != 0 scatters back
 * (the backward).  H, W even, C % 8 == 0. */
int tok_patch_merge(int B, int H, int W, int C, const void* src, void* dst, int inverse, void* stream);
/* exact (erf) GELU of timm's Mlp (Mlp.act between fc1 and fc2).  Backward over an (n / C, C) matrix; `dbias` (nullable,
 * C floats, ACCUMULATED, needs C % 128 == 0) receives the column sums of dx = the bias gradient of Mlp.fc1. */
int tok_gelu_fwd(long long n, const void* x, void* y, void* stream);
int tok_gelu_bwd(long long n, int C, const void* x, const void* dy, void* dx, float* dbias, void* stream);
/* WindowAttention.forward on the (B, H, W, 3C) qkv tensor ([3][heads][32] per token): cosine attention with
 * exp(min(logit_scale, ln 100)), additive bias[heads][N][N], shifted-window mask (-100) computed from `shift`; windows are
 * gathered from / scattered to their home positions (no roll / partition copies).  window <= 8, head_dim == 32.
 * Backward (tcgen05: S, dP, dV, dK, dQ as UMMA chains, P / dS kept on chip) recomputes the probabilities; dbias /
 * dlogit_scale are atomically ACCUMULATED.  `dqkv_colsum` (nullable, 3C floats, ACCUMULATED; the k third is left
 * untouched): column sums of dq and dv = the q_bias / v_bias gradients of timm's WindowAttention.qkv. */
int tok_window_attn_fwd(int B, int H, int W, int C, int heads, int ws, int shift, const void* qkv,
                        const float* logit_scale, const float* bias, void* out, void* stream);
int tok_window_attn_bwd(int B, int H, int W, int C, int heads, int ws, int shift, const void* qkv,
                        const float* logit_scale, const float* bias, const void* dout, void* dqkv, float* dbias,
                        float* dlogit_scale, float* dqkv_colsum, void* stream);

/* ---- HRNet / segmentation passes (tok_seg.cu) ---------------------------------------------------------------------------
 * timm HighResolutionModule fuse (torchok/models/backbones/hrnet.py:167-192): out = relu(sum_t nearest_up(term_t)), term t
 * at resolution (h >> shifts[t], w >> shifts[t]); bits (nullable) = ReLU mask, 1 bit/element.  Backward per term. */
int tok_fuse_sum_fwd(int n, int h, int w, int c, int nterms, const void* const* terms, const int* shifts, int relu,
                     void* out, void* bits, void* stream);
int tok_fuse_sum_bwd(int n, int h, int w, int c, int shift, const void* dout, const void* bits, void* dterm,
                     void* stream);
/* F.interpolate(mode='bilinear', align_corners=False) (necks/segmentation/hrnet.py:35-38, heads/segmentation/base.py:37)
 * writing into channels [dst_c_offset, dst_c_offset + c) of a dst_c-channel NHWC tensor (= the torch.cat of the neck). */
int tok_bilinear_fwd(int n, int hi, int wi, int c, int ho, int wo, const void* src, void* dst, int dst_c,
                     int dst_c_offset, void* stream);
int tok_bilinear_bwd(int n, int hi, int wi, int c, int ho, int wo, const void* dout, int dout_c, int dout_c_offset,
                     void* dsrc, void* stream);
/* CrossEntropyLoss over many short rows (segmentation logits, C <= 64, row pitch ld): forward (dlogits == NULL) adds the
 * NLL sum and the number of non-ignored rows to loss_sum / count; backward writes (softmax - onehot) * gscale *
 * (*gscale_dev) * (*inv_count_dev). */
int tok_softmax_xent_small(long long rows, int C, int ld, const void* logits, const long long* target,
                           float* loss_sum, float* count, void* dlogits, const float* inv_count_dev, float gscale,
                           const float* gscale_dev, long long ignore_index, void* stream);
/* DiceLoss(mode='multiclass', from_logits=True) of torchok/losses/segmentation/dice.py:85-188 on (rows, C <= 64) bf16
 * logits (NHWC rows of pitch ld) with int64 targets: tok_dice_stats ACCUMULATES per class [3][64] = (sum p_c [t = c],
 * sum p_c, count_c) with p = softmax; tok_dice_finalize turns them into the scalar loss (mean over classes, classes
 * without a true pixel masked, optional -log) and the per-class coefficients coef[2][64] = (a_c, b_c) of
 * d loss / d p_c(pixel) = a_c [t = c] + b_c; tok_dice_bwd writes dlogits = gscale * p o (dp - sum_k p_k dp_k). */
int tok_dice_stats(long long rows, int C, int ld, const void* logits, const long long* target, float* stats,
                   void* stream);
int tok_dice_finalize(int C, const float* stats, float smooth, float eps, int log_loss, float* loss, float* coef,
                      void* stream);
int tok_dice_bwd(long long rows, int C, int ld, const void* logits, const long long* target, const float* coef,
                 const float* gscale_dev, void* dlogits, void* stream);

/* ---- embedding heads and the pairwise loss (tok_heads.cu) ------------------------------------------------------------
 * F.normalize (LinearHead(normalize=True), linear_head.py:33-35; ArcFaceHead, arcface_head.py:125-126):
 *   xhat = scale * x / max(|x|, 1e-12) per row, bf16 with pitch ld_out (pad columns zeroed), inv_norm = 1/max(|x|,eps);
 *   backward dx = scale*inv_norm*(g - u(u.g)), u = x/|x|; dx is bf16 (stored) or fp32 (accumulated if `accumulate`). */
int tok_rownorm_fwd(int rows, int d, const void* x, int x_is_bf16, float scale, void* xhat_bf16, int ld_out,
                    float* inv_norm, void* stream);
int tok_rownorm_bwd(int rows, int d, const void* x, int x_is_bf16, const float* inv_norm, float scale, const void* g,
                    int g_is_bf16, int ld_g, void* dx, int dx_is_bf16, int accumulate, void* stream);
/* ArcFaceHead.__add_margin (arcface_head.py:95-108).  `logits` = scale*cosine comes from tok_linear_fwd on
 * xs = scale*x_hat and wh = w_hat; the forward replaces the target column of each row by scale*phi(cos_t) (cos_t
 * recomputed in fp32 and saved), the backward multiplies dlogits[r, target] by dphi/dcos. */
int tok_arcface_margin_fwd(int rows, int d, int ld_x, const void* xs_bf16, const void* wh_bf16,
                           const long long* target, int num_classes, void* logits_bf16, long long ld_logits,
                           float scale, float margin, int easy_margin, float* cos_t, void* stream);
int tok_arcface_margin_bwd(int rows, const long long* target, int num_classes, const float* cos_t,
                           void* dlogits_bf16, long long ld_logits, float scale, float margin, int easy_margin,
                           void* stream);
/* ContrastiveLoss.calc_loss (losses/representation/pairwise.py:126-136): S = cdist(emb1, emb2) (B x M, fp32),
 * loss_rows[i] = sum_j (1-R)relu(margin-S)^2 + R S^2; backward given d(loss_rows). */
int tok_contrastive_fwd(int B, int M, int d, const float* emb1, const float* emb2, const float* R, float margin,
                        float* S, float* loss_rows, void* stream);
int tok_contrastive_bwd(int B, int M, int d, const float* emb1, const float* emb2, const float* R, const float* S,
                        const float* grad_rows, float margin, float* d_emb1, float* d_emb2, void* stream);

/* ---- retrieval metric (IndexBasedMeter.compute, torchok/metrics/index_base_metric.py:170-270: normalise ->
 *      faiss IndexFlatIP/L2 .search(q, k+1) in batches, :444-545).  Three steps, all on the device:
 *      1. tok_l2_normalize_rows: xn = x / |x| per row (normalize != 0; SURVEY S6: the reference's golden answers encode
 *         row-wise cosine) or a copy; also writes a zero-padded bf16 copy with pitch ld_bf16 and the squared norms.
 *      2. tok_topk_candidates: S = Q * G^T on tcgen05 with a fused running top-kp per query row (kp in {8,16,32}); S is
 *         never materialised.  g_sqnorm == NULL ranks by inner product, otherwise by -(|g|^2 - 2 q.g) (= L2 order).
 *      3. tok_topk_rerank: exact fp32 re-scoring of the kp candidates and emission of the best k in faiss order
 *         (IP descending / squared L2 ascending, ties -> lower index; missing entries -1 / -+inf), int64 indices. */
int tok_l2_normalize_rows(int n, int d, int normalize, const float* x, float* xn, void* xn_bf16, int ld_bf16,
                          float* sqnorm, void* stream);
int tok_topk_candidates(int nq, int ng, int d, int kp, const void* q_bf16, const void* g_bf16, const float* g_sqnorm,
                        float* cand_score, int* cand_idx, void* stream);
int tok_topk_rerank(int nq, int d, int kp, int k, int metric, const float* q_f32, const float* g_f32,
                    const int* cand_idx, float* out_score, long long* out_idx, void* stream);

/* ---- layout (task boundary is NCHW, torchok/tasks/classification.py:108-109) ------------------------------------ */
int tok_nchw_to_nhwc(int n, int c, int hw, int cp, int src_is_bf16, const void* src, void* dst, void* stream);
int tok_nhwc_to_nchw(int n, int c, int hw, int cp, int dst_is_bf16, const void* src, void* dst, void* stream);

/* ---- optimizer steps on flat fp32 parameter arenas, fused with the bf16 shadow-weight cast
 *      (torch.optim.SGD / Adam / AdamW registered in torchok/optim/optimizers/__init__.py:9-19) ------------------- */
int tok_sgd_step(long long n, float* param, const float* grad, float* momentum_buf, void* shadow_bf16, float lr,
                 float momentum, float weight_decay, float dampening, int nesterov, float grad_scale, int first_step,
                 void* stream);
int tok_adam_step(long long n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled, int step,
                  float grad_scale, void* stream);
/* Graph-replayable forms used by the CUDA-stream step loop (replaces Lightning's optimizer_step around
 * torchok/tasks/base.py:125-133): the learning rate and the step counter live in device memory (lr_dev[0];
 * step_dev[0] is advanced by one on the device before it is used, so bias correction / first-step momentum follow
 * the replay count), and with zero_grad != 0 the consumed gradient arena is cleared for the next step's atomic
 * accumulation. */
int tok_sgd_step_dev(long long n, float* param, float* grad, float* momentum_buf, void* shadow_bf16,
                     const float* lr_dev, int* step_dev, float momentum, float weight_decay, float dampening,
                     int nesterov, float grad_scale, int zero_grad, void* stream);
int tok_adam_step_dev(long long n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                      const float* lr_dev, int* step_dev, float beta1, float beta2, float eps, float weight_decay,
                      int decoupled, float grad_scale, int zero_grad, void* stream);
/* The same steps with `paramwise_cfg` (torchok/constructor/constructor.py:162-251: bias_lr_mult, bias_decay_mult,
 * norm_decay_mult, dwconv_decay_mult, custom_keys): the reference builds one torch.optim param group per parameter; here
 * the flat kernel looks up lr / weight-decay MULTIPLIERS in a sorted per-parameter segment table (seg_begin[k] = first
 * arena element of segment k, seg_begin[0] = 0; device arrays; n_segs = 0 disables the lookup). */
int tok_sgd_step_dev_groups(long long n, float* param, float* grad, float* momentum_buf, void* shadow_bf16,
                            const float* lr_dev, int* step_dev, float momentum, float weight_decay, float dampening,
                            int nesterov, float grad_scale, int zero_grad, const int* seg_begin,
                            const float* seg_lr_mult, const float* seg_wd_mult, int n_segs, void* stream);
int tok_adam_step_dev_groups(long long n, float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                             void* shadow_bf16, const float* lr_dev, int* step_dev, float beta1, float beta2, float eps,
                             float weight_decay, int decoupled, float grad_scale, int zero_grad, const int* seg_begin,
                             const float* seg_lr_mult, const float* seg_wd_mult, int* seg_steps, int n_segs,
                             void* stream);
/* Padded operand forms for layers whose channel counts are not multiples of 8 and that do NOT run on the halo kernels
 * (timm HRNet fuse / transition convs, torchok/models/backbones/hrnet.py:140-192): dst bf16 [Kp][T][Cp] =
 * zero-padded src fp32 [K][T][C] (T = R*S filter taps; inv_map, nullable: source channel of every padded input position
 * or -1), and the reverse for the weight gradient: dst fp32 [K][T][C] += src fp32 [K][T][Cp] (map, nullable: padded
 * position of every input channel). */
int tok_pad_weight(int K, int T, int C, int Kp, int Cp, const float* src, const int* inv_map, void* dst, void* stream);
int tok_unpad_wgrad_add(int K, int T, int C, int Cp, const float* src, const int* map, float* dst, void* stream);
/* `seg_steps` (nullable, n_segs ints on the device): per-parameter Adam step counts (torch.optim.Adam's state['step']);
 * advanced by this call for every segment whose multipliers are not both zero and used for the bias correction, so a
 * parameter thawed by FreezeUnfreeze (torchok/callbacks/freeze_unfreeze.py:51-184) restarts at step 1 as in the
 * reference.  Segments with lr_mult == wd_mult == 0 (frozen) are skipped entirely by both step kernels. */
int tok_cast_f32_bf16(long long n, const float* src, void* dst, void* stream);

/* ---- Swin-V2 continuous position bias (timm WindowAttention.cpb_mlp + relative_position_index gather + 16 * sigmoid,
 *      used through torchok/models/backbones/swin.py:71-81): coords [(2ws-1)^2][2], w1 [512][2], b1 [512], w2 [heads][512]
 *      -> hidden [(2ws-1)^2][512], table [(2ws-1)^2][heads] (kept for the backward), bias [heads][ws^2][ws^2].  The backward
 *      ACCUMULATES dw1 / db1 / dw2 (fp32) from dbias; dtable is scratch. */
int tok_cpb_bias_fwd(int ws, int heads, int hidden_dim, const float* coords, const float* w1, const float* b1,
                     const float* w2, float* hidden, float* table, float* bias, void* stream);
int tok_cpb_bias_bwd(int ws, int heads, int hidden_dim, const float* coords, const float* w2, const float* hidden,
                     const float* table, const float* dbias, float* dtable, float* dw1, float* db1, float* dw2,
                     void* stream);

/* ---- object-context head and U-Net decoder glue ----------------------------------------------------
 * nearest-neighbour resize written into a channel slice of a padded NHWC concat buffer: F.interpolate(mode='nearest') +
 * torch.cat of DecoderBlock.forward (torchok/models/necks/segmentation/unet.py:40-58); same argument meaning as
 * tok_bilinear_fwd / _bwd. */
int tok_nearest_fwd(int n, int hi, int wi, int c, int ho, int wo, const void* src, void* dst, int dst_c,
                    int dst_c_offset, void* stream);
int tok_nearest_bwd(int n, int hi, int wi, int c, int ho, int wo, const void* dout, int dout_c, int dout_c_offset,
                    void* dsrc, void* stream);
/* out[n, :, c] = x[n, :, c] * scale[n][c]: nn.Dropout2d (ocr.py:126) with the mask drawn by the caller. */
int tok_channel_scale(int n, long long hw, int c, const void* x, const float* scale, void* out, void* stream);
/* SpatialGather_Module.forward (torchok/models/heads/segmentation/ocr.py:37-46): p = softmax over the hw positions of
 * every class map (logits [b][hw][kp] bf16, k <= kp classes); ctx[b][k][c] = sum_hw p * feats[b][hw][c].
 * stats [b][k][2] fp32 (max, sum of exp) is kept for the backward; ctx_f32 [b][k][c] is fp32 scratch. */
int tok_spatial_gather_fwd(int b, int hw, int c, int k, int kp, const void* feats, const void* logits, float* stats,
                           float* ctx_f32, void* ctx, void* stream);
int tok_spatial_gather_bwd(int b, int hw, int c, int k, int kp, const void* feats, const void* logits,
                           const float* stats, const void* ctx, const void* dctx, void* dfeats, void* dlogits,
                           void* stream);
/* ObjectAttentionBlock.forward at scale 1 (ocr.py:77-101): out[b][hw][:] = softmax_k(scale * q[b][hw] . key[b][k]) .
 * value[b][k][:]; q / out [b][hw][kc], key / value [b][k][kc], all bf16.  The backward returns dq and, through fp32
 * scratch, dkey / dvalue. */
int tok_object_attn_fwd(int b, int hw, int kc, int k, float scale, const void* q, const void* key, const void* value,
                        void* out, void* stream);
int tok_object_attn_bwd(int b, int hw, int kc, int k, float scale, const void* q, const void* key, const void* value,
                        const void* dout, void* dq, float* dkey_f32, float* dvalue_f32, void* dkey, void* dvalue,
                        void* stream);

/* ---- data-parallel gradient exchange over NVLink / NVSwitch PEER MEMORY, fused with the optimizer --------------------
 * Replaces Lightning's DDP gradient all-reduce + optimizer.step (`trainer.strategy: ddp`,
 * torchok/constructor/config_structure.py:137-140; examples/configs/classification_imagenet.yaml:121-122).  One process
 * per GPU; every rank allocates its parameter / gradient / bf16-shadow arenas and a flag block with tok_ipc_alloc and
 * opens the other ranks' with tok_ipc_open (CUDA IPC, peer access over NVLink).  tok_peer_step then runs, per gradient
 * bucket, ONE kernel per rank: barrier (flags in peer memory) -> each rank sums ITS 1/world slice of the bucket from
 * all ranks' gradient arenas (reduce-scatter by peer loads) -> SGD / Adam(W) on that slice (momentum / moment state
 * exists only for the slice: ZeRO-1) -> the new fp32 master and bf16 shadow values are stored into EVERY rank's arenas
 * (all-gather by peer stores) -> barrier -> the local gradients of the bucket are cleared.  No host synchronisation, so
 * the launch can be captured in a CUDA graph. */
#define TOK_PEER_MAX_RANKS 8
#define TOK_PEER_MAX_BUCKETS 256
/* cudaMalloc + zero fill + cudaIpcGetMemHandle; `handle64` receives the 64-byte cudaIpcMemHandle_t. */
int tok_ipc_alloc(size_t bytes, void** dev_ptr, void* handle64);
int tok_ipc_free(void* dev_ptr);
/* cudaIpcOpenMemHandle (peer access enabled lazily) of a handle produced by ANOTHER process. */
int tok_ipc_open(const void* handle64, void** dev_ptr);
int tok_ipc_close(void* dev_ptr);
/* bytes a flag block must have (two phases x buckets x ranks 32-bit epochs + per-bucket epoch / ticket words) */
size_t tok_peer_flag_bytes(void);
typedef struct {
  int world, rank;
  float* master[TOK_PEER_MAX_RANKS];     /* fp32 parameter arenas of every rank (index = rank; [rank] is local) */
  float* grad[TOK_PEER_MAX_RANKS];       /* fp32 gradient arenas */
  void* shadow[TOK_PEER_MAX_RANKS];      /* bf16 shadow arenas */
  unsigned* flags[TOK_PEER_MAX_RANKS];   /* flag blocks */
} tokPeerArenas;
/* kind 0: SGD (h0 momentum, h1 weight_decay, h2 dampening, i0 nesterov; state0 = momentum buffer or NULL);
 * kind 1: Adam / AdamW (h0 beta1, h1 beta2, h2 eps, h3 weight_decay, i0 decoupled; state0 exp_avg, state1 exp_avg_sq).
 * [begin, end) = the bucket's element range in the arenas (multiples of 64); `bucket` < TOK_PEER_MAX_BUCKETS indexes
 * the flags; grad_scale multiplies the SUMMED gradient (1/world for DDP's mean); advance_step != 0 on the first bucket
 * launched in a step (advances *step_dev and the per-segment Adam counters). */
int tok_peer_step(const tokPeerArenas* arenas, int bucket, long long begin, long long end, int kind, float* state0,
                  float* state1, const float* lr_dev, int* step_dev, float h0, float h1, float h2, float h3, int i0,
                  float grad_scale, const int* seg_begin, const float* seg_lr_mult, const float* seg_wd_mult,
                  int* seg_steps, int n_segs, int advance_step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOKB200_H_ */
