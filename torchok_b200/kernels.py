"""Host-side launch layer: torch tensors in, libtokb200.so (sm_100a CUDA behind the C ABI) out.

Everything here is plumbing: shape bookkeeping, output allocation through torch's caching allocator, the current CUDA
stream, and `torch.autograd.Function`s whose forward/backward are sequences of C-ABI calls.  No arithmetic on
activations or weights is done by torch ops in this file.

Layout contract: an activation is a bf16 tensor of logical shape (N, C, H, W) whose memory is NHWC ("channels_last")
with a channel pitch Cp that is a multiple of 8 (Cp == C for every ResNet tensor; HRNet's 18/36-channel branches are
views `[:, :C]` of a Cp-channel buffer whose pad lanes are zero).  The reference keeps NCHW fp32/AMP tensors
(torchok/tasks/classification.py:108-119); logical shapes are identical, so reference-side code that looks at
`.shape` keeps working.
"""
import ctypes as C

import os

import torch

from ._lib import lib, tokConvDesc

BF16 = torch.bfloat16
F32 = torch.float32


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def ceil8(c):
    return (c + 7) // 8 * 8


def require_cuda(t, what='tensor'):
    if not t.is_cuda:
        raise RuntimeError(f'torchok_b200: {what} is on {t.device}; the sm_100a kernels need a CUDA tensor '
                           '(there is no CPU fallback — use oracle/ for CPU reference results)')


# ------------------------------------------------------------------------------------------------------ layout
def nhwc_empty(n, c, h, w, device, cp=None):
    """bf16 (n, c, h, w) tensor backed by an NHWC buffer of channel pitch cp (default ceil8(c))."""
    cp = cp or ceil8(c)
    buf = torch.empty((n, h, w, cp), dtype=BF16, device=device)
    t = buf.permute(0, 3, 1, 2)
    return t if cp == c else t[:, :c]


def nhwc_zeros(n, c, h, w, device, cp=None):
    cp = cp or ceil8(c)
    buf = torch.zeros((n, h, w, cp), dtype=BF16, device=device)
    t = buf.permute(0, 3, 1, 2)
    return t if cp == c else t[:, :c]


def nhwc_pitch(x):
    """Channel pitch of a 4-D tensor if its memory is NHWC-with-pitch as produced by this module, else None."""
    if x.dim() != 4 or x.dtype != BF16:
        return None
    n, c, h, w = x.shape
    s = x.stride()
    if c > 1 and s[1] != 1:
        return None
    if w > 1:
        cp = s[3]
    elif h > 1:
        cp = s[2]
    elif n > 1:
        cp = s[0]
    else:
        cp = ceil8(c)
    if cp % 8 or cp < c:
        return None
    if (w > 1 and s[3] != cp) or (h > 1 and s[2] != w * cp) or (n > 1 and s[0] != h * w * cp):
        return None
    if x.data_ptr() % 16:
        return None
    return cp


def to_nhwc(x):
    """Any (N,C,H,W) float tensor -> bf16 NHWC-with-pitch (no-op if it already is)."""
    require_cuda(x, 'activation')
    if nhwc_pitch(x) is not None:
        return x
    n, c, h, w = x.shape
    if x.dtype not in (F32, BF16):
        x = x.float()
    x = x.contiguous()
    out = nhwc_empty(n, c, h, w, x.device)
    lib().tok_nchw_to_nhwc(n, c, h * w, ceil8(c), int(x.dtype == BF16), _p(x), _p(out), _st())
    return out


def to_nchw(x, dtype=F32):
    """bf16 NHWC-with-pitch -> contiguous NCHW tensor of `dtype` (fp32 or bf16)."""
    cp = nhwc_pitch(x)
    if cp is None:
        raise RuntimeError('to_nchw expects an NHWC bf16 activation')
    n, c, h, w = x.shape
    out = torch.empty((n, c, h, w), dtype=dtype, device=x.device)
    lib().tok_nhwc_to_nchw(n, c, h * w, cp, int(dtype == BF16), _p(x), _p(out), _st())
    return out


class _ToNHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.dtype = x.dtype
        return to_nhwc(x)

    @staticmethod
    def backward(ctx, g):
        return to_nchw(to_nhwc(g), ctx.dtype if ctx.dtype in (F32, BF16) else F32).to(ctx.dtype)


class _ToNCHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        return to_nchw(x, dtype)

    @staticmethod
    def backward(ctx, g):
        return to_nhwc(g), None


def to_nhwc_autograd(x):
    if nhwc_pitch(x) is not None:
        return x
    return _ToNHWC.apply(x) if x.requires_grad else to_nhwc(x)


def to_nchw_autograd(x, dtype=F32):
    return _ToNCHW.apply(x, dtype)


# ------------------------------------------------------------------------------------------------------ weights
def cast_bf16(src_f32, dst_bf16=None):
    """fp32 -> bf16 elementwise over the flat storage (memory layout preserved)."""
    if dst_bf16 is None:
        dst_bf16 = torch.empty_like(src_f32, dtype=BF16)
    lib().tok_cast_f32_bf16(src_f32.numel(), _p(src_f32), _p(dst_bf16), _st())
    return dst_bf16


def is_krsc(w):
    """True if a (K,C,R,S) weight tensor is stored [K][R][S][C] densely."""
    return w.permute(0, 2, 3, 1).is_contiguous()


def shadow_of(param):
    """bf16 copy of an fp32 master parameter, same memory order.  Arena-managed parameters carry a persistent
    shadow that the optimizer kernel refreshes (`_tok_shadow`); otherwise the cast runs now."""
    sh = getattr(param, '_tok_shadow', None)
    if sh is not None:
        return sh
    require_cuda(param, 'parameter')
    return cast_bf16(param.detach())


def grad_ready(param):
    """Tell the gradient-bucket reducer (engine.ParamArena) that this parameter's gradient is complete."""
    b = getattr(param, '_tok_bucket', None)
    if b is not None:
        b[0].ready(b[1])


def grad_buffer(param):
    """fp32 gradient accumulator of a parameter (same memory order); created zero-filled on first use."""
    if param.grad is None:
        param.grad = torch.zeros_like(param)  # preserve_format keeps [K][R][S][C]
    return param.grad


# ------------------------------------------------------------------------------------------------------ side stream
# Weight gradients have no consumer inside the backward pass (only the gradient exchange / optimizer read them), so they
# are launched on a side stream forked off the main one: in the captured step graph they become parallel branches that
# run next to the HBM-bound BatchNorm / LayerNorm passes of the following units instead of in line with them.
_WGRAD_ASYNC = os.environ.get('TOK_WGRAD_STREAM', '1') == '1'
_side_streams, _side_state = {}, {'pending': False, 'keep': [], 'queued': False}


def side_stream():
    dev = torch.cuda.current_device()
    s = _side_streams.get(dev)
    if s is None:
        s = _side_streams[dev] = torch.cuda.Stream()
    return s


def join_side():
    """The current stream waits for everything launched on the side stream; operands kept alive for it are released."""
    if _side_state['pending']:
        torch.cuda.current_stream().wait_stream(side_stream())
        _side_state['pending'] = False
    _side_state['keep'].clear()
    _side_state['queued'] = False


def wgrad_async(operands, launch):
    """Run `launch(stream_ptr)` (one weight-gradient kernel) on the side stream, ordered after the work enqueued on the
    current stream so far.  `operands` stay referenced until the join, so the caching allocator cannot hand their
    memory to a later main-stream allocation while the side stream still reads it.  The join happens when a gradient
    bucket is handed to the exchange, in ParamArena.finish(), and — for plain autograd use — in a callback at the end of
    the backward pass."""
    if not _WGRAD_ASYNC:
        launch(_st())
        return
    cur, side = torch.cuda.current_stream(), side_stream()
    side.wait_stream(cur)
    _side_state['keep'].extend(operands)
    launch(side.cuda_stream)
    _side_state['pending'] = True
    if not _side_state['queued']:
        try:
            torch.autograd.Variable._execution_engine.queue_callback(join_side)
            _side_state['queued'] = True
        except RuntimeError:      # not inside a backward pass (direct call): order the streams now
            join_side()


# ------------------------------------------------------------------------------------------------------ conv
def conv_desc(n, h, w, c, k, r, s, stride, pad, dil, wk=0, wc=0):
    """`wk` / `wc`: dimensions of an UNPADDED weight tensor (0 = k / c); see include/tokb200.h tokConvDesc."""
    d = tokConvDesc(n, h, w, c, k, r, s, stride, pad, dil, wk, wc)
    p, q = C.c_int(), C.c_int()
    lib().tok_conv_out_hw(C.byref(d), C.byref(p), C.byref(q))
    return d, p.value, q.value


def conv_fprop(d, x, w, y, stats=None, addend=None, bias=None, relu=False):
    lib().tok_conv_fprop(C.byref(d), _p(x), _p(w), _p(y), _p(stats[0]) if stats is not None else None,
                         _p(stats[1]) if stats is not None else None, _p(addend), _p(bias), int(relu), _st())


def conv_dgrad(d, dy, w, dx, addend=None, addend_bits=None):
    """`addend_bits`: 1-bit ReLU mask of `addend` (tok_conv_dgrad_masked: dx = dgrad(dy) + addend * mask)."""
    nbytes = lib().tok_conv_dgrad_workspace_bytes(C.byref(d))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dy.device) if nbytes else None
    if addend_bits is not None:
        lib().tok_conv_dgrad_masked(C.byref(d), _p(dy), _p(w), _p(dx), _p(addend), _p(addend_bits), _p(ws), _st())
    else:
        lib().tok_conv_dgrad(C.byref(d), _p(dy), _p(w), _p(dx), _p(addend), _p(ws), _st())


# TOK_MASKED_ADDEND=0: the residual gradient of a block is materialised by the tail's BatchNorm backward (r2 path)
_MASKED_ADDEND = os.environ.get('TOK_MASKED_ADDEND', '1') != '0'


def dgrad_masked_supported(d):
    return _MASKED_ADDEND and bool(lib().tok_conv_dgrad_masked_supported(C.byref(d)))


def conv_wgrad(d, x, dy, dw):
    lib().tok_conv_wgrad(C.byref(d), _p(x), _p(dy), _p(dw), _st())


# ------------------------------------------------------------------------------------------------------ BN state
class BNState:
    """What a BatchNorm2d module lends to the fused conv+BN unit (torch.nn.BatchNorm2d semantics,
    torchok/models/modules/bricks/convbnact.py:44-46)."""
    __slots__ = ('weight', 'bias', 'running_mean', 'running_var', 'eps', 'momentum', 'training', 'acc', 'cp', 'cv')

    def __init__(self, weight, bias, running_mean, running_var, eps, momentum, training, acc, cp, cv=None):
        """`cp`: channel pitch of the activations (multiple of 8); `cv`: channels that exist in weight / bias / running
        statistics (HRNet's BatchNorm2d(18): cp 24, cv 18 — the *_cv finalize kernels keep the pad lanes at zero)."""
        self.weight, self.bias, self.running_mean, self.running_var = weight, bias, running_mean, running_var
        self.eps, self.momentum, self.training, self.acc, self.cp = eps, momentum, training, acc, cp
        self.cv = cp if cv is None else cv


MASK_NONE, MASK_Y, MASK_BITS = 0, 1, 2


# Fused BatchNorm finalize (last-CTA ticket).  A/B on one B200, ResNet-50 bs256 graph replay (profiles/r1_resnet50_step.md
# section 8): forward fusion 21.64-21.68 ms vs 21.25-21.42 ms unfused (the fence + ticket + finalize tail of the
# persistent conv kernel costs more than the 4 us single-CTA launch it replaces), backward fusion neutral (21.39) and
# 53 launches fewer for the eager multi-GPU loop.  Hence: forward off, backward on.
_FUSE_FWD_FIN = os.environ.get('TOK_BN_FUSE_FWD', '0') == '1'
_FUSE_BWD_FIN = os.environ.get('TOK_BN_FUSE_BWD', '1') == '1'
# r2 experiment (opt-in, TOK_BN_FUSE_APPLY=1): the forward finalize inside the APPLY pass (every CTA derives scale / shift
# from the sums; a ticket only to zero them).  Correct (the GPU suite passes with it) but SLOWER: 22.19 vs 20.25 ms per
# ResNet-50 step — the apply grids have ~2 400 CTAs and their tickets serialise on one L2 address (~20 ns each), which
# costs more than the 53 single-CTA finalize launches it removes.  Kept off.
_FUSE_APPLY_FIN = os.environ.get('TOK_BN_FUSE_APPLY', '0') == '1'
# r2 "chain" form of the same fusion (opt-in, TOK_BN_FUSE_CHAIN=1): no ticket — the batch sums of a unit are zeroed by the
# NEXT fused apply launch of the stream (CTA 0 of it), which is ordered after every reader of those sums;
# `_chain_pending` is the accumulator tensor the last chain launch left non-zero.  It removes the single-CTA finalize
# launch of every unit (53 per ResNet-50 step, 306 per HRNet step) and the ticket atomics — and is STILL slower, on all
# three workloads (A/B on one B200, graph replay): ResNet-50 21.70 vs 20.08 ms, HRNet-W18 seg 53.7 vs 48.7 ms,
# ResNet-18 CIFAR 1.298 vs 1.267 ms.  So it was never the ticket: every CTA of the apply grid paying ~32 dependent L2
# loads + a rsqrt per thread before its first vector costs more than one 4 us launch.  Kept off.
_FUSE_CHAIN = os.environ.get('TOK_BN_FUSE_CHAIN', '0') == '1'
# r2 experiment (opt-in, TOK_BN_FUSED_BWD_MB=<megabytes>): BatchNorm backward of an L2-sized tensor (g + y up to that size)
# as ONE launch — reduce, finalize in the last CTA, grid-wide release, apply on the rows each CTA already read
# (tok_bn_bwd_fused_cv).  Correct (the GPU suite passes with it) and SLOWER on every workload: ResNet-50 19.94 (110 MB) /
# 20.03 (60 MB) vs 19.85 ms off, HRNet-W18 seg 52.4 vs 48.3 ms, ResNet-18 CIFAR 1.293 vs 1.267 ms; 48 us per fused launch
# against 20 + 21 us for the pair it replaces — the one-wave grid at 2 CTAs / SM (100-128 registers) waiting for its
# slowest CTA plus the release round trip costs more than the second read from L2 saves.  Kept off.
_FUSED_BWD_MB = float(os.environ.get('TOK_BN_FUSED_BWD_MB', '0'))
_chain_pending = [None]


def chain_flush():
    """Zero the accumulators the last chain launch left behind (needed only by code that reads / re-uses them outside
    the chain, e.g. tests that inspect them)."""
    acc = _chain_pending[0]
    if acc is not None:
        acc[0:2].zero_()
        _chain_pending[0] = None


def unit_forward(x, d, pq, w, bn, relu, residual=None, keep=True):
    """conv -> BatchNorm (batch stats in training) -> (+residual) -> (ReLU).

    x: NHWC activation; d: tokConvDesc for x; w: bf16 [Kp][R][S][Cp] weights; bn: BNState.
    Returns (out, saved) where saved = (x, y, bits_or_None, small, mask_mode) feeds `unit_backward`.  The backward
    never re-reads `out`: a plain ReLU unit rebuilds its mask from y and the forward's scale/shift (MASK_Y), a
    residual tail stores one bit per element (MASK_BITS).
    """
    L = lib()
    n, kp, (p, q) = d.n, d.k, pq
    dev = x.device
    y = torch.empty((n, p, q, kp), dtype=BF16, device=dev)
    small = torch.empty((4, kp), dtype=F32, device=dev)  # scale, shift, save_mean, save_invstd
    rows = n * p * q
    st = _st()
    fused_apply = False
    chain = bn.training and _FUSE_CHAIN and not _FUSE_APPLY_FIN and not _FUSE_FWD_FIN and \
        (keep and relu and residual is not None or L.tok_bn_apply_train_supported(rows, kp))
    if chain:
        acc = bn.acc
        if _chain_pending[0] is acc:     # the same layer twice in a row: its sums must be zero before the conv adds to them
            chain_flush()
        L.tok_conv_fprop(C.byref(d), _p(x), _p(w), _p(y), _p(acc[0]), _p(acc[1]), None, None, 0, st)
    elif bn.training and bn.acc.shape[0] > 4 and _FUSE_APPLY_FIN and not _FUSE_FWD_FIN and bn.cv == kp and \
            (keep and relu and residual is not None or L.tok_bn_apply_train_supported(rows, kp)):
        # conv (+ statistics), then ONE pass that finalizes the statistics and applies them (acc[4] word 2: its ticket)
        acc = bn.acc
        L.tok_conv_fprop(C.byref(d), _p(x), _p(w), _p(y), _p(acc[0]), _p(acc[1]), None, None, 0, st)
        fused_apply = True
    elif bn.training:
        acc = bn.acc
        if acc.shape[0] > 4 and _FUSE_FWD_FIN and bn.cv == kp:   # conv + statistics + finalize in one launch (acc[4]: ticket counters)
            L.tok_conv_fprop_bn(C.byref(d), _p(x), _p(w), _p(y), _p(acc[0]), _p(acc[1]), _p(bn.weight), _p(bn.bias),
                                bn.eps, bn.momentum, _p(bn.running_mean), _p(bn.running_var), _p(small[0]),
                                _p(small[1]), _p(small[2]), _p(small[3]), _p(acc[4]), st)
        else:
            L.tok_conv_fprop(C.byref(d), _p(x), _p(w), _p(y), _p(acc[0]), _p(acc[1]), None, None, 0, st)
            L.tok_bn_finalize_train_cv(kp, bn.cv, float(rows), _p(acc[0]), _p(acc[1]), _p(bn.weight), _p(bn.bias), bn.eps,
                                       bn.momentum, _p(bn.running_mean), _p(bn.running_var), _p(small[0]), _p(small[1]),
                                       _p(small[2]), _p(small[3]), st)
    else:
        L.tok_conv_fprop(C.byref(d), _p(x), _p(w), _p(y), None, None, None, None, 0, st)
        L.tok_bn_finalize_eval_cv(kp, bn.cv, _p(bn.running_mean), _p(bn.running_var), _p(bn.weight), _p(bn.bias), bn.eps,
                                  _p(small[0]), _p(small[1]), st)
    out = torch.empty_like(y) if keep else y
    bits, mode = None, MASK_NONE
    fin_args = ()
    if fused_apply:
        fin_args = (_p(acc[0]), _p(acc[1]), _p(bn.weight), _p(bn.bias), bn.eps, bn.momentum, _p(bn.running_mean),
                    _p(bn.running_var), _p(small[0]), _p(small[1]), _p(small[2]), _p(small[3]), acc[4].data_ptr() + 8)
    if chain:
        prev = _chain_pending[0]
        chain_args = (_p(acc[0]), _p(acc[1]), _p(bn.weight), _p(bn.bias), bn.eps, bn.momentum, _p(bn.running_mean),
                      _p(bn.running_var), _p(small[0]), _p(small[1]), _p(small[2]), _p(small[3]),
                      _p(prev[0]) if prev is not None else None, 2 * prev.shape[1] if prev is not None else 0)
        _chain_pending[0] = acc
    if keep and relu and residual is not None:
        bits = torch.empty((rows * kp // 8,), dtype=torch.uint8, device=dev)
        mode = MASK_BITS
        if chain:
            L.tok_bn_apply_bits_chain(rows, kp, bn.cv, _p(y), *chain_args, _p(residual), _p(out), _p(bits), st)
        elif fused_apply:
            L.tok_bn_apply_bits_train(rows, kp, _p(y), *fin_args, _p(residual), _p(out), _p(bits), st)
        else:
            L.tok_bn_apply_bits(rows, kp, _p(y), _p(small[0]), _p(small[1]), _p(residual), _p(out), _p(bits), st)
    else:
        if relu:
            mode = MASK_Y
        if chain:
            L.tok_bn_apply_chain(rows, kp, bn.cv, _p(y), *chain_args, _p(residual), int(relu), _p(out), st)
        elif fused_apply:
            L.tok_bn_apply_train(rows, kp, _p(y), *fin_args, _p(residual), int(relu), _p(out), st)
        else:
            L.tok_bn_apply(rows, kp, _p(y), _p(small[0]), _p(small[1]), _p(residual), int(relu), _p(out), st)
    saved = (x, y, bits, small, mode) if keep else None
    return out.permute(0, 3, 1, 2), saved


def unit_backward(saved, d, w, bn, dout, dout2=None, need_dx=True, dx_addend=None, want_dres=False,
                  wgrad_into=None, dgamma=None, dbeta=None, compact_dx=False, wgrad_direct=False, dx_addend_bits=None):
    """Backward of `unit_forward`.  Returns (dx or None, dres or None).  Parameter gradients are ACCUMULATED into
    wgrad_into / dgamma / dbeta (fp32, may be None for frozen parameters)."""
    L = lib()
    x, y, bits, small, mode = saved
    kp = d.k
    rows = y.numel() // kp
    dev = y.device
    st = _st()
    acc = bn.acc
    coefs = torch.empty((3, kp), dtype=F32, device=dev)
    if not bn.training:
        # Frozen statistics (torchok/callbacks/freeze_unfreeze.py puts BatchNorm into eval mode while the layers around
        # it train): out = y * scale + shift with constants, so dy = scale * g; gamma / beta still receive
        # dgamma = sum g * (y - running_mean) * invstd, dbeta = sum g when they are trainable.
        if dgamma is not None or dbeta is not None:
            L.tok_bn_bwd_reduce2(rows, kp, _p(dout), _p(dout2), _p(y), mode, _p(bits), _p(small[0]), _p(small[1]),
                                 _p(acc[2]), _p(acc[3]), st)
            invstd = torch.rsqrt(bn.running_var.float() + bn.eps)
            cv = bn.cv
            if dgamma is not None:
                dgamma += (acc[3, :cv] - bn.running_mean.float() * acc[2, :cv]) * invstd
            if dbeta is not None:
                dbeta += acc[2, :cv]
            acc[2:4].zero_()
        coefs[0].copy_(small[0])
        coefs[1:].zero_()
        dy = torch.empty_like(y)
        dres = torch.empty_like(y) if want_dres else None
        L.tok_bn_bwd_apply2(rows, kp, _p(dout), _p(dout2), _p(y), mode, _p(bits), _p(small[0]), _p(small[1]),
                            _p(coefs[0]), _p(coefs[1]), _p(coefs[2]), _p(dy), _p(dres), st)
        return _unit_backward_tail(L, st, d, w, x, y, dy, dres, need_dx, compact_dx, dx_addend, wgrad_into, rows, kp, dev,
                                   wgrad_direct, dx_addend_bits)
    if acc.shape[0] > 4 and _FUSE_BWD_FIN and _FUSED_BWD_MB > 0 and \
            rows * kp * 2 * (3 if dout2 is not None else 2) <= _FUSED_BWD_MB * 1e6:
        # small enough to stay in L2 between the passes: reduce + finalize + apply in ONE launch (acc[4] words 1 / 3:
        # ticket and release word of the layer)
        dy = torch.empty_like(y)
        dres = torch.empty_like(y) if want_dres else None
        L.tok_bn_bwd_fused_cv(rows, kp, bn.cv, _p(dout), _p(dout2), _p(y), mode, _p(bits), _p(small[0]), _p(small[1]),
                              _p(acc[2]), _p(acc[3]), _p(small[2]), _p(small[3]), _p(bn.weight), _p(coefs[0]),
                              _p(coefs[1]), _p(coefs[2]), _p(dgamma), _p(dbeta), 1, acc[4].data_ptr() + 4,
                              acc[4].data_ptr() + 12, _p(dy), _p(dres), st)
        return _unit_backward_tail(L, st, d, w, x, y, dy, dres, need_dx, compact_dx, dx_addend, wgrad_into, rows, kp, dev,
                                   wgrad_direct, dx_addend_bits)
    if acc.shape[0] > 4 and _FUSE_BWD_FIN:   # reduce + finalize in one launch (acc[4]: the layer's ticket counters)
        L.tok_bn_bwd_reduce2_finalize_cv(rows, kp, bn.cv, _p(dout), _p(dout2), _p(y), mode, _p(bits), _p(small[0]),
                                         _p(small[1]), _p(acc[2]), _p(acc[3]), _p(small[2]), _p(small[3]), _p(bn.weight),
                                         _p(coefs[0]), _p(coefs[1]), _p(coefs[2]), _p(dgamma), _p(dbeta), 1,
                                         acc[4].data_ptr() + 4, st)
    else:
        L.tok_bn_bwd_reduce2(rows, kp, _p(dout), _p(dout2), _p(y), mode, _p(bits), _p(small[0]), _p(small[1]),
                             _p(acc[2]), _p(acc[3]), st)
        L.tok_bn_bwd_finalize_cv(kp, bn.cv, float(rows), _p(acc[2]), _p(acc[3]), _p(small[2]), _p(small[3]), _p(bn.weight),
                                 _p(coefs[0]), _p(coefs[1]), _p(coefs[2]), _p(dgamma), _p(dbeta), 1, st)
    dy = torch.empty_like(y)
    dres = torch.empty_like(y) if want_dres else None
    L.tok_bn_bwd_apply2(rows, kp, _p(dout), _p(dout2), _p(y), mode, _p(bits), _p(small[0]), _p(small[1]),
                        _p(coefs[0]), _p(coefs[1]), _p(coefs[2]), _p(dy), _p(dres), st)
    return _unit_backward_tail(L, st, d, w, x, y, dy, dres, need_dx, compact_dx, dx_addend, wgrad_into, rows, kp, dev,
                               wgrad_direct, dx_addend_bits)


def _unit_backward_tail(L, st, d, w, x, y, dy, dres, need_dx, compact_dx, dx_addend, wgrad_into, rows, kp, dev,
                        wgrad_direct=False, dx_addend_bits=None):
    """Data and weight gradients of the conv once dy (the gradient at the conv output) is known."""
    dx = None
    if need_dx and compact_dx:
        # strided 1x1 conv: the data gradient lives on the (p*stride, q*stride) sub-lattice only; return it compact
        # ([N, P, Q, C] = a plain GEMM) and let the caller merge it with tok_strided_add
        assert d.r == 1 and d.s == 1 and d.pad == 0 and dx_addend is None
        dx = torch.empty((d.n, y.shape[1], y.shape[2], d.c), dtype=BF16, device=dev)
        L.tok_linear_dgrad(rows, kp, d.c, _p(dy), _p(w), _p(dx), st)
    elif need_dx:
        dx = torch.empty((d.n, d.h, d.w, d.c), dtype=BF16, device=dev)
        conv_dgrad(d, dy, w, dx, dx_addend, dx_addend_bits)
        dx = dx.permute(0, 3, 1, 2)
    if wgrad_into is not None:
        if wgrad_direct:
            wgrad_async((x, dy, wgrad_into), lambda s: L.tok_conv_wgrad(C.byref(d), _p(x), _p(dy), _p(wgrad_into), s))
        else:   # a padded temporary that torch ops merge on the main stream right after: stay in line
            L.tok_conv_wgrad(C.byref(d), _p(x), _p(dy), _p(wgrad_into), st)
    return dx, (dres.permute(0, 3, 1, 2) if dres is not None else None)


def strided_add(dst, src_compact, stride):
    """dst[n, :, p*stride, q*stride] += src_compact[n, p, q, :]  (dst: NHWC-backed (N,C,H,W); src: [N,P,Q,C])."""
    n, c, h, w = dst.shape
    lib().tok_strided_add(n, h, w, nhwc_pitch(dst), stride, _p(src_compact), _p(dst), _st())
    return dst


# ------------------------------------------------------------------------------------------------------ pooling
def maxpool_fwd(x, k, s, pad, want_arg=True):
    n, c, h, w = x.shape
    cp = nhwc_pitch(x)
    p, q = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
    out = torch.empty((n, p, q, cp), dtype=BF16, device=x.device)
    arg = torch.empty((n, p, q, cp), dtype=torch.uint8, device=x.device)
    lib().tok_maxpool_fwd(n, h, w, cp, k, s, pad, _p(x), _p(out), _p(arg), _st())
    o = out.permute(0, 3, 1, 2)
    return (o if cp == c else o[:, :c]), arg


def maxpool_bwd(dout, arg, xshape, cp, k, s, pad):
    n, c, h, w = xshape
    dx = torch.empty((n, h, w, cp), dtype=BF16, device=dout.device)
    lib().tok_maxpool_bwd(n, h, w, cp, k, s, pad, _p(dout), _p(arg), _p(dx), _st())
    o = dx.permute(0, 3, 1, 2)
    return o if cp == c else o[:, :c]


class MaxPoolFn(torch.autograd.Function):
    """torch.nn.MaxPool2d (torchok/models/backbones/resnet.py:510)."""

    @staticmethod
    def forward(ctx, x, k, s, pad):
        x = to_nhwc(x)
        out, arg = maxpool_fwd(x, k, s, pad)
        ctx.save_for_backward(arg)
        ctx.meta = (tuple(x.shape), nhwc_pitch(x), k, s, pad)
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        shape, cp, k, s, pad = ctx.meta
        return maxpool_bwd(_dense_grad(g, cp), arg, shape, cp, k, s, pad), None, None, None


def _dense_grad(g, cp=None):
    """Incoming gradient -> NHWC bf16 with the producer's pitch (autograd may hand us anything)."""
    g = to_nhwc(g) if nhwc_pitch(g) is None else g
    if cp is not None and nhwc_pitch(g) != cp:
        n, c, h, w = g.shape
        buf = nhwc_zeros(n, c, h, w, g.device, cp)
        buf.copy_(g)
        g = buf
    return g


POOL_MODES = {'avg': 0, 'max': 1, 'avgmax': 2}


class GlobalPoolFn(torch.autograd.Function):
    """timm SelectAdaptivePool2d(output_size=1, flatten=True) (torchok/models/poolings/classification/pooling.py:8-12).
    Output: (N, C) bf16."""

    @staticmethod
    def forward(ctx, x, mode):
        x = to_nhwc(x)
        n, c, h, w = x.shape
        cp = nhwc_pitch(x)
        out = torch.empty((n, cp), dtype=BF16, device=x.device)
        lib().tok_gap_fwd(n, h * w, cp, POOL_MODES[mode], _p(x), _p(out), _st())
        ctx.meta = (n, c, h, w, cp, mode)
        if mode != 'avg':
            ctx.save_for_backward(x, out)
        return out if cp == c else out[:, :c]

    @staticmethod
    def backward(ctx, g):
        n, c, h, w, cp, mode = ctx.meta
        if cp != c:
            gp = torch.zeros((n, cp), dtype=BF16, device=g.device)
            gp[:, :c] = g
            g = gp
        g = g.to(BF16).contiguous()
        dx = torch.empty((n, h, w, cp), dtype=BF16, device=g.device)
        if mode == 'avg':
            lib().tok_gap_bwd(n, h * w, cp, _p(g), _p(dx), _st())
        else:
            x, _ = ctx.saved_tensors
            lib().tok_gap_bwd_max(n, h * w, cp, POOL_MODES[mode], _p(g), _p(x), _p(dx), _st())
        o = dx.permute(0, 3, 1, 2)
        return (o if cp == c else o[:, :c]), None


# ------------------------------------------------------------------------------------------------------ linear
class ResidualLink:
    """Side channel for the gradient of the skip connection of a residual block `x + f(x)`.

    Autograd would sum the two gradients of `x` (skip path, branch path) with a separate elementwise pass.  Instead
    the block's tail (LayerNormFn with `res_link`) parks the skip gradient here and returns no gradient for its
    residual input, and the first linear layer of the branch (`linear(..., res_link=...)`, whose backward always runs
    after the tail's) adds it in the epilogue of its dgrad GEMM."""
    __slots__ = ('g',)

    def __init__(self):
        self.g = None

    def take(self):
        g, self.g = self.g, None
        return g


def _dgrad_with_link(L, link, m, n, k, g, w, dx, st):
    add = link.take() if link is not None else None
    if add is not None:
        L.tok_linear_dgrad_add(m, n, k, _p(g), _p(w), _p(add), _p(dx), st)
    else:
        L.tok_linear_dgrad(m, n, k, _p(g), _p(w), _p(dx), st)


class LinearFn(torch.autograd.Function):
    """torch.nn.Linear (torchok/models/heads/representation/linear_head.py:25-31).  x (M, K) bf16, weight (N, K)
    fp32 master (bf16 shadow used), bias fp32.  N is padded to a multiple of 8 inside; K must be one."""

    @staticmethod
    def forward(ctx, x, weight, bias, bias_grad_external=False, res_link=None):
        require_cuda(x, 'linear input')
        m, k = x.shape
        n = weight.shape[0]
        ctx.bias_grad_external = bool(bias_grad_external)
        ctx.res_link = res_link
        if k % 8:
            raise ValueError(f'linear: in_features must be a multiple of 8 (got {k})')
        np_ = ceil8(n)
        x = x.to(BF16).contiguous()
        w = padded_linear_shadow(weight)
        b = bias
        if bias is not None and np_ != n:
            b = torch.zeros(np_, dtype=F32, device=x.device)
            b[:n] = bias.detach()
        elif bias is not None:
            b = bias.detach().float()
        y = torch.empty((m, np_), dtype=BF16, device=x.device)
        lib().tok_linear_fwd(m, np_, k, _p(x), _p(w), _p(b), _p(y), _st())
        ctx.save_for_backward(x, w)
        ctx.params = (weight, bias)
        ctx.dims = (m, n, np_, k)
        return y if np_ == n else y[:, :n]

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        weight, bias = ctx.params
        m, n, np_, k = ctx.dims
        L = lib()
        st = _st()
        if np_ != n or g.dtype != BF16 or not g.is_contiguous():
            gp = torch.zeros((m, np_), dtype=BF16, device=g.device) if np_ != n else None
            if gp is not None:
                gp[:, :n] = g
                g = gp
            else:
                g = g.to(BF16).contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((m, k), dtype=BF16, device=g.device)
            _dgrad_with_link(L, ctx.res_link, m, np_, k, g, w, dx, st)
        if weight.requires_grad:
            gw = grad_buffer(weight)
            if np_ == n:
                wgrad_async((x, g, gw), lambda s: L.tok_linear_wgrad(m, np_, k, _p(x), _p(g), _p(gw), s))
            else:
                tmp = torch.zeros((np_, k), dtype=F32, device=g.device)
                L.tok_linear_wgrad(m, np_, k, _p(x), _p(g), _p(tmp), st)
                gw += tmp[:n]
        if bias is not None and bias.requires_grad and not ctx.bias_grad_external:
            acc = torch.zeros((2, np_), dtype=F32, device=g.device)
            L.tok_bn_bwd_reduce(m, np_, _p(g), None, None, _p(g), _p(acc[0]), _p(acc[1]), st)
            grad_buffer(bias).add_(acc[0, :n])
        grad_ready(weight)
        if bias is not None:
            grad_ready(bias)   # with bias_grad_external the consumer's backward (which ran first) has filled it
        return dx, None, None, None, None


def padded_linear_shadow(weight):
    n, k = weight.shape
    np_ = ceil8(n)
    if np_ == n:
        return shadow_of(weight)
    w = torch.zeros((np_, k), dtype=BF16, device=weight.device)
    cast_bf16(weight.detach().contiguous(), w[:n])
    return w


def linear(x, weight, bias=None, bias_grad_external=False, res_link=None):
    """`bias_grad_external`: the op consuming this output (layernorm / gelu with `colsum_param=bias`) accumulates the
    bias gradient as the column sums of its own input gradient, so this layer skips its pass over dy.
    `res_link`: see ResidualLink."""
    return LinearFn.apply(x, weight, bias, bias_grad_external, res_link)


# ------------------------------------------------------------------------------------------------------ loss
class SoftmaxXentFn(torch.autograd.Function):
    """torch.nn.CrossEntropyLoss(reduction='mean') (torchok/losses/__init__.py:26) on (rows, C) logits.
    `correct` (optional int32 scalar tensor) is incremented by the number of rows whose argmax equals the target."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index, correct):
        require_cuda(logits, 'logits')
        rows, c = logits.shape
        lg = logits
        if lg.dtype != BF16 or lg.stride(1) != 1:
            lg = lg.to(BF16).contiguous()
        target = target.long().contiguous()
        loss = torch.zeros((), dtype=F32, device=lg.device)
        lib().tok_softmax_xent(rows, c, lg.stride(0), _p(lg), _p(target), _p(loss), None, 1.0 / rows, 0.0, None,
                               ignore_index, _p(correct), _st())
        ctx.save_for_backward(lg, target)
        ctx.meta = (rows, c, ignore_index, logits.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        lg, target = ctx.saved_tensors
        rows, c, ignore_index, in_dtype = ctx.meta
        ld = lg.stride(0)
        dlog = torch.empty((rows, ld), dtype=BF16, device=lg.device)
        if ld != c:
            dlog.zero_()
        # g is the scalar JointLoss weight chain (torchok/losses/base.py:80); read on the device by the kernel.
        g = g.to(F32).contiguous()
        lib().tok_softmax_xent(rows, c, ld, _p(lg), _p(target), None, _p(dlog), 0.0, 1.0 / rows, _p(g), ignore_index,
                               None, _st())
        d = dlog if ld == c else dlog[:, :c]
        return (d if in_dtype == BF16 else d.to(in_dtype)), None, None, None


def softmax_xent(logits, target, ignore_index=-100, correct=None):
    return SoftmaxXentFn.apply(logits, target, ignore_index, correct)


class RowNormFn(torch.autograd.Function):
    """F.normalize(x, p=2, dim=1) on (rows, d) bf16/fp32 input -> bf16 (linear_head.py:33-35)."""

    @staticmethod
    def forward(ctx, x):
        require_cuda(x, 'embeddings')
        if x.dtype not in (BF16, F32):
            x = x.float()
        x = x.contiguous()
        rows, d = x.shape
        out = torch.empty((rows, d), dtype=BF16, device=x.device)
        inv = torch.empty((rows,), dtype=F32, device=x.device)
        lib().tok_rownorm_fwd(rows, d, _p(x), int(x.dtype == BF16), 1.0, _p(out), d, _p(inv), _st())
        ctx.save_for_backward(x, inv)
        return out

    @staticmethod
    def backward(ctx, g):
        x, inv = ctx.saved_tensors
        rows, d = x.shape
        if g.dtype not in (BF16, F32):
            g = g.float()
        g = g.contiguous()
        dx = torch.empty((rows, d), dtype=x.dtype, device=x.device)
        lib().tok_rownorm_bwd(rows, d, _p(x), int(x.dtype == BF16), _p(inv), 1.0, _p(g), int(g.dtype == BF16), d,
                              _p(dx), int(x.dtype == BF16), 0, _st())
        return dx


def l2_normalize(x):
    return RowNormFn.apply(x)


class ArcFaceFn(torch.autograd.Function):
    """Training branch of ArcFaceHead.forward (arcface_head.py:125-128): scale * where(onehot, phi(cos), cos) with
    cos = normalize(x) @ normalize(W)^T.  x (B, D) bf16/fp32, weight (C, D) fp32 master, target (B,) int64."""

    @staticmethod
    def forward(ctx, x, weight, target, scale, margin, easy_margin):
        require_cuda(x, 'ArcFace input')
        L = lib()
        st = _st()
        if x.dtype not in (BF16, F32):
            x = x.float()
        x = x.contiguous()
        b, d = x.shape
        c = weight.shape[0]
        dp, cp = ceil8(d), ceil8(c)
        dev = x.device
        w = weight.detach().contiguous()
        xs = torch.empty((b, dp), dtype=BF16, device=dev)        # scale * x_hat
        x_inv = torch.empty((b,), dtype=F32, device=dev)
        wh = torch.zeros((cp, dp), dtype=BF16, device=dev)       # w_hat (pad rows stay zero)
        w_inv = torch.empty((c,), dtype=F32, device=dev)
        L.tok_rownorm_fwd(b, d, _p(x), int(x.dtype == BF16), float(scale), _p(xs), dp, _p(x_inv), st)
        L.tok_rownorm_fwd(c, d, _p(w), 0, 1.0, _p(wh), dp, _p(w_inv), st)
        logits = torch.empty((b, cp), dtype=BF16, device=dev)
        L.tok_linear_fwd(b, cp, dp, _p(xs), _p(wh), None, _p(logits), st)
        target = target.long().contiguous()
        cos_t = torch.zeros((b,), dtype=F32, device=dev)
        L.tok_arcface_margin_fwd(b, dp, dp, _p(xs), _p(wh), _p(target), c, _p(logits), cp, float(scale),
                                 float(margin), int(easy_margin), _p(cos_t), st)
        ctx.save_for_backward(x, xs, x_inv, wh, w_inv, target, cos_t)
        ctx.meta = (weight, b, d, c, dp, cp, float(scale), float(margin), int(easy_margin))
        return logits if cp == c else logits[:, :c]

    @staticmethod
    def backward(ctx, g):
        x, xs, x_inv, wh, w_inv, target, cos_t = ctx.saved_tensors
        weight, b, d, c, dp, cp, scale, margin, easy = ctx.meta
        L = lib()
        st = _st()
        dev = x.device
        dl = torch.zeros((b, cp), dtype=BF16, device=dev) if cp != c else torch.empty((b, cp), dtype=BF16, device=dev)
        dl[:, :c] = g  # private copy: the target column is rescaled in place
        L.tok_arcface_margin_bwd(b, _p(target), c, _p(cos_t), _p(dl), cp, scale, margin, easy, st)
        dx = None
        if ctx.needs_input_grad[0]:
            dxs = torch.empty((b, dp), dtype=BF16, device=dev)   # gradient w.r.t. xs = scale * x_hat
            L.tok_linear_dgrad(b, cp, dp, _p(dl), _p(wh), _p(dxs), st)
            dx = torch.empty((b, d), dtype=x.dtype, device=dev)
            L.tok_rownorm_bwd(b, d, _p(x), int(x.dtype == BF16), _p(x_inv), scale, _p(dxs), 1, dp, _p(dx),
                              int(x.dtype == BF16), 0, st)
        if weight.requires_grad:
            dwh = torch.zeros((cp, dp), dtype=F32, device=dev)   # gradient w.r.t. w_hat
            L.tok_linear_wgrad(b, cp, dp, _p(xs), _p(dl), _p(dwh), st)
            gw = grad_buffer(weight)
            L.tok_rownorm_bwd(c, d, _p(weight.detach().contiguous()), 0, _p(w_inv), 1.0, _p(dwh), 0, dp, _p(gw), 0, 1, st)
            grad_ready(weight)
        return dx, None, None, None, None, None


def arcface(x, weight, target, scale, margin, easy_margin=False):
    return ArcFaceFn.apply(x, weight, target, scale, margin, easy_margin)


class ContrastiveFn(torch.autograd.Function):
    """ContrastiveLoss.calc_loss (losses/representation/pairwise.py:126-136) -> per-row loss (B,) fp32."""

    @staticmethod
    def forward(ctx, emb1, emb2, R, margin):
        require_cuda(emb1, 'emb1')
        e1, e2 = emb1.float().contiguous(), emb2.float().contiguous()
        R = R.float().contiguous()
        b, d = e1.shape
        m = e2.shape[0]
        S = torch.empty((b, m), dtype=F32, device=e1.device)
        rows = torch.empty((b,), dtype=F32, device=e1.device)
        lib().tok_contrastive_fwd(b, m, d, _p(e1), _p(e2), _p(R), float(margin), _p(S), _p(rows), _st())
        ctx.save_for_backward(e1, e2, R, S)
        ctx.meta = (float(margin), emb1.dtype, emb2.dtype)
        return rows

    @staticmethod
    def backward(ctx, g):
        e1, e2, R, S = ctx.saved_tensors
        margin, t1, t2 = ctx.meta
        b, d = e1.shape
        m = e2.shape[0]
        g = g.float().contiguous()
        d1 = torch.empty_like(e1) if ctx.needs_input_grad[0] else None
        d2 = torch.empty_like(e2) if ctx.needs_input_grad[1] else None
        lib().tok_contrastive_bwd(b, m, d, _p(e1), _p(e2), _p(R), _p(S), _p(g), margin, _p(d1), _p(d2), _st())
        return (d1.to(t1) if d1 is not None else None), (d2.to(t2) if d2 is not None else None), None, None


def contrastive_rows(emb1, emb2, R, margin):
    return ContrastiveFn.apply(emb1, emb2, R, margin)


class SoftmaxXentNHWCFn(torch.autograd.Function):
    """torch.nn.CrossEntropyLoss(reduction='mean', ignore_index) on (B, C, H, W) logits with (B, H, W) targets
    (segmentation, torchok/losses/__init__.py:26): NHWC rows of C <= 64 classes, one thread per pixel."""

    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        lg = to_nhwc(logits)
        n, c, h, w = lg.shape
        cp = nhwc_pitch(lg)
        rows = n * h * w
        target = target.long().contiguous()
        if target.numel() != rows:
            raise ValueError(f'CrossEntropyLoss: target shape {tuple(target.shape)} does not match logits {tuple(lg.shape)}')
        acc = torch.zeros((2,), dtype=F32, device=lg.device)
        lib().tok_softmax_xent_small(rows, c, cp, _p(lg), _p(target), _p(acc[0:1]), _p(acc[1:2]), None, None, 1.0,
                                     None, ignore_index, _st())
        inv = 1.0 / acc[1].clamp_min(1.0)
        ctx.save_for_backward(lg, target, inv.reshape(1))
        ctx.meta = (n, c, h, w, cp, ignore_index)
        return acc[0] * inv

    @staticmethod
    def backward(ctx, g):
        lg, target, inv = ctx.saved_tensors
        n, c, h, w, cp, ignore_index = ctx.meta
        d = torch.empty((n, h, w, cp), dtype=BF16, device=lg.device)
        g = g.to(F32).contiguous()
        lib().tok_softmax_xent_small(n * h * w, c, cp, _p(lg), _p(target), None, None, _p(d), _p(inv), 1.0, _p(g),
                                     ignore_index, _st())
        d = d.permute(0, 3, 1, 2)
        return (d if cp == c else d[:, :c]), None, None


class DiceMulticlassFn(torch.autograd.Function):
    """DiceLoss(mode='multiclass', from_logits=True) (torchok/losses/segmentation/dice.py:85-188) on (B, C, H, W) logits
    with (B, H, W) int64 targets: statistics pass, one-warp finalize (loss + per-class coefficients), gradient pass."""

    @staticmethod
    def forward(ctx, logits, target, smooth, eps, log_loss):
        lg = to_nhwc(logits)
        n, c, h, w = lg.shape
        cp = nhwc_pitch(lg)
        rows = n * h * w
        target = target.long().contiguous()
        if target.numel() != rows:
            raise ValueError(f"Shapes of input {tuple(lg.shape)} and target {tuple(target.shape)} tensors don't match!")
        stats = torch.zeros((3, 64), dtype=F32, device=lg.device)
        out = torch.empty((1 + 2 * 64,), dtype=F32, device=lg.device)   # loss, then coef[2][64]
        L, st = lib(), _st()
        L.tok_dice_stats(rows, c, cp, _p(lg), _p(target), _p(stats), st)
        L.tok_dice_finalize(c, _p(stats), float(smooth), float(eps), int(bool(log_loss)), _p(out[0:1]), _p(out[1:]), st)
        ctx.save_for_backward(lg, target, out)
        ctx.meta = (n, c, h, w, cp)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        lg, target, out = ctx.saved_tensors
        n, c, h, w, cp = ctx.meta
        d = torch.empty((n, h, w, cp), dtype=BF16, device=lg.device)
        g = g.to(F32).reshape(1).contiguous()
        lib().tok_dice_bwd(n * h * w, c, cp, _p(lg), _p(target), _p(out[1:]), _p(g), _p(d), _st())
        d = d.permute(0, 3, 1, 2)
        return (d if cp == c else d[:, :c]), None, None, None, None


def dice_multiclass(logits, target, smooth=0.0, eps=1e-7, log_loss=False):
    if logits.dim() != 4 or logits.shape[1] > 64:
        raise NotImplementedError('DiceLoss: (B, C, H, W) logits with up to 64 classes')
    return DiceMulticlassFn.apply(logits, target, smooth, eps, log_loss)


def softmax_xent_nhwc(logits, target, ignore_index=-100):
    if logits.shape[1] > 64:
        raise NotImplementedError('CrossEntropyLoss on 4-D logits supports up to 64 classes')
    return SoftmaxXentNHWCFn.apply(logits, target, ignore_index)


# ------------------------------------------------------------------------------------------------------ HRNet / seg
class FuseSumFn(torch.autograd.Function):
    """y = relu(sum_t nearest_upsample(term_t)) — the fuse step of timm's HighResolutionModule.forward
    (torchok/models/backbones/hrnet.py:167-192).  terms[0] defines the output resolution; term t may be 2^s smaller."""

    @staticmethod
    def forward(ctx, relu, *terms):
        import ctypes as C2
        terms = [to_nhwc(t) for t in terms]
        n, c, h, w = terms[0].shape
        cp = nhwc_pitch(terms[0])
        shifts = []
        for t in terms:
            if nhwc_pitch(t) != cp or t.shape[1] != c:
                raise ValueError('fuse_sum: all terms must have the same channel count')
            f = h // t.shape[2]
            if f < 1 or f & (f - 1) or t.shape[2] * f != h or t.shape[3] * f != w:
                raise ValueError(f'fuse_sum: term of size {tuple(t.shape)} is not a power-of-two reduction of {h}x{w}')
            shifts.append(f.bit_length() - 1)
        out = torch.empty((n, h, w, cp), dtype=BF16, device=terms[0].device)
        bits = torch.empty((n * h * w * cp // 8,), dtype=torch.uint8, device=out.device) if relu else None
        ptrs = (C2.c_void_p * len(terms))(*[t.data_ptr() for t in terms])
        sh = (C2.c_int * len(terms))(*shifts)
        lib().tok_fuse_sum_fwd(n, h, w, cp, len(terms), ptrs, sh, int(relu), _p(out), _p(bits), _st())
        ctx.meta = (n, c, h, w, cp, shifts)
        ctx.save_for_backward(bits) if relu else None
        ctx.relu = relu
        o = out.permute(0, 3, 1, 2)
        return o if cp == c else o[:, :c]

    @staticmethod
    def backward(ctx, g):
        n, c, h, w, cp, shifts = ctx.meta
        bits = ctx.saved_tensors[0] if ctx.relu else None
        g = _dense_grad(g, cp)
        grads = []
        for i, s in enumerate(shifts):
            if not ctx.needs_input_grad[1 + i]:
                grads.append(None)
                continue
            d = torch.empty((n, h >> s, w >> s, cp), dtype=BF16, device=g.device)
            lib().tok_fuse_sum_bwd(n, h, w, cp, s, _p(g), _p(bits), _p(d), _st())
            d = d.permute(0, 3, 1, 2)
            grads.append(d if cp == c else d[:, :c])
        return (None,) + tuple(grads)


def fuse_sum(terms, relu=True):
    return FuseSumFn.apply(relu, *terms)


class BilinearCatFn(torch.autograd.Function):
    """cat([interpolate(x_k, size, 'bilinear', align_corners=False) for k], dim=1) in one padded NHWC buffer:
    segment k occupies ceil8(C_k) channels (necks/segmentation/hrnet.py:35-40).  With a single input and the same
    channel count it is plain F.interpolate (heads/segmentation/base.py:37)."""

    @staticmethod
    def forward(ctx, size, *xs):
        return _resize_cat_fwd('bilinear', ctx, size, xs)

    @staticmethod
    def backward(ctx, g):
        return _resize_cat_bwd('bilinear', ctx, g)


def _resize_cat_fwd(mode, ctx, size, xs):
    xs = [to_nhwc(x) for x in xs]
    n = xs[0].shape[0]
    ho, wo = size
    pitches = [nhwc_pitch(x) for x in xs]
    total = sum(pitches)
    out = torch.empty((n, ho, wo, total), dtype=BF16, device=xs[0].device)
    off = 0
    fwd = lib().tok_bilinear_fwd if mode == 'bilinear' else lib().tok_nearest_fwd
    for x, cp in zip(xs, pitches):
        fwd(n, x.shape[2], x.shape[3], cp, ho, wo, _p(x), _p(out), total, off, _st())
        off += cp
    ctx.meta = (n, ho, wo, total, [(x.shape[1], x.shape[2], x.shape[3], cp) for x, cp in zip(xs, pitches)])
    return out.permute(0, 3, 1, 2)

def _resize_cat_bwd(mode, ctx, g):
    n, ho, wo, total, shapes = ctx.meta
    g = _dense_grad(g, total)
    grads, off = [], 0
    bwd = lib().tok_bilinear_bwd if mode == 'bilinear' else lib().tok_nearest_bwd
    for i, (c, h, w, cp) in enumerate(shapes):
        if ctx.needs_input_grad[1 + i]:
            d = torch.empty((n, h, w, cp), dtype=BF16, device=g.device)
            bwd(n, h, w, cp, ho, wo, _p(g), total, off, _p(d), _st())
            d = d.permute(0, 3, 1, 2)
            grads.append(d if cp == c else d[:, :c])
        else:
            grads.append(None)
        off += cp
    return (None,) + tuple(grads)


def bilinear_cat(xs, size):
    """Returns the (N, sum ceil8(C_k), H, W) padded concat; the consumer conv maps its input channels with
    Conv2d.set_input_layout([C_k...])."""
    return BilinearCatFn.apply(tuple(size), *xs)


class NearestCatFn(torch.autograd.Function):
    """cat([interpolate(x_k, size, 'nearest') for k], dim=1): the upsample + skip concat of the U-Net decoder blocks
    (torchok/models/necks/segmentation/unet.py:48-56), one pass per input straight into the padded concat buffer."""

    @staticmethod
    def forward(ctx, size, *xs):
        return _resize_cat_fwd('nearest', ctx, size, xs)

    @staticmethod
    def backward(ctx, g):
        return _resize_cat_bwd('nearest', ctx, g)


def nearest_cat(xs, size):
    return NearestCatFn.apply(tuple(size), *xs)


class SpatialGatherFn(torch.autograd.Function):
    """SpatialGather_Module.forward (torchok/models/heads/segmentation/ocr.py:37-46): soft object regions.
    feats (B, C, H, W), class logits (B, K, H, W) -> context (B, C, K, 1)."""

    @staticmethod
    def forward(ctx, feats, logits):
        feats, logits = to_nhwc(feats), to_nhwc(logits)
        b, c, h, w = feats.shape
        k, kp = logits.shape[1], nhwc_pitch(logits)
        if nhwc_pitch(feats) != c:
            raise NotImplementedError('spatial gather: feature channels must be a multiple of 8')
        dev = feats.device
        stats = torch.empty((b, k, 2), dtype=F32, device=dev)
        scratch = torch.empty((b, k, c), dtype=F32, device=dev)
        out = torch.empty((b, k, 1, c), dtype=BF16, device=dev)
        lib().tok_spatial_gather_fwd(b, h * w, c, k, kp, _p(feats), _p(logits), _p(stats), _p(scratch), _p(out), _st())
        ctx.save_for_backward(feats, logits, stats, out)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        feats, logits, stats, out = ctx.saved_tensors
        b, c, h, w = feats.shape
        k, kp = logits.shape[1], nhwc_pitch(logits)
        g = _dense_grad(g, c)
        dfeats = torch.empty((b, h, w, c), dtype=BF16, device=g.device)
        dlogits = torch.empty((b, h, w, kp), dtype=BF16, device=g.device)
        lib().tok_spatial_gather_bwd(b, h * w, c, k, kp, _p(feats), _p(logits), _p(stats), _p(out), _p(g), _p(dfeats),
                                     _p(dlogits), _st())
        dl = dlogits.permute(0, 3, 1, 2)
        return dfeats.permute(0, 3, 1, 2), (dl if kp == k else dl[:, :k])


def spatial_gather(feats, logits):
    return SpatialGatherFn.apply(feats, logits)


class ObjectAttnFn(torch.autograd.Function):
    """ObjectAttentionBlock.forward's attention (ocr.py:84-95): query (B, Kc, H, W), key / value (B, Kc, K, 1)
    -> context (B, Kc, H, W) = softmax_k(Kc^-.5 q.key) . value."""

    @staticmethod
    def forward(ctx, q, key, value, scale):
        q, key, value = to_nhwc(q), to_nhwc(key), to_nhwc(value)
        b, kc, h, w = q.shape
        k = key.shape[2]
        if kc % 8 or nhwc_pitch(key) != kc or nhwc_pitch(value) != kc:
            raise NotImplementedError('object attention: key_channels must be a multiple of 8')
        out = torch.empty((b, h, w, kc), dtype=BF16, device=q.device)
        lib().tok_object_attn_fwd(b, h * w, kc, k, float(scale), _p(q), _p(key), _p(value), _p(out), _st())
        ctx.save_for_backward(q, key, value)
        ctx.scale = float(scale)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, g):
        q, key, value = ctx.saved_tensors
        b, kc, h, w = q.shape
        k = key.shape[2]
        g = _dense_grad(g, kc)
        dev = g.device
        dq = torch.empty((b, h, w, kc), dtype=BF16, device=dev)
        dk32, dv32 = torch.empty((b, k, kc), dtype=F32, device=dev), torch.empty((b, k, kc), dtype=F32, device=dev)
        dk, dv = torch.empty((b, k, 1, kc), dtype=BF16, device=dev), torch.empty((b, k, 1, kc), dtype=BF16, device=dev)
        lib().tok_object_attn_bwd(b, h * w, kc, k, ctx.scale, _p(q), _p(key), _p(value), _p(g), _p(dq), _p(dk32),
                                  _p(dv32), _p(dk), _p(dv), _st())
        return dq.permute(0, 3, 1, 2), dk.permute(0, 3, 1, 2), dv.permute(0, 3, 1, 2), None


def object_attention(q, key, value, scale):
    return ObjectAttnFn.apply(q, key, value, scale)


class ChannelScaleFn(torch.autograd.Function):
    """x * scale[n, c] (nn.Dropout2d with the caller's mask, ocr.py:126)."""

    @staticmethod
    def forward(ctx, x, scale):
        x = to_nhwc(x)
        n, c, h, w = x.shape
        cp = nhwc_pitch(x)
        s = torch.zeros((n, cp), dtype=F32, device=x.device)
        s[:, :c] = scale
        out = torch.empty((n, h, w, cp), dtype=BF16, device=x.device)
        lib().tok_channel_scale(n, h * w, cp, _p(x), _p(s), _p(out), _st())
        ctx.save_for_backward(s)
        ctx.c = c
        o = out.permute(0, 3, 1, 2)
        return o if cp == c else o[:, :c]

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        n, cp = s.shape
        g = _dense_grad(g, cp)
        _, c, h, w = g.shape
        out = torch.empty((n, h, w, cp), dtype=BF16, device=g.device)
        lib().tok_channel_scale(n, h * w, cp, _p(g), _p(s), _p(out), _st())
        o = out.permute(0, 3, 1, 2)
        return (o if cp == ctx.c else o[:, :ctx.c]), None


def dropout2d(x, p, training):
    """nn.Dropout2d: whole channels of a sample are zeroed with probability p and the rest scaled by 1 / (1 - p)."""
    if not training or p == 0.0:
        return x
    keep = (torch.rand((x.shape[0], x.shape[1]), device=x.device) >= p).float() / (1.0 - p)
    return ChannelScaleFn.apply(x, keep)


def bilinear_resize(x, size):
    c = x.shape[1]
    out = BilinearCatFn.apply(tuple(size), x)
    return out if out.shape[1] == c else out[:, :c]


# ------------------------------------------------------------------------------------------------------ Swin-V2
class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last dim of a (rows, C) bf16 matrix; with `residual` computes
    residual + rowscale[sample] * LN(x) — the res-post-norm block tail `x + drop_path(norm(f(x)))` of timm's
    SwinTransformerBlock.  weight / bias gradients are accumulated into `.grad` directly."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, residual, rowscale, rows_per_sample, colsum_param=None, res_link=None):
        require_cuda(x, 'LayerNorm input')
        ctx.colsum_param = colsum_param
        ctx.res_link = res_link if residual is not None else None
        x = x.to(BF16).contiguous()
        rows, c = x.shape
        out = torch.empty_like(x)
        stats = torch.empty((2, rows), dtype=F32, device=x.device)
        res = residual.to(BF16).contiguous() if residual is not None else None
        lib().tok_layernorm_fwd(rows, c, _p(x), _p(weight), _p(bias), float(eps), _p(res), _p(rowscale),
                                int(rows_per_sample), _p(out), _p(stats[0]), _p(stats[1]), _st())
        ctx.save_for_backward(x, stats, rowscale)
        ctx.params = (weight, bias)
        ctx.meta = (rows, c, int(rows_per_sample), residual is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        x, stats, rowscale = ctx.saved_tensors
        weight, bias = ctx.params
        rows, c, rps, has_res = ctx.meta
        g = g.to(BF16).contiguous()
        dx = torch.empty_like(x)
        gw = grad_buffer(weight) if weight.requires_grad else None
        gb = grad_buffer(bias) if bias.requires_grad else None
        cp = ctx.colsum_param
        gx = grad_buffer(cp) if cp is not None and cp.requires_grad else None
        lib().tok_layernorm_bwd(rows, c, _p(x), _p(weight), _p(stats[0]), _p(stats[1]), _p(g), _p(rowscale), rps,
                                _p(dx), _p(gw), _p(gb), _p(gx), _st())
        grad_ready(weight)
        grad_ready(bias)
        gres = g if has_res else None
        if ctx.res_link is not None:   # the branch's first linear layer adds the skip gradient in its dgrad epilogue
            ctx.res_link.g, gres = g, None
        return dx, None, None, None, gres, None, None, None, None


def layernorm(x, weight, bias, eps=1e-5, residual=None, rowscale=None, rows_per_sample=1, colsum_param=None,
              res_link=None):
    """`colsum_param`: fp32 (C,) bias of the linear layer that produced `x`; its gradient (column sums of dx) is
    accumulated by the LayerNorm backward (only if `layernorm_fuses_colsum(C)`)."""
    return LayerNormFn.apply(x, weight, bias, eps, residual, rowscale, rows_per_sample, colsum_param, res_link)


def layernorm_fuses_colsum(c):
    return bool(lib().tok_layernorm_has_dxsum(int(c)))


def gelu_fuses_colsum(c):
    return c % 128 == 0


class GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, colsum_param=None):
        x = x.to(BF16).contiguous()
        y = torch.empty_like(x)
        lib().tok_gelu_fwd(x.numel(), _p(x), _p(y), _st())
        ctx.save_for_backward(x)
        ctx.colsum_param = colsum_param
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.to(BF16).contiguous()
        dx = torch.empty_like(x)
        cp = ctx.colsum_param
        gx = grad_buffer(cp) if cp is not None and cp.requires_grad else None
        lib().tok_gelu_bwd(x.numel(), x.shape[-1], _p(x), _p(g), _p(dx), _p(gx), _st())
        return dx, None


def gelu(x, colsum_param=None):
    """`colsum_param`: bias of the linear layer that produced `x` (timm Mlp.fc1); see `layernorm`."""
    return GeluFn.apply(x, colsum_param)


# timm Mlp tail `fc2(act(h))` as ONE autograd node: the forward is the GELU pass + the fc2 GEMM as before; the backward
# applies gelu'(h) in the epilogue of fc2's data-gradient GEMM (tok_linear_dgrad_gelu) instead of writing the gradient of
# the activation, reading it back together with h and writing the product (three passes over the widest tensor of the
# block).  MEASURED SLOWER (r5, Swin-T bs256, same box): the 12 fused launches take 4.88 ms against 1.29 ms (plain dgrad)
# + 1.97 ms (tok_gelu_bwd) for the launches they replace, step 26.6 vs 25.2 ms — the derivative costs ~14 instructions
# and 2 MUFU per element, and in the GEMM epilogue only the 8 epilogue warps of the CTA (2 per scheduler) execute it,
# while the stand-alone pass spreads the same work over 64 warps per SM and runs at 4.2 TB/s.  So the default is the
# two-node form; TOK_GELU_DGRAD=1 selects the fused node (kept correct by tests/test_swin_gpu.py).  Re-measured after the
# derivative went from ~33 to ~14 instructions per element (rcp.approx / ex2.approx): 2.92 ms fused against 1.14 + 1.53 ms,
# step 24.08 vs 24.05 ms — even, so the default stays.
_GELU_DGRAD = os.environ.get('TOK_GELU_DGRAD', '0') == '1'


class GeluLinearFn(torch.autograd.Function):
    """y = Linear(GELU(h)): h (M, K) bf16 = the fc1 output, weight (N, K) fp32 master, bias (N,) fp32.
    `colsum_param`: fc1's bias — its gradient is the column sum of dh, accumulated by the fused kernel.
    `bias_grad_external`: as in LinearFn (the LayerNorm behind fc2 accumulates fc2's bias gradient)."""

    @staticmethod
    def forward(ctx, h, weight, bias, colsum_param, bias_grad_external):
        require_cuda(h, 'gelu-linear input')
        m, k = h.shape
        n = weight.shape[0]
        h = h.to(BF16).contiguous()
        a = torch.empty_like(h)
        L = lib()
        st = _st()
        L.tok_gelu_fwd(h.numel(), _p(h), _p(a), st)
        w = shadow_of(weight)
        b = bias.detach().float() if bias is not None else None
        y = torch.empty((m, n), dtype=BF16, device=h.device)
        L.tok_linear_fwd(m, n, k, _p(a), _p(w), _p(b), _p(y), st)
        ctx.save_for_backward(h, a, w)
        ctx.params = (weight, bias, colsum_param)
        ctx.bias_grad_external = bool(bias_grad_external)
        ctx.dims = (m, n, k)
        return y

    @staticmethod
    def backward(ctx, g):
        h, a, w = ctx.saved_tensors
        weight, bias, cp = ctx.params
        m, n, k = ctx.dims
        L = lib()
        st = _st()
        g = g.to(BF16).contiguous()
        dh = None
        if ctx.needs_input_grad[0]:
            dh = torch.empty((m, k), dtype=BF16, device=g.device)
            gx = grad_buffer(cp) if cp is not None and cp.requires_grad else None
            L.tok_linear_dgrad_gelu(m, n, k, _p(g), _p(w), _p(h), _p(dh), _p(gx), st)
        if weight.requires_grad:
            gw = grad_buffer(weight)
            wgrad_async((a, g, gw), lambda s: L.tok_linear_wgrad(m, n, k, _p(a), _p(g), _p(gw), s))
        if bias is not None and bias.requires_grad and not ctx.bias_grad_external:
            acc = torch.zeros((2, n), dtype=F32, device=g.device)
            L.tok_bn_bwd_reduce(m, n, _p(g), None, None, _p(g), _p(acc[0]), _p(acc[1]), st)
            grad_buffer(bias).add_(acc[0])
        grad_ready(weight)
        if bias is not None:
            grad_ready(bias)
        return dh, None, None, None, None


def gelu_dgrad_fused(hidden, out_features):
    """True when `gelu_linear` runs as the fused node (then the bias gradient of the layer in front of the GELU always
    comes out of its backward, whatever the hidden width)."""
    return _GELU_DGRAD and hidden % 8 == 0 and out_features % 8 == 0


def gelu_linear(h, weight, bias, colsum_param=None, bias_grad_external=False):
    """fc2(GELU(h)) with the fused backward when the shapes allow it (both widths multiples of 8)."""
    n, k = weight.shape
    if gelu_dgrad_fused(k, n) and h.dim() == 2:
        return GeluLinearFn.apply(h, weight, bias, colsum_param, bias_grad_external)
    return linear(gelu(h, colsum_param), weight, bias, bias_grad_external)


class WindowAttnFn(torch.autograd.Function):
    """timm WindowAttention core on the (B*H*W, 3C) qkv matrix: cosine attention + logit scale + bias + shift mask ->
    (B*H*W, C).  `bias` (heads, N, N) fp32 is an autograd input (its gradient flows back to the cpb MLP);
    `logit_scale` gets its gradient accumulated into `.grad`."""

    @staticmethod
    def forward(ctx, qkv, bias, logit_scale, geom, qv_bias=None):
        b, h, w, c, heads, ws, shift = geom
        ctx.qv_bias = qv_bias
        qkv = qkv.to(BF16).contiguous()
        bias = bias.float().contiguous()
        ls = logit_scale.detach().float().reshape(-1).contiguous()
        out = torch.empty((b * h * w, c), dtype=BF16, device=qkv.device)
        lib().tok_window_attn_fwd(b, h, w, c, heads, ws, shift, _p(qkv), _p(ls), _p(bias), _p(out), _st())
        ctx.save_for_backward(qkv, bias, ls)
        ctx.meta = (geom, logit_scale)
        return out

    @staticmethod
    def backward(ctx, g):
        qkv, bias, ls = ctx.saved_tensors
        (b, h, w, c, heads, ws, shift), logit_scale = ctx.meta
        g = g.to(BF16).contiguous()
        dqkv = torch.empty_like(qkv)
        dbias = torch.zeros_like(bias)
        dls = torch.zeros_like(ls)
        col = torch.zeros(3 * c, dtype=F32, device=qkv.device) if ctx.qv_bias is not None else None
        lib().tok_window_attn_bwd(b, h, w, c, heads, ws, shift, _p(qkv), _p(ls), _p(bias), _p(g), _p(dqkv), _p(dbias),
                                  _p(dls), _p(col), _st())
        if logit_scale.requires_grad:
            grad_buffer(logit_scale).add_(dls.reshape(logit_scale.shape))
            grad_ready(logit_scale)
        if col is not None:   # q_bias / v_bias gradients = column sums of dq / dv, accumulated by the kernel
            q_bias, v_bias = ctx.qv_bias
            grad_buffer(q_bias).add_(col[:c])
            grad_buffer(v_bias).add_(col[2 * c:])
        return dqkv, dbias, None, None, None


def attn_fuses_qv_bias_grad():
    return os.environ.get('TOK_ATTN_BWD_CUDA_CORES') != '1'


def window_attention(qkv, bias, logit_scale, geom, qv_bias=None):
    """`qv_bias=(q_bias, v_bias)`: their gradients are produced by the attention backward (see `linear`)."""
    return WindowAttnFn.apply(qkv, bias, logit_scale, geom, qv_bias)


class CpbBiasFn(torch.autograd.Function):
    """timm WindowAttention's bias table: 16 * sigmoid(cpb_mlp(relative_coords_table))[relative_position_index] ->
    (heads, N, N) fp32.  Parameters: cpb_mlp.0.weight (512, 2), cpb_mlp.0.bias (512), cpb_mlp.2.weight (heads, 512);
    their gradients are accumulated into `.grad` by the backward kernels."""

    @staticmethod
    def forward(ctx, coords, w1, b1, w2, ws):
        heads, hid = w2.shape
        t = (2 * ws - 1) ** 2
        n = ws * ws
        dev = w1.device
        c = coords.reshape(t, 2).float().contiguous()
        w1c, b1c, w2c = w1.detach().float().contiguous(), b1.detach().float().contiguous(), w2.detach().float().contiguous()
        hidden = torch.empty((t, hid), dtype=F32, device=dev)
        table = torch.empty((t, heads), dtype=F32, device=dev)
        bias = torch.empty((heads, n, n), dtype=F32, device=dev)
        lib().tok_cpb_bias_fwd(ws, heads, hid, _p(c), _p(w1c), _p(b1c), _p(w2c), _p(hidden), _p(table), _p(bias), _st())
        ctx.save_for_backward(c, w2c, hidden, table)
        ctx.params = (w1, b1, w2)
        ctx.ws = ws
        return bias

    @staticmethod
    def backward(ctx, g):
        c, w2c, hidden, table = ctx.saved_tensors
        w1, b1, w2 = ctx.params
        heads, hid = w2c.shape
        g = g.float().contiguous()
        dtable = torch.empty_like(table)
        lib().tok_cpb_bias_bwd(ctx.ws, heads, hid, _p(c), _p(w2c), _p(hidden), _p(table), _p(g), _p(dtable),
                               _p(grad_buffer(w1)), _p(grad_buffer(b1)), _p(grad_buffer(w2)), _st())
        for prm in (w1, b1, w2):
            grad_ready(prm)
        return None, None, None, None, None


def cpb_bias(coords, w1, b1, w2, ws):
    return CpbBiasFn.apply(coords, w1, b1, w2, ws)


class QkvLinearFn(torch.autograd.Function):
    """F.linear(x, qkv.weight, cat(q_bias, zeros, v_bias)) of timm's WindowAttention (k has no bias)."""

    @staticmethod
    def forward(ctx, x, weight, q_bias, v_bias, bias_grad_external=False, res_link=None):
        m, k = x.shape
        n = weight.shape[0]
        ctx.bias_grad_external = bool(bias_grad_external)
        ctx.res_link = res_link
        x = x.to(BF16).contiguous()
        w = shadow_of(weight)
        b = None
        if q_bias is not None:
            b = torch.cat([q_bias.detach(), torch.zeros_like(v_bias), v_bias.detach()]).float().contiguous()
        y = torch.empty((m, n), dtype=BF16, device=x.device)
        lib().tok_linear_fwd(m, n, k, _p(x), _p(w), _p(b), _p(y), _st())
        ctx.save_for_backward(x, w)
        ctx.params = (weight, q_bias, v_bias)
        return y

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        weight, q_bias, v_bias = ctx.params
        m, k = x.shape
        n = weight.shape[0]
        g = g.to(BF16).contiguous()
        L, st = lib(), _st()
        dx = torch.empty((m, k), dtype=BF16, device=g.device)
        _dgrad_with_link(L, ctx.res_link, m, n, k, g, w, dx, st)
        if weight.requires_grad:
            L.tok_linear_wgrad(m, n, k, _p(x), _p(g), _p(grad_buffer(weight)), st)
        if q_bias is not None:
            if not ctx.bias_grad_external:
                acc = torch.zeros((2, n), dtype=F32, device=g.device)
                L.tok_bn_bwd_reduce(m, n, _p(g), None, None, _p(g), _p(acc[0]), _p(acc[1]), st)
                c = n // 3
                grad_buffer(q_bias).add_(acc[0, :c])
                grad_buffer(v_bias).add_(acc[0, 2 * c:])
            grad_ready(q_bias)
            grad_ready(v_bias)
        grad_ready(weight)
        return dx, None, None, None, None, None


def qkv_linear(x, weight, q_bias, v_bias, bias_grad_external=False, res_link=None):
    return QkvLinearFn.apply(x, weight, q_bias, v_bias, bias_grad_external, res_link)


class PatchMergeFn(torch.autograd.Function):
    """The gather of timm's PatchMerging: (B*H*W, C) tokens -> (B*H/2*W/2, 4C); a permutation in both directions."""

    @staticmethod
    def forward(ctx, x, b, h, w):
        x = x.to(BF16).contiguous()
        c = x.shape[-1]
        out = torch.empty((b * (h // 2) * (w // 2), 4 * c), dtype=BF16, device=x.device)
        lib().tok_patch_merge(b, h, w, c, _p(x), _p(out), 0, _st())
        ctx.geom = (b, h, w, c)
        return out

    @staticmethod
    def backward(ctx, g):
        b, h, w, c = ctx.geom
        g = g.to(BF16).contiguous()
        dx = torch.empty((b * h * w, c), dtype=BF16, device=g.device)
        lib().tok_patch_merge(b, h, w, c, _p(g), _p(dx), 1, _st())
        return dx, None, None, None


def patch_merge(x, b, h, w):
    return PatchMergeFn.apply(x, b, h, w)


_PATCH_GEMM = os.environ.get('TOK_PATCH_EMBED_GEMM', '1') != '0'   # 0: the r2 CUDA-core patch-embedding kernels


class PatchEmbedFn(torch.autograd.Function):
    """timm PatchEmbed.proj (Conv2d(3, E, 4, stride 4)) + flatten(2).transpose(1, 2) on the NCHW image: (B*H/4*W/4, E)
    bf16 tokens.  r3: the patches are laid out once as a bf16 [tokens][48] matrix in the weight's memory order
    (tok_patchify) and the projection / its weight gradient run as tensor-core GEMMs (tok_linear_fwd / tok_linear_wgrad on
    the [E][4][4][3] weight seen as [E][48]); the bias gradient is the column sum of the incoming gradient.  The image gets
    no gradient.  TOK_PATCH_EMBED_GEMM=0 keeps the CUDA-core kernels that read the fp32 image directly."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        require_cuda(x, 'image')
        b, c, h, w = x.shape
        e = weight.shape[0]
        gemm = _PATCH_GEMM and c == 3 and e % 8 == 0 and is_krsc(weight) and weight.shape[2] == 4 and weight.shape[3] == 4
        ctx.gemm = gemm
        ctx.params = (weight, bias)
        ctx.geom = (b, h, w, e)
        if gemm:
            x = x.detach()
            if x.dtype not in (F32, BF16):
                x = x.float()
            x = x.contiguous()
            m = b * (h // 4) * (w // 4)
            patches = torch.empty((m, 48), dtype=BF16, device=x.device)
            lib().tok_patchify(b, c, h, w, 4, int(x.dtype == BF16), _p(x), _p(patches), _st())
            out = torch.empty((m, e), dtype=BF16, device=x.device)
            bvec = bias.detach().float() if bias is not None else None
            lib().tok_linear_fwd(m, e, 48, _p(patches), _p(shadow_of(weight)), _p(bvec), _p(out), _st())
            ctx.save_for_backward(patches)
            return out
        x = x.detach().float().contiguous()
        wst = (C.c_int * 4)(*weight.stride())
        out = torch.empty((b * (h // 4) * (w // 4), e), dtype=BF16, device=x.device)
        lib().tok_patch_embed_fwd(b, h, w, e, _p(x), _p(weight), _p(bias), wst, _p(out), _st())
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        weight, bias = ctx.params
        b, h, w, e = ctx.geom
        g = g.to(BF16).contiguous()
        gw = grad_buffer(weight)
        if gw.stride() != weight.stride():
            raise RuntimeError('patch_embed: gradient buffer and weight must share their memory layout')
        gb = grad_buffer(bias) if bias is not None and bias.requires_grad else None
        if ctx.gemm:
            m = x.shape[0]
            L, st = lib(), _st()
            if weight.requires_grad:
                wgrad_async((x, g, gw), lambda s: L.tok_linear_wgrad(m, e, 48, _p(x), _p(g), _p(gw), s))
            if gb is not None:
                acc = torch.zeros((2, e), dtype=F32, device=g.device)
                L.tok_bn_bwd_reduce(m, e, _p(g), None, None, _p(g), _p(acc[0]), _p(acc[1]), st)
                gb.add_(acc[0])
        else:
            wst = (C.c_int * 4)(*weight.stride())
            lib().tok_patch_embed_bwd(b, h, w, e, _p(x), _p(g), wst, _p(gw), _p(gb), _st())
        grad_ready(weight)
        if bias is not None:
            grad_ready(bias)
        return None, None, None


def patch_embed_supported(cin, patch, h, w, e):
    return bool(lib().tok_patch_embed_supported(int(cin), int(patch), int(h), int(w), int(e)))


def patch_embed(x, weight, bias):
    # weight / bias are passed as non-differentiable handles: their gradients go straight into the arena buffers
    return PatchEmbedFn.apply(x, weight, bias)
