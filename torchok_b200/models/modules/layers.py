"""Parameter-holding layers and the fused autograd Functions built on libtokb200.

`Conv2d` / `BatchNorm2d` subclass the torch modules only to keep parameter names, shapes, init code and
`isinstance` checks of reference-style code working (state-dict contract); their arithmetic never goes
through torch.nn.functional.  The unit of execution is conv -> BN -> (+shortcut) -> (ReLU)
(torchok/models/modules/bricks/convbnact.py:48-53; timm BasicBlock/Bottleneck as built by
torchok/models/backbones/resnet.py:363-405), run as a sequence of C-ABI calls by one autograd Function per unit, per
residual block, or per stem.
"""
import ctypes as C

import torch
import torch.nn as nn

from ... import kernels as K
from ..._lib import lib

BF16, F32 = K.BF16, K.F32
# TOK_DIRECT_HALO=0: A/B aid — padded temporaries around every conv of a channel count that is not a multiple of 8 (r1 path)
_DIRECT_HALO = __import__('os').environ.get('TOK_DIRECT_HALO', '1') != '0'


def _single(v, what):
    if isinstance(v, (tuple, list)):
        if len(set(v)) != 1:
            raise NotImplementedError(f'{what} must be the same along H and W (got {v})')
        v = v[0]
    return int(v)


class Conv2d(nn.Conv2d):
    """nn.Conv2d parameters (weight OIHW fp32, optional bias) stored in [K][R][S][C] memory; computed by
    tok_conv_fprop / tok_conv_dgrad / tok_conv_wgrad."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 padding_mode='zeros', device=None, dtype=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         padding_mode, device, dtype)
        if groups != 1:
            raise NotImplementedError('torchok_b200.Conv2d: grouped convolution is out of scope (SURVEY §2)')
        if padding_mode != 'zeros' or isinstance(padding, str):
            raise NotImplementedError('torchok_b200.Conv2d: only explicit zero padding')
        self.tok_stride = _single(self.stride, 'stride')
        self.tok_pad = _single(self.padding, 'padding')
        self.tok_dil = _single(self.dilation, 'dilation')
        self.cin_p, self.cout_p = K.ceil8(in_channels), K.ceil8(out_channels)
        self.weight.data = self.weight.data.contiguous(memory_format=torch.channels_last)
        self._descs = {}
        self.tok_in_map = None  # optional LongTensor: padded position of every input channel (see set_input_layout)

    def set_input_layout(self, segments):
        """The input arrives as a concatenation of channel segments, each padded to a multiple of 8 (the HRNet
        segmentation neck's torch.cat, necks/segmentation/hrnet.py:40): map logical input channels to padded slots."""
        pos, off = [], 0
        for c in segments:
            pos.extend(range(off, off + c))
            off += K.ceil8(c)
        assert len(pos) == self.in_channels
        self.tok_in_map = torch.tensor(pos, dtype=torch.long)
        self.cin_p = off
        self._descs = {}

    def desc(self, x, allow_direct=True):
        """tokConvDesc for this input.  A padded channel count (18 -> 24) whose three operations all run on the halo
        3x3 kernels gets the DIRECT form: `wk` / `wc` name the real weight dimensions, so the kernels read the arena's
        unpadded bf16 shadow and accumulate into the parameter's own fp32 gradient — no padded temporaries, no torch
        copy kernels around the launch (they were ~26 tiny launches per unit and step in HRNet)."""
        n, _, h, w = x.shape
        key = (n, h, w, allow_direct)
        hit = self._descs.get(key)
        if hit is None:
            r, s = self.kernel_size
            d, p, q = K.conv_desc(n, h, w, self.cin_p, self.cout_p, r, s, self.tok_stride, self.tok_pad, self.tok_dil)
            if allow_direct and self.padded and self.tok_in_map is None and _DIRECT_HALO and \
                    lib().tok_conv_halo_caps(C.byref(d)) == 7:
                d, p, q = K.conv_desc(n, h, w, self.cin_p, self.cout_p, r, s, self.tok_stride, self.tok_pad, self.tok_dil,
                                      self.out_channels, self.in_channels)
            hit = self._descs[key] = (d, (p, q))
        return hit

    @staticmethod
    def is_direct(d):
        return d is not None and (d.wk != 0 or d.wc != 0)

    @property
    def padded(self):
        return self.cin_p != self.in_channels or self.cout_p != self.out_channels or self.tok_in_map is not None

    def _in_index(self, device):
        if self.tok_in_map is None:
            return slice(0, self.in_channels)
        if self.tok_in_map.device != device:
            self.tok_in_map = self.tok_in_map.to(device)
        return self.tok_in_map

    def shadow(self, d=None):
        """bf16 [Kp][R][S][Cp] weights for the kernels ([K][R][S][C] unpadded for a direct descriptor)."""
        w = self.weight
        if not K.is_krsc(w):
            w.data = w.data.contiguous(memory_format=torch.channels_last)
            if not K.is_krsc(w):  # 1x1 / single-channel corner cases of torch's stride normalisation
                w.data = w.data.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        if not self.padded or self.is_direct(d):
            return K.shadow_of(w)
        r, s = self.kernel_size
        if not K.is_krsc(w) or w.dtype != F32:
            full = torch.zeros((self.cout_p, r, s, self.cin_p), dtype=BF16, device=w.device)
            tmp = K.cast_bf16(w.detach())
            full[:self.out_channels, :, :, self._in_index(w.device)] = tmp.permute(0, 2, 3, 1)
            return full
        # one launch: zero pad + cast + (optional) channel scatter
        full = torch.empty((self.cout_p, r, s, self.cin_p), dtype=BF16, device=w.device)
        maps = self._maps(w.device)
        lib().tok_pad_weight(self.out_channels, r * s, self.in_channels, self.cout_p, self.cin_p, K._p(w.detach()),
                             K._p(maps[1]) if maps is not None else None, K._p(full), K._st())
        return full

    def _maps(self, device):
        """(padded position of every input channel, source channel of every padded position or -1) as int32 tensors, or
        None when the input channels are simply the first `in_channels` lanes."""
        if self.tok_in_map is None:
            return None
        hit = getattr(self, '_tok_maps', None)
        if hit is None or hit[0].device != device:
            fwd = self.tok_in_map.to(device=device, dtype=torch.int32)
            inv = torch.full((self.cin_p,), -1, dtype=torch.int32, device=device)
            inv[fwd.long()] = torch.arange(self.in_channels, dtype=torch.int32, device=device)
            hit = self._tok_maps = (fwd.contiguous(), inv.contiguous())
        return hit

    def wgrad_target(self, d=None):
        """(buffer the wgrad kernel accumulates into, finish()) — the parameter's own fp32 .grad when unpadded (or when the
        descriptor is direct)."""
        if not self.weight.requires_grad:
            return None, None
        g = K.grad_buffer(self.weight)
        if (not self.padded or self.is_direct(d)) and K.is_krsc(g):
            return g, None
        r, s = self.kernel_size
        tmp = torch.zeros((self.cout_p, r, s, self.cin_p), dtype=F32, device=g.device)

        def finish():
            if K.is_krsc(g) and g.dtype == F32:   # one launch: gather the real lanes of the padded gradient and add
                maps = self._maps(g.device)
                lib().tok_unpad_wgrad_add(self.out_channels, r * s, self.in_channels, self.cin_p, K._p(tmp),
                                          K._p(maps[0]) if maps is not None else None, K._p(g), K._st())
            else:
                g.add_(tmp[:self.out_channels, :, :, self._in_index(g.device)].permute(0, 3, 1, 2))
        return tmp, finish

    def forward(self, x):
        return ConvFn.apply(x, self, False, self.weight, self.bias)


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d parameters/buffers; the statistics, normalisation and backward run in tok_bn_* kernels fused
    around the producing convolution."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True, device=None,
                 dtype=None):
        super().__init__(num_features, eps, momentum, affine, track_running_stats, device, dtype)
        if not (affine and track_running_stats) or momentum is None:
            raise NotImplementedError('torchok_b200.BatchNorm2d: affine=True, track_running_stats=True, momentum set')
        self.cp = K.ceil8(num_features)
        # fwd sum / sqsum, bwd sum_g / sum_gy: zero between uses (the finalize kernels hand them back zeroed); row 4
        # holds the 32-bit ticket counters of the fused reduce+finalize launches (word 0 forward, word 1 backward)
        self.register_buffer('_tok_acc', torch.zeros(5, self.cp), persistent=False)
        self._pending_batches = 0
        self._register_state_dict_hook(_flush_batches)

    def state(self, count_batch=True):
        """BNState for the fused unit.  When the channel count is not a multiple of 8 (HRNet's 18 / 36-channel
        branches) the activations carry zero pad lanes up to `cp`; weight / bias / running statistics keep their real
        size `cv` and the *_cv finalize kernels give the pad lanes scale = shift = 0."""
        if self._tok_acc.dtype != F32:
            self._tok_acc = self._tok_acc.float()
        # `track_running_stats` switched off after construction (FreezeUnfreeze's `bn_track_running_stats: false`,
        # torchok/callbacks/freeze_unfreeze.py:113-117): torch then normalises with batch statistics in training mode
        # and leaves the running buffers alone — a momentum of zero in the fused finalize step does exactly that.
        momentum = self.momentum if self.track_running_stats else 0.0
        if self.training and count_batch and self.track_running_stats:
            self._pending_batches += 1
        return K.BNState(self.weight, self.bias, self.running_mean, self.running_var, self.eps, momentum,
                         self.training, self._tok_acc, self.cp, self.num_features)

    def commit(self):
        """(r1 copied running statistics back from padded temporaries; the *_cv finalize kernels now work on the real
        buffers, so there is nothing to do.)"""

    def forward(self, x):
        raise NotImplementedError('torchok_b200.BatchNorm2d is executed fused with its producer conv '
                                  '(use conv_bn_act / ConvBnAct), not stand-alone')


def _flush_batches(module, state_dict, prefix, local_metadata):
    if module._pending_batches:
        module.num_batches_tracked += module._pending_batches
        module._pending_batches = 0
        state_dict[prefix + 'num_batches_tracked'] = module.num_batches_tracked
    return state_dict


def _bn_grads(bn):
    gw = K.grad_buffer(bn.weight) if bn.weight.requires_grad else None
    gb = K.grad_buffer(bn.bias) if bn.bias.requires_grad else None
    return gw, gb


def _unit_fwd(x, conv, bn, relu, residual, keep):
    d, pq = conv.desc(x)
    w = conv.shadow(d)
    conv._tok_last_shadow = w if (conv.padded and not conv.is_direct(d)) else None   # the backward of this step re-uses it
    res = K.unit_forward(x, d, pq, w, bn.state(), relu, residual, keep)
    return res, d


def _unit_bwd(saved, d, conv, bn, dout, **kw):
    wbuf, finish = conv.wgrad_target(d)
    gw, gb = _bn_grads(bn)
    st = K.BNState(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, bn.training,
                   bn._tok_acc, bn.cp, bn.num_features)
    w = getattr(conv, '_tok_last_shadow', None)
    if w is None or conv.is_direct(d):
        w = conv.shadow(d)
    out = K.unit_backward(saved, d, w, st, dout, wgrad_into=wbuf, dgamma=gw, dbeta=gb,
                          wgrad_direct=finish is None, **kw)
    if finish:
        finish()
    K.grad_ready(conv.weight)
    K.grad_ready(bn.weight)
    K.grad_ready(bn.bias)
    return out


def _params(*mods):
    ps = []
    for m in mods:
        if m is not None:
            ps.extend(p for p in m.parameters(recurse=False))
    return ps


# ------------------------------------------------------------------------------------------------------ single unit
class ConvBnActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, conv, bn, relu, keep, *params):
        x = K.to_nhwc(x)
        if residual is not None:
            residual = K.to_nhwc(residual)
        (out, saved), d = _unit_fwd(x, conv, bn, relu, residual, keep)
        if keep:
            ctx.mods = (conv, bn, d, residual is not None)
            ctx.has_bits, ctx.mode = saved[2] is not None, saved[4]
            ctx.save_for_backward(*[t for t in saved[:4] if t is not None])
        return out if conv.cout_p == conv.out_channels else out[:, :conv.out_channels]

    @staticmethod
    def backward(ctx, dout):
        conv, bn, d, has_res = ctx.mods
        t = list(ctx.saved_tensors)
        saved = (t[0], t[1], t[2] if ctx.has_bits else None, t[-1], ctx.mode)
        dout = K._dense_grad(dout, d.k)
        dx, dres = _unit_bwd(saved, d, conv, bn, dout, need_dx=ctx.needs_input_grad[0], want_dres=has_res)
        if dx is not None and conv.cin_p != conv.in_channels and conv.tok_in_map is None:
            dx = dx[:, :conv.in_channels]
        if dres is not None and conv.cout_p != conv.out_channels:
            dres = dres[:, :conv.out_channels]
        return (dx, dres, None, None, None, None) + (None,) * (len(ctx.needs_input_grad) - 6)


def conv_bn_act(x, conv, bn, relu=True, residual=None):
    params = _params(conv, bn)
    keep = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
    return ConvBnActFn.apply(x, residual, conv, bn, relu, keep, *params)


class ConvFn(torch.autograd.Function):
    """Plain convolution (+bias) (+ReLU) without normalisation: heads and FPN-style convs."""

    @staticmethod
    def forward(ctx, x, conv, relu, weight, bias):
        x = K.to_nhwc(x)
        d, (p, q) = conv.desc(x, allow_direct=False)   # bias / ReLU epilogues live in the generic kernel
        w = conv.shadow()
        b = bias
        if bias is not None:
            b = bias.detach().float()
            if conv.cout_p != conv.out_channels:
                b = torch.zeros(conv.cout_p, dtype=F32, device=x.device)
                b[:conv.out_channels] = bias.detach()
        y = torch.empty((d.n, p, q, d.k), dtype=BF16, device=x.device)
        K.conv_fprop(d, x, w, y, bias=b, relu=relu)
        ctx.conv, ctx.d, ctx.relu = conv, d, relu
        ctx.save_for_backward(x, y if relu else None)   # y: post-ReLU output (its sign pattern is the mask)
        out = y.permute(0, 3, 1, 2)
        return out if conv.cout_p == conv.out_channels else out[:, :conv.out_channels]

    @staticmethod
    def backward(ctx, dout):
        conv, d = ctx.conv, ctx.d
        x, y = ctx.saved_tensors
        dy = K._dense_grad(dout, d.k)
        if ctx.relu:
            # conv -> ReLU without normalisation (ConvBnAct(use_batchnorm=False), unet.py:37-38): dy = dout * (y > 0) through
            # the BatchNorm backward-apply kernel with identity coefficients (mask rebuilt from y: scale 1, shift 0)
            rows = dy.shape[0] * dy.shape[2] * dy.shape[3]   # pixels (dy may be a [:, :C] view of a padded buffer)
            one = torch.ones(d.k, dtype=F32, device=dy.device)
            zero = torch.zeros(d.k, dtype=F32, device=dy.device)
            masked = torch.empty_like(dy)
            lib().tok_bn_bwd_apply2(rows, d.k, K._p(dy), None, K._p(y), K.MASK_Y, None, K._p(one), K._p(zero), K._p(one),
                                    K._p(zero), K._p(zero), K._p(masked), None, K._st())
            dy = masked
        dx = None
        if ctx.needs_input_grad[0]:
            dxb = torch.empty((d.n, d.h, d.w, d.c), dtype=BF16, device=dy.device)
            K.conv_dgrad(d, dy, conv.shadow(), dxb)
            dx = dxb.permute(0, 3, 1, 2)
            if conv.cin_p != conv.in_channels:
                dx = dx[:, :conv.in_channels]
        wbuf, finish = conv.wgrad_target()
        if wbuf is not None:
            K.conv_wgrad(d, x, dy, wbuf)
            if finish:
                finish()
        if conv.bias is not None and conv.bias.requires_grad:
            rows = dy.shape[0] * dy.shape[2] * dy.shape[3]   # pixels (dy may be a [:, :C] view of a padded buffer)
            acc = torch.zeros((2, d.k), dtype=F32, device=dy.device)
            lib().tok_bn_bwd_reduce(rows, d.k, K._p(dy), None, None, K._p(dy), K._p(acc[0]), K._p(acc[1]), K._st())
            K.grad_buffer(conv.bias).add_(acc[0, :conv.out_channels])
        K.grad_ready(conv.weight)
        if conv.bias is not None:
            K.grad_ready(conv.bias)
        return dx, None, None, None, None


# ------------------------------------------------------------------------------------------------------ residual block
class ResidualBlockFn(torch.autograd.Function):
    """A whole timm BasicBlock / Bottleneck (torchok/models/backbones/resnet.py:14, built at :393-398):
    units[0..n-2] are conv+BN+ReLU, units[n-1] is conv+BN, then `+ shortcut` and ReLU; shortcut = x or
    downsample conv+BN.  One Function per block so that the two gradient paths into x are merged by the dgrad
    epilogue (`addend`) instead of a separate add pass."""

    @staticmethod
    def forward(ctx, x, block, keep, *params):
        x = K.to_nhwc(x)
        units = block.tok_units()
        ds = block.tok_downsample()
        saved_all, descs = [], []
        sc = x
        if ds is not None:
            (sc, sv), d = _unit_fwd(x, ds[0], ds[1], False, None, keep)
            saved_all.append(sv)
            descs.append(d)
        h = x
        for conv, bn in units[:-1]:
            (h, sv), d = _unit_fwd(h, conv, bn, True, None, keep)
            saved_all.append(sv)
            descs.append(d)
        conv, bn = units[-1]
        (out, sv), d = _unit_fwd(h, conv, bn, True, sc, keep)
        saved_all.append(sv)
        descs.append(d)
        if keep:
            flat, layout = [], []
            for sv in saved_all:
                layout.append((tuple(t is not None for t in sv[:4]), sv[4]))
                flat.extend(t for t in sv[:4] if t is not None)
            ctx.save_for_backward(*flat)
            ctx.layout, ctx.descs, ctx.block = layout, descs, block
            ctx.cin = x.shape[1]
        cout = units[-1][0].out_channels
        return out if out.shape[1] == cout else out[:, :cout]

    @staticmethod
    def backward(ctx, dout):
        block, descs = ctx.block, ctx.descs
        units = block.tok_units()
        ds = block.tok_downsample()
        it = iter(ctx.saved_tensors)
        saved_all = [tuple(next(it) if present else None for present in lay) + (mode,) for lay, mode in ctx.layout]
        off = 1 if ds is not None else 0
        dout = K._dense_grad(dout, descs[-1].k)
        # last unit: ReLU mask from the block output.  The gradient that flows to the shortcut is g = dout * mask; r3: it
        # is not materialised — its consumers read dout and the forward's mask bits themselves (the identity shortcut
        # through the masked addend of the first conv's dgrad, a downsample unit through MASK_BITS of its own BatchNorm
        # backward), so the tail's backward apply writes 2 of its 8 bytes per element less.
        conv, bn = units[-1]
        tail_bits = saved_all[-1][2] if saved_all[-1][4] == K.MASK_BITS else None
        need_dx = ctx.needs_input_grad[0]
        lazy_g = tail_bits is not None and (
            (ds is None and need_dx and K.dgrad_masked_supported(descs[off])) or (ds is not None and K._MASKED_ADDEND))
        dh, g = _unit_bwd(saved_all[-1], descs[-1], conv, bn, dout, want_dres=not lazy_g)
        for i in range(len(units) - 2, 0, -1):
            conv, bn = units[i]
            dh, _ = _unit_bwd(saved_all[off + i], descs[off + i], conv, bn, dh)
        addend, addend_bits, compact = g, None, None
        if ds is not None:
            dd = descs[0]
            sv = saved_all[0]
            if lazy_g:   # the downsample unit has no activation of its own: its incoming gradient is dout under the tail's mask
                sv, g = (sv[0], sv[1], tail_bits, sv[3], K.MASK_BITS), dout
            if dd.stride > 1 and dd.r == 1 and dd.s == 1 and dd.pad == 0:
                # strided 1x1 shortcut: compact GEMM now, merged into dx below (no zero-filled scatter tensor)
                compact, _ = _unit_bwd(sv, dd, ds[0], ds[1], g, need_dx=need_dx, compact_dx=True)
                addend = None
            else:
                addend, _ = _unit_bwd(sv, dd, ds[0], ds[1], g, need_dx=need_dx)
        elif lazy_g:
            addend, addend_bits = dout, tail_bits
        conv, bn = units[0]
        dx, _ = _unit_bwd(saved_all[off], descs[off], conv, bn, dh, need_dx=need_dx, dx_addend=addend,
                          dx_addend_bits=addend_bits)
        if compact is not None and dx is not None:
            K.strided_add(dx, compact, descs[0].stride)
        if dx is not None and dx.shape[1] != ctx.cin:
            dx = dx[:, :ctx.cin]
        return (dx, None, None) + (None,) * (len(ctx.needs_input_grad) - 3)


def residual_block(x, block):
    params = [p for p in block.parameters()]
    keep = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
    return ResidualBlockFn.apply(x, block, keep, *params)


# ------------------------------------------------------------------------------------------------------ ResNet stem
class StemFn(torch.autograd.Function):
    """conv 7x7 s2 p3 (Cin <= 4) -> BN -> ReLU -> maxpool 3x3 s2 p1 (torchok/models/backbones/resnet.py:488-490,510,
    542-545).  Returns (act1 or None, pooled); the image is consumed straight from its NCHW fp32/bf16 layout.  With the
    standard 3/2/1 pool the BN-apply, ReLU and pooling run as ONE pass over the conv output (and the backward gathers
    the pooled gradient on the fly), so the 112x112 activation only exists in HBM when the caller asks for it."""

    @staticmethod
    def forward(ctx, image, conv, bn, pool, keep, need_act, *params):
        K.require_cuda(image, 'image')
        ctx.set_materialize_grads(False)  # `act` is usually unused: do not build (and convert) a zero gradient for it
        L = lib()
        st = K._st()
        n, c, h, w = image.shape
        if image.dtype not in (F32, BF16):
            image = image.float()
        image = image.contiguous()
        k = conv.out_channels
        P, Q, H2, W2 = (C.c_int() for _ in range(4))
        L.tok_stem_geometry(h, w, C.byref(P), C.byref(Q), C.byref(H2), C.byref(W2))
        P, Q, H2, W2 = P.value, Q.value, H2.value, W2.value
        dev = image.device
        xs2d = torch.empty((n, H2, W2, 16), dtype=BF16, device=dev)
        L.tok_stem_pack_input(n, c, h, w, int(image.dtype == BF16), K._p(image), K._p(xs2d), st)
        wp = torch.empty((k, 256), dtype=BF16, device=dev)
        wt = conv.weight
        if not K.is_krsc(wt):
            wt.data = wt.data.contiguous(memory_format=torch.channels_last)
        L.tok_stem_pack_weight(k, c, K._p(wt), K._p(wp), st)
        bs = bn.state()
        y = torch.empty((n, P, Q, k), dtype=BF16, device=dev)
        small = torch.empty((4, k), dtype=F32, device=dev)
        rows = n * P * Q
        if bs.training:
            L.tok_stem_conv_fprop(n, h, w, k, K._p(xs2d), K._p(wp), K._p(y), K._p(bs.acc[0]), K._p(bs.acc[1]), st)
            L.tok_bn_finalize_train(k, float(rows), K._p(bs.acc[0]), K._p(bs.acc[1]), K._p(bs.weight), K._p(bs.bias),
                                    bs.eps, bs.momentum, K._p(bs.running_mean), K._p(bs.running_var), K._p(small[0]),
                                    K._p(small[1]), K._p(small[2]), K._p(small[3]), st)
        else:
            L.tok_stem_conv_fprop(n, h, w, k, K._p(xs2d), K._p(wp), K._p(y), None, None, st)
            L.tok_bn_finalize_eval(k, K._p(bs.running_mean), K._p(bs.running_var), K._p(bs.weight), K._p(bs.bias),
                                   bs.eps, K._p(small[0]), K._p(small[1]), st)
        pk, ps, pp = pool
        fused = (pk, ps, pp) == (3, 2, 1) and P % 2 == 0 and Q % 2 == 0
        act = None
        if fused:
            PP, QQ = (P + 2 - 3) // 2 + 1, (Q + 2 - 3) // 2 + 1
            if need_act:
                act = torch.empty_like(y)
            pooled_raw = torch.empty((n, PP, QQ, k), dtype=BF16, device=dev)
            arg = torch.empty((n, PP, QQ, k), dtype=torch.uint8, device=dev)
            L.tok_stem_bn_relu_pool_fwd(n, P, Q, k, K._p(y), K._p(small[0]), K._p(small[1]), K._p(act),
                                        K._p(pooled_raw), K._p(arg), st)
            pooled = pooled_raw.permute(0, 3, 1, 2)
            if act is not None:
                act = act.permute(0, 3, 1, 2)
        else:
            act = torch.empty_like(y) if keep else y
            L.tok_bn_apply(rows, k, K._p(y), K._p(small[0]), K._p(small[1]), None, 1, K._p(act), st)
            act = act.permute(0, 3, 1, 2)
            pooled, arg = K.maxpool_fwd(act, pk, ps, pp)
        if keep:
            ctx.save_for_backward(xs2d, y, small, arg)
            ctx.meta = (conv, bn, pool, (n, c, h, w), (P, Q), fused)
        return act, pooled

    @staticmethod
    def backward(ctx, d_act, d_pooled):
        conv, bn, (pk, ps, pp), (n, c, h, w), (P, Q), fused = ctx.meta
        xs2d, y, small, arg = ctx.saved_tensors
        L = lib()
        st = K._st()
        k = conv.out_channels
        dev = y.device
        rows = n * P * Q
        if d_pooled is None and d_act is None:
            return (None,) * len(ctx.needs_input_grad)
        acc = bn._tok_acc
        gw, gb = _bn_grads(bn)
        coefs = torch.empty((3, k), dtype=F32, device=dev)
        dy = torch.empty_like(y)
        if fused and d_pooled is not None:
            dp = K._dense_grad(d_pooled, k)
            da = K._dense_grad(d_act, k) if d_act is not None else None
            L.tok_stem_bwd_reduce(n, P, Q, k, K._p(dp), K._p(arg), K._p(da), K._p(y), K._p(small[0]), K._p(small[1]),
                                  K._p(acc[2]), K._p(acc[3]), st)
            L.tok_bn_bwd_finalize(k, float(rows), K._p(acc[2]), K._p(acc[3]), K._p(small[2]), K._p(small[3]),
                                  K._p(bn.weight), K._p(coefs[0]), K._p(coefs[1]), K._p(coefs[2]), K._p(gw), K._p(gb),
                                  1, st)
            L.tok_stem_bwd_apply(n, P, Q, k, K._p(dp), K._p(arg), K._p(da), K._p(y), K._p(small[0]), K._p(small[1]),
                                 K._p(coefs[0]), K._p(coefs[1]), K._p(coefs[2]), K._p(dy), st)
        else:
            if d_pooled is not None:
                PP, QQ = arg.shape[1], arg.shape[2]
                dout = K.maxpool_bwd(K._dense_grad(d_pooled, k), arg, (n, k, P, Q), k, pk, ps, pp)
                dout2 = K._dense_grad(d_act, k) if d_act is not None else None
            else:
                dout, dout2 = K._dense_grad(d_act, k), None
            # ReLU mask rebuilt from y and the forward's scale/shift (MASK_Y): `act` is not kept for the backward
            L.tok_bn_bwd_reduce2(rows, k, K._p(dout), K._p(dout2), K._p(y), K.MASK_Y, None, K._p(small[0]),
                                 K._p(small[1]), K._p(acc[2]), K._p(acc[3]), st)
            L.tok_bn_bwd_finalize(k, float(rows), K._p(acc[2]), K._p(acc[3]), K._p(small[2]), K._p(small[3]),
                                  K._p(bn.weight), K._p(coefs[0]), K._p(coefs[1]), K._p(coefs[2]), K._p(gw), K._p(gb),
                                  1, st)
            L.tok_bn_bwd_apply2(rows, k, K._p(dout), K._p(dout2), K._p(y), K.MASK_Y, None, K._p(small[0]),
                                K._p(small[1]), K._p(coefs[0]), K._p(coefs[1]), K._p(coefs[2]), K._p(dy), None, st)
        if conv.weight.requires_grad:
            dwp = torch.zeros((k, 256), dtype=F32, device=dev)
            L.tok_stem_conv_wgrad(n, h, w, k, K._p(xs2d), K._p(dy), K._p(dwp), st)
            L.tok_stem_unpack_wgrad(k, c, K._p(dwp), K._p(K.grad_buffer(conv.weight)), 1, st)
        for prm in (conv.weight, bn.weight, bn.bias):
            K.grad_ready(prm)
        if ctx.needs_input_grad[0]:
            raise NotImplementedError('gradient w.r.t. the input image is not produced by the stem kernel')
        return (None,) * len(ctx.needs_input_grad)


def stem(image, conv, bn, pool=(3, 2, 1), need_act=True):
    params = _params(conv, bn)
    keep = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return StemFn.apply(image, conv, bn, pool, keep, need_act, *params)


# ------------------------------------------------------------------------------------------------------ small modules
class ReLU(nn.ReLU):
    """Marker module: the ReLU itself is applied inside the fused unit that precedes it."""

    def forward(self, x):
        return x


class MaxPool2d(nn.MaxPool2d):
    def forward(self, x):
        return K.MaxPoolFn.apply(x, _single(self.kernel_size, 'kernel_size'), _single(self.stride, 'stride'),
                                 _single(self.padding, 'padding'))
