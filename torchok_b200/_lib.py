"""ctypes binding of libtokb200.so (include/tokb200.h).

The reference has no FFI: torchok modules call torch.nn and torch dispatches to cuDNN/cuBLAS
(e.g. torchok/models/modules/bricks/convbnact.py:38-53).  This file is the stub a maintainer would add to call the
sm_100a kernels instead; see INTEGRATION.md.  There is NO CPU fallback: if the shared library is missing every call
raises `TokLibraryError`.
"""
import ctypes as C
import os
import re
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, 'libtokb200.so')
HEADER_PATH = os.path.join(_ROOT, 'include', 'tokb200.h')


class TokLibraryError(RuntimeError):
    pass


class TokError(RuntimeError):
    """A libtokb200 entry point returned a negative tokStatus."""


class tokConvDesc(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('n', 'h', 'w', 'c', 'k', 'r', 's', 'stride', 'pad', 'dil', 'wk', 'wc')]


class tokPeerArenas(C.Structure):
    """include/tokb200.h: pointer table of tok_peer_step (8 = TOK_PEER_MAX_RANKS)."""
    _fields_ = [('world', C.c_int), ('rank', C.c_int), ('master', C.c_void_p * 8), ('grad', C.c_void_p * 8),
                ('shadow', C.c_void_p * 8), ('flags', C.c_void_p * 8)]


_vp, _i, _ll, _f, _d, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_size_t
_pd = C.POINTER(tokConvDesc)
_pi = C.POINTER(C.c_int)

# name -> (restype, argtypes).  Status-returning functions (restype int, name not in _RAW) are wrapped to raise.
_SIGS = {
    'tok_version': (_i, []),
    'tok_last_error': (C.c_char_p, []),
    'tok_device_ok': (_i, []),
    'tok_debug_conv_profile': (_i, [_vp, _i]),
    'tok_debug_attn_profile': (_i, [_vp, _i]),
    'tok_conv_out_hw': (None, [_pd, _pi, _pi]),
    'tok_conv_halo_caps': (_i, [_pd]),
    'tok_conv_fprop': (_i, [_pd, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    'tok_conv_fprop_bn': (_i, [_pd, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_conv_dgrad_workspace_bytes': (_sz, [_pd]),
    'tok_conv_dgrad': (_i, [_pd, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_conv_dgrad_masked_supported': (_i, [_pd]),
    'tok_conv_dgrad_masked': (_i, [_pd, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_conv_wgrad': (_i, [_pd, _vp, _vp, _vp, _vp]),
    'tok_linear_fwd': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'tok_linear_dgrad': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    'tok_linear_dgrad_add': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'tok_patch_embed_supported': (_i, [_i, _i, _i, _i, _i]),
    'tok_patchify': (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'tok_patch_embed_fwd': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_patch_embed_bwd': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_patch_merge': (_i, [_i, _i, _i, _i, _vp, _vp, _i, _vp]),
    'tok_linear_dgrad_gelu': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_linear_wgrad': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    'tok_stem_geometry': (None, [_i, _i, _pi, _pi, _pi, _pi]),
    'tok_stem_pack_input': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'tok_stem_pack_weight': (_i, [_i, _i, _vp, _vp, _vp]),
    'tok_stem_unpack_wgrad': (_i, [_i, _i, _vp, _vp, _i, _vp]),
    'tok_stem_conv_fprop': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_stem_conv_wgrad': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'tok_bn_finalize_train': (_i, [_i, _d, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_bn_finalize_eval': (_i, [_i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    'tok_bn_finalize_train_cv': (_i, [_i, _i, _d, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_bn_finalize_eval_cv': (_i, [_i, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    'tok_bn_bwd_finalize_cv': (_i, [_i, _i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    'tok_bn_bwd_reduce2_finalize_cv': (_i, [_ll, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                            _vp, _vp, _i, _vp, _vp]),
    'tok_bn_apply': (_i, [_ll, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    'tok_bn_apply_train_supported': (_i, [_ll, _i]),
    'tok_bn_apply_train': (_i, [_ll, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    'tok_bn_apply_bits_train': (_i, [_ll, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_bn_apply_chain': (_i, [_ll, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp]),
    'tok_bn_apply_bits_chain': (_i, [_ll, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    'tok_bn_bwd_reduce': (_i, [_ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_bn_bwd_finalize': (_i, [_i, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    'tok_bn_bwd_apply': (_i, [_ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_bn_apply_bits': (_i, [_ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_bn_bwd_reduce2': (_i, [_ll, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_bn_bwd_reduce2_finalize': (_i, [_ll, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                         _vp, _vp, _i, _vp, _vp]),
    'tok_bn_bwd_fused_cv': (_i, [_ll, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    'tok_bn_bwd_apply2': (_i, [_ll, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_stem_bn_relu_pool_fwd': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_stem_bwd_reduce': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_stem_bwd_apply': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_strided_add': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'tok_maxpool_fwd': (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'tok_maxpool_bwd': (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'tok_gap_fwd': (_i, [_i, _i, _i, _i, _vp, _vp, _vp]),
    'tok_gap_bwd': (_i, [_i, _i, _i, _vp, _vp, _vp]),
    'tok_gap_bwd_max': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'tok_softmax_xent': (_i, [_i, _i, _ll, _vp, _vp, _vp, _vp, _f, _f, _vp, _ll, _vp, _vp]),
    'tok_layernorm_fwd': (_i, [_ll, _i, _vp, _vp, _vp, _f, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    'tok_layernorm_bwd': (_i, [_ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    'tok_layernorm_has_dxsum': (_i, [_i]),
    'tok_gelu_fwd': (_i, [_ll, _vp, _vp, _vp]),
    'tok_gelu_bwd': (_i, [_ll, _i, _vp, _vp, _vp, _vp, _vp]),
    'tok_window_attn_fwd': (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'tok_window_attn_bwd': (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_fuse_sum_fwd': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    'tok_fuse_sum_bwd': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'tok_bilinear_fwd': (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    'tok_bilinear_bwd': (_i, [_i, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    'tok_softmax_xent_small': (_i, [_ll, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _ll, _vp]),
    'tok_dice_stats': (_i, [_ll, _i, _i, _vp, _vp, _vp, _vp]),
    'tok_dice_finalize': (_i, [_i, _vp, _f, _f, _i, _vp, _vp, _vp]),
    'tok_dice_bwd': (_i, [_ll, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_rownorm_fwd': (_i, [_i, _i, _vp, _i, _f, _vp, _i, _vp, _vp]),
    'tok_rownorm_bwd': (_i, [_i, _i, _vp, _i, _vp, _f, _vp, _i, _i, _vp, _i, _i, _vp]),
    'tok_arcface_margin_fwd': (_i, [_i, _i, _i, _vp, _vp, _vp, _i, _vp, _ll, _f, _f, _i, _vp, _vp]),
    'tok_arcface_margin_bwd': (_i, [_i, _vp, _i, _vp, _vp, _ll, _f, _f, _i, _vp]),
    'tok_contrastive_fwd': (_i, [_i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    'tok_contrastive_bwd': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    'tok_l2_normalize_rows': (_i, [_i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    'tok_topk_candidates': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_topk_rerank': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_nchw_to_nhwc': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'tok_nhwc_to_nchw': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'tok_sgd_step': (_i, [_ll, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _f, _i, _vp]),
    'tok_adam_step': (_i, [_ll, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _f, _i, _i, _f, _vp]),
    'tok_sgd_step_dev': (_i, [_ll, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _i, _f, _i, _vp]),
    'tok_adam_step_dev': (_i, [_ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _f, _i, _vp]),
    'tok_sgd_step_dev_groups': (_i, [_ll, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _i, _f, _i, _vp, _vp, _vp, _i, _vp]),
    'tok_adam_step_dev_groups': (_i, [_ll, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _f, _i, _vp, _vp, _vp, _vp,
                                      _i, _vp]),
    'tok_cpb_bias_fwd': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_cpb_bias_bwd': (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_nearest_fwd': (_i, [_i, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    'tok_nearest_bwd': (_i, [_i, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp]),
    'tok_channel_scale': (_i, [_i, _ll, _i, _vp, _vp, _vp, _vp]),
    'tok_spatial_gather_fwd': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_spatial_gather_bwd': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_object_attn_fwd': (_i, [_i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    'tok_object_attn_bwd': (_i, [_i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'tok_ipc_alloc': (_i, [_sz, C.POINTER(C.c_void_p), _vp]),
    'tok_ipc_free': (_i, [_vp]),
    'tok_ipc_open': (_i, [C.c_char_p, C.POINTER(C.c_void_p)]),
    'tok_ipc_close': (_i, [_vp]),
    'tok_peer_flag_bytes': (_sz, []),
    'tok_peer_step': (_i, [C.POINTER(tokPeerArenas), _i, _ll, _ll, _i, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _f, _vp, _vp,
                           _vp, _vp, _i, _i, _vp]),
    'tok_cast_f32_bf16': (_i, [_ll, _vp, _vp, _vp]),
    'tok_pad_weight': (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'tok_unpad_wgrad_add': (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp]),
}
_RAW = {'tok_conv_halo_caps', 'tok_conv_dgrad_masked_supported', 'tok_bn_apply_train_supported', 'tok_peer_flag_bytes', 'tok_debug_conv_profile', 'tok_debug_attn_profile', 'tok_layernorm_has_dxsum', 'tok_patch_embed_supported', 'tok_version', 'tok_last_error', 'tok_device_ok', 'tok_conv_out_hw', 'tok_conv_dgrad_workspace_bytes',
        'tok_stem_geometry'}


def header_symbols(path=HEADER_PATH):
    """Every function name declared in include/tokb200.h (used by the symbol-coverage test)."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(tok_[a-z0-9_]+)\s*\(', text)))


class _Lib:
    def __init__(self, path):
        if not os.path.exists(path):
            raise TokLibraryError(
                f'{path} is missing: build it with `make` (or `python -c "import __graft_entry__ as g; g.build()"`). '
                'torchok_b200 has no CPU or torch.nn fallback for its kernels.')
        self._dll = C.CDLL(path)
        self.path = path
        self.tracer = None
        self.launches = 0  # kernels-launching ABI calls made through this binding (bench.py's gpu_launches claim)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self._dll, name)
            fn.restype = res
            fn.argtypes = args
            if res is _i and name not in _RAW:
                fn = self._checked(name, fn)
            setattr(self, name, fn)

    def _checked(self, name, fn):
        def call(*a):
            tr = self.tracer
            if tr is not None:     # bench.py's per-call CUDA-event timing (off on the product path)
                tok = tr.before(name, a)
                rc = fn(*a)
                tr.after(tok)
            else:
                rc = fn(*a)
            self.launches += 1
            if rc != 0:
                raise TokError(f'{name} failed ({rc}): {self._dll.tok_last_error().decode()}')
            return rc
        call.__name__ = name
        return call

    def has(self, name):
        try:
            getattr(self._dll, name)
            return True
        except AttributeError:
            return False


_lib = None
_lock = threading.Lock()


def lib():
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                _lib = _Lib(LIB_PATH)
    return _lib


def build(verbose=False):
    """Compile libtokb200.so for sm_100a in-tree (Makefile at the repo root)."""
    out = subprocess.run(['make', '-C', _ROOT, '-j8'], capture_output=True, text=True)
    if out.returncode != 0:
        raise TokLibraryError('make failed:\n' + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH
