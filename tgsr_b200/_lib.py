"""ctypes binding of libtgsr_sm100.so (the C ABI in include/tgsr.h).

The library is mandatory: importing a kernel entry point without the built .so, or on a host
without a CUDA device, raises -- there is no CPU or PyTorch fallback for the arithmetic.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libtgsr_sm100.so')

ACT_NONE, ACT_GLU, ACT_TANH = 0, 1, 2
FMT_BF16, FMT_F16 = 0, 1
CONV_REVERSE = 1
CONV_L2_IN_FIRST, CONV_L2_OUT_LAST = 2, 4

_lib = None
_lock = threading.Lock()
_inited_devices = set()

c_f32p = C.c_void_p
_SIGS = {
    'tgsr_version': ([], C.c_int),
    'tgsr_error_string': ([C.c_int], C.c_char_p),
    'tgsr_last_cuda_error': ([], C.c_int),
    'tgsr_init': ([C.c_int], C.c_int),
    'tgsr_sm_count': ([], C.c_int),
    'tgsr_launch_count': ([], C.c_ulonglong),
    'tgsr_attn_fwd_f32': ([C.c_void_p] * 6 + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_attn_f32_ws_floats': ([C.c_int] * 3, C.c_size_t),
    'tgsr_attn_f32_tc_supported': ([C.c_int] * 6, C.c_int),
    'tgsr_attn_fwd_f32_tc': ([C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_func_attn_fwd_f32': ([C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p], C.c_int),
    'tgsr_conv2d_f32': ([C.c_void_p] * 6 + [C.c_float, C.c_float, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p], C.c_int),
    'tgsr_bn_fold_f32': ([C.c_void_p] * 4 + [C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p], C.c_int),
    'tgsr_linear_glu_f32': ([C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_void_p], C.c_int),
    'tgsr_ca_reparam_f32': ([C.c_void_p] * 3 + [C.c_int] * 2 + [C.c_void_p], C.c_int),
    'tgsr_nchw_to_blocked': ([C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p], C.c_int),
    'tgsr_blocked_to_nchw': ([C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p], C.c_int),
    'tgsr_packed_conv3x3_elems': ([C.c_int] * 3, C.c_size_t),
    'tgsr_pack_conv3x3': ([C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_conv3x3_tc': ([C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                         C.c_void_p, C.c_int, C.c_int] + [C.c_int] * 9 + [C.c_void_p], C.c_int),
    'tgsr_pack_head': ([C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p], C.c_int),
    'tgsr_img_head_tc': ([C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
                         + [C.c_int] * 7 + [C.c_void_p, C.c_void_p], C.c_int),
    'tgsr_conv_in_blocked': ([C.c_void_p] * 6 + [C.c_int] * 7 + [C.c_void_p], C.c_int),
    'tgsr_bn_stats_f32': ([C.c_void_p] * 5 + [C.c_float, C.c_float, C.c_void_p] + [C.c_int] * 3 + [C.c_void_p], C.c_int),
    'tgsr_bn_act_f32': ([C.c_void_p] * 7 + [C.c_int] * 4 + [C.c_void_p], C.c_int),
    'tgsr_bn_act_bwd_f32': ([C.c_void_p] * 10 + [C.c_int] * 4 + [C.c_void_p], C.c_int),
    'tgsr_conv_weight_flip_f32': ([C.c_void_p, C.c_void_p] + [C.c_int] * 3 + [C.c_void_p], C.c_int),
    'tgsr_sumpool2x2_f32': ([C.c_void_p, C.c_void_p] + [C.c_int] * 3 + [C.c_void_p], C.c_int),
    'tgsr_conv2d_wgrad_f32': ([C.c_void_p] * 3 + [C.c_int] * 7 + [C.c_void_p], C.c_int),
    'tgsr_tanh_head_bwd_f32': ([C.c_void_p] * 3 + [C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p], C.c_int),
    'tgsr_attn_bwd_f32': ([C.c_void_p] * 9 + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_damsm_pairs_fwd_f32': ([C.c_void_p] * 3 + [C.c_float] * 3 + [C.c_void_p] * 4 + [C.c_int] * 6 + [C.c_void_p], C.c_int),
    'tgsr_damsm_ce_f32': ([C.c_void_p] * 3 + [C.c_float] + [C.c_void_p] * 3 + [C.c_int, C.c_void_p], C.c_int),
    'tgsr_damsm_pairs_bwd_f32': ([C.c_void_p] * 11 + [C.c_float] * 3 + [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_sent_sim_f32': ([C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p], C.c_int),
    'tgsr_sent_bwd_f32': ([C.c_void_p] * 7 + [C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p], C.c_int),
    'tgsr_mse_f32': ([C.c_void_p] * 4 + [C.c_size_t, C.c_void_p], C.c_int),
    'tgsr_mse_bwd_f32': ([C.c_void_p] * 4 + [C.c_size_t, C.c_void_p], C.c_int),
    'tgsr_adam_f32': ([C.c_void_p] * 4 + [C.c_size_t] + [C.c_float] * 7 + [C.c_void_p], C.c_int),
    'tgsr_conv_tf32_supported': ([C.c_int] * 4, C.c_int),
    'tgsr_conv_prep_tf32': ([C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p], C.c_int),
    'tgsr_conv2d_tf32': ([C.c_void_p] * 3 + [C.c_int] * 7 + [C.c_void_p], C.c_int),
    'tgsr_conv2d_wgrad_tf32': ([C.c_void_p] * 3 + [C.c_int] * 7 + [C.c_void_p], C.c_int),
    'tgsr_adam_devstep_f32': ([C.c_void_p] * 4 + [C.c_size_t] + [C.c_float] * 4 + [C.c_void_p, C.c_float, C.c_void_p], C.c_int),
    'tgsr_quantize_u8': ([C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p], C.c_int),
    'tgsr_attn_word_stats_f32': ([C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_float, C.c_int, C.c_void_p], C.c_int),
    'tgsr_kl_loss_f32': ([C.c_void_p] * 3 + [C.c_size_t, C.c_void_p], C.c_int),
    'tgsr_kl_loss_bwd_f32': ([C.c_void_p] * 5 + [C.c_size_t, C.c_void_p], C.c_int),
    'tgsr_cosine_similarity_f32': ([C.c_void_p] * 4 + [C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_void_p], C.c_int),
    'tgsr_cosine_similarity_bwd_f32': ([C.c_void_p] * 6 + [C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_void_p], C.c_int),
    'tgsr_resize_u8_ws_bytes': ([C.c_int] * 6, C.c_size_t),
    'tgsr_resize_u8': ([C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p], C.c_int),
    'tgsr_gaussian_blur_u8': ([C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_float, C.c_int, C.c_void_p], C.c_int),
    'tgsr_normalize_u8_f32': ([C.c_void_p] * 2 + [C.c_int] * 4 + [C.c_void_p], C.c_int),
    'tgsr_pyramid_supported': ([C.c_int] * 5, C.c_int),
    'tgsr_pyramid_levels_u8': ([C.c_void_p] + [C.c_int] * 5 + [C.c_void_p] * 5 + [C.c_float, C.c_void_p], C.c_int),
    'tgsr_attn_sfrag_bytes': ([C.c_int], C.c_size_t),
    'tgsr_attn_sfrag_bytes_t': ([C.c_int, C.c_int], C.c_size_t),
    'tgsr_attn_project': ([C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_attn_project_multi': ([C.c_void_p] * 7 + [C.c_int] * 6 + [C.c_void_p], C.c_int),
    'tgsr_attn_fwd_blocked': ([C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                               C.c_int, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_bn16_stats': ([C.c_void_p] * 5 + [C.c_float, C.c_float, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p], C.c_int),
    'tgsr_bn16_act': ([C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_bn16_act_bwd': ([C.c_void_p] * 10 + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_conv3x3_wgrad_tc': ([C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p], C.c_int),
    'tgsr_head_wgrad16': ([C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p], C.c_int),
    'tgsr_maxpool3s2_nhwc16': ([C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_maxpool3s2_nhwc16_bwd': ([C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_head_dgrad16': ([C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p], C.c_int),
    'tgsr_relu_bwd_nhwc16': ([C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p], C.c_int),
    'tgsr_upsample2x_blocked': ([C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p], C.c_int),
    'tgsr_sumpool2x2_blocked': ([C.c_void_p, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_lstm_prepare_f32': ([C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p], C.c_int),
    'tgsr_lstm_bidir_fwd_tab_f32': ([C.c_void_p] * 6 + [C.c_int] * 5 + [C.c_void_p], C.c_int),
    'tgsr_lstm_bidir_fwd_f32': ([C.c_void_p] * 10 + [C.c_int] * 6 + [C.c_void_p], C.c_int),
}


def load() -> C.CDLL:
    """dlopen the library (no GPU needed) and attach signatures."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError(
                    'tgsr_b200: %s is missing -- build it with `python -m tgsr_b200.build` '
                    '(there is no fallback path)' % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, (argtypes, restype) in _SIGS.items():
                fn = getattr(lib, name)          # AttributeError if the .so does not export it
                fn.argtypes = argtypes
                fn.restype = restype
            _lib = lib
    return _lib


def declared_symbols():
    return sorted(_SIGS)


def _ensure_device(dev: torch.device) -> C.CDLL:
    lib = load()
    if dev.type != 'cuda':
        raise RuntimeError('tgsr_b200 kernels need CUDA tensors (got %s); there is no CPU path' % dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _inited_devices:
        with torch.cuda.device(idx):
            check(lib.tgsr_init(idx), 'tgsr_init')
        _inited_devices.add(idx)
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        lib = load()
        msg = lib.tgsr_error_string(status).decode()
        extra = ''
        if status == -3:
            extra = ' (cudaError %d)' % lib.tgsr_last_cuda_error()
        raise RuntimeError('tgsr_b200: %s failed: %s%s' % (what, msg, extra))


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# ---------------------------------------------------------------- derived-parameter cache keys
# The kernels write parameters and BatchNorm running statistics through raw pointers (fused Adam, bn_stats), which
# bumps neither data_ptr nor tensor._version.  Every such writer calls bump_weights_epoch(); every derived-parameter
# cache (folded BN, packed 16-bit operands, LSTM tables, captured eval graphs) keys on tensor_key(), which includes
# the epoch, so an eval after training steps refolds / repacks instead of replaying stale weights.
_weights_epoch = 0


def weights_epoch() -> int:
    return _weights_epoch


def bump_weights_epoch() -> None:
    global _weights_epoch
    _weights_epoch += 1


def no_grad_path(what: str, *tensors) -> None:
    """The frozen-text SR training step is the only backward path built (SURVEY.md 8d cfg 5: text and image encoders
    frozen).  An input that would receive a gradient in the reference but has no backward kernel here must not silently
    train nothing: raise instead."""
    if torch.is_grad_enabled():
        for t in tensors:
            if t is not None and t.requires_grad:
                raise NotImplementedError(
                    'tgsr_b200: %s: a gradient was requested for an input whose backward is not implemented (the CUDA path '
                    'trains the generator against frozen text / image-text encoders); detach() it or freeze its parameters' % what)


_manual_epoch = 0


def bump_manual_epoch() -> None:
    """invalidate_caches(): also drops the caches of modules the library's own writers never touch (frozen encoders)."""
    global _manual_epoch
    _manual_epoch += 1
    bump_weights_epoch()


def tensor_key(t, trained: bool = True):
    """Cache key of a parameter / buffer.  trained=False: a tensor only the user can change (the frozen text encoder's
    weights): it ignores the epoch the optimiser and BatchNorm kernels bump every training step."""
    return None if t is None else (t.data_ptr(), t._version, str(t.device), _weights_epoch if trained else _manual_epoch)


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_fn_cache = {}


def call(name: str, dev: torch.device, *args) -> None:
    """Enqueue one C-ABI entry point on torch's current stream of `dev` (hot path: keep it lean)."""
    idx = dev.index
    if dev.type != 'cuda' or idx is None or idx not in _inited_devices:
        _ensure_device(dev)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(load(), name)
    if idx != torch.cuda.current_device():
        with torch.cuda.device(idx):
            status = fn(*args, C.c_void_p(torch.cuda.current_stream(idx).cuda_stream))
    else:
        status = fn(*args, C.c_void_p(torch.cuda.current_stream(idx).cuda_stream))
    if status != 0:
        check(status, name)
