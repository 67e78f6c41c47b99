"""tgsr_b200 -- B200-native (sm_100a) implementation of the TGSR text-conditioned SR generator
forward path behind the reference's nn.Module API.  See DESIGN.md / INTEGRATION.md."""
from .config import cfg, cfg_from_file, cfg_from_dict, set_eval_defaults  # noqa: F401
from .attention import GlobalAttentionGeneral, func_attention  # noqa: F401

__version__ = '0.1.0'

_PRECISION = 'fp32'
SMOKE_PRECISIONS = ('fp32', 'mixed16')
DEFAULT_BENCH_PRECISION = 'mixed16'
_MODES = ('fp32', 'bf16', 'fp16', 'mixed16')


def set_precision(mode: str) -> None:
    """'fp32'   : exact-precision CUDA-core kernels (fp32 NCHW everywhere).
    'bf16'    : tcgen05 implicit-GEMM conv stages, bf16 operands/activations, fp32 accumulation.
    'fp16'    : same kernels with fp16 operands/activations.
    'mixed16' : bf16 for G_SR_NET_low, fp16 for NetG_high* -- the throughput configuration that meets
                the 2e-2 max-abs bar on the shipped checkpoint (pure bf16 does not)."""
    global _PRECISION
    if mode not in _MODES:
        raise ValueError('precision must be one of %s' % (_MODES,))
    _PRECISION = mode


def invalidate_caches() -> None:
    """Drop every derived-parameter cache (folded BatchNorm, packed 16-bit operands, LSTM tables, captured eval graphs).
    The caches follow `tensor._version`, `load_state_dict`, the library's own optimiser and BatchNorm updates; writes that
    bypass the version counter -- `param.data.copy_(...)` / `.data.mul_(...)` as in the reference's `load_params`
    (miscc/utils.py:462-464), or raw-pointer writes from other libraries -- need this call before the next eval forward."""
    from ._lib import bump_manual_epoch
    bump_manual_epoch()


def get_precision() -> str:
    return _PRECISION


def is_16bit() -> bool:
    return _PRECISION != 'fp32'


def fmt_for(net: str = 'low') -> int:
    """Operand format (C-ABI TGSR_FMT_*) for 'low' (G_SR_NET_low / single blocks) or 'high' (NetG_high*)."""
    from ._lib import FMT_BF16, FMT_F16
    if _PRECISION == 'fp16' or (_PRECISION == 'mixed16' and net == 'high'):
        return FMT_F16
    return FMT_BF16
