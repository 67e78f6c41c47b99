"""tgsr_b200 -- B200-native (sm_100a) implementation of the TGSR text-conditioned SR generator
forward path behind the reference's nn.Module API.  See DESIGN.md / INTEGRATION.md."""
from .config import cfg, cfg_from_file, cfg_from_dict, set_eval_defaults  # noqa: F401
from .attention import GlobalAttentionGeneral, func_attention  # noqa: F401

__version__ = '0.1.0'

_PRECISION = 'fp32'
SMOKE_PRECISIONS = ('fp32',)
DEFAULT_BENCH_PRECISION = 'fp32'      # switched to 'bf16' once the tcgen05 path is parity-green


def set_precision(mode: str) -> None:
    """'fp32': exact-precision CUDA-core kernels.  'bf16': tcgen05 implicit-GEMM conv stages on
    bf16 blocked activations with fp32 accumulation (the throughput path)."""
    global _PRECISION
    if mode not in ('fp32', 'bf16'):
        raise ValueError(mode)
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION
