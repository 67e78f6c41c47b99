"""CPU-only checks of the host side: C-ABI library loads and exports every symbol that
include/tgsr.h declares, the module API mirrors the reference (state-dict keys, constructors),
config merge semantics, and the product path fails loudly without a GPU (no fallback)."""
import os
import re

import pytest
import torch

import tgsr_b200
from tgsr_b200 import _lib, config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_in_header():
    src = open(os.path.join(ROOT, 'include', 'tgsr.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(tgsr_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from tgsr_b200 import build
    build.build()                                  # cross-compiles for sm_100a without a GPU
    lib = _lib.load()
    names = _declared_in_header()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), 'libtgsr_sm100.so does not export %s' % n
    assert sorted(_lib.declared_symbols()) == names, 'ctypes signatures out of sync with include/tgsr.h'
    assert lib.tgsr_version() >= 100
    assert lib.tgsr_error_string(-2) == b'unsupported dimensions for this kernel'


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    tgsr_b200.set_eval_defaults()
    att = tgsr_b200.GlobalAttentionGeneral(32, 256)
    with pytest.raises(RuntimeError, match='CUDA'):
        att(torch.zeros(1, 32, 4, 4), torch.zeros(1, 256, 5))


def test_state_dict_contract(golden):
    """Shipped Checkpoint/face_S8 state dicts load strict=True; derived buffers stay out."""
    tgsr_b200.set_eval_defaults()
    from tgsr_b200 import nets, nets16, text_encoder
    w = golden('face_S8_weights.pt')
    gl = nets.G_SR_NET_low(); gl.load_state_dict(w['netG'], strict=True)
    # 104 keys / 911 488 elements and 121 keys / 284 724 elements (num_batches_tracked included)
    assert len(gl.state_dict()) == 104
    assert sum(v.numel() for v in gl.state_dict().values()) == 911488
    assert {k: tuple(v.shape) for k, v in gl.state_dict().items()} == {k: tuple(v.shape) for k, v in w['netG'].items()}
    gh = nets.NetG_highweight(weightmap=False, low='lr'); gh.load_state_dict(w['netGH'], strict=True)
    assert len(gh.state_dict()) == 121 and 'a' not in gh.state_dict()
    assert sum(v.numel() for v in gh.state_dict().values()) == 284724
    assert {k: tuple(v.shape) for k, v in gh.state_dict().items()} == {k: tuple(v.shape) for k, v in w['netGH'].items()}
    nets.NetG_high().load_state_dict(w['netGH'], strict=True)
    assert set(gl.state_dict()) == set(w['netG']) and set(gh.state_dict()) == set(w['netGH'])
    x16 = golden('x16.pt')
    m = nets16.NetG_high(); m.load_state_dict(x16['sdH'], strict=True)
    assert set(m.state_dict()) == set(x16['sdH'])
    g16 = nets16.G_SR_NET_low()
    sd = g16.state_dict()
    assert len(sd) == 136 and len({v.data_ptr() for v in sd.values()}) == 71     # aliases
    enc = text_encoder.RNN_ENCODER(41, nhidden=256)
    enc.load_state_dict(golden('rnn_encoder.pt')['sd'], strict=True)
    assert enc.nhidden == 128 and enc.init_hidden(3)[0].shape == (2, 3, 128)


def test_config_merge(tmp_path):
    p = tmp_path / 'c.yml'
    p.write_text("TREE:\n    BRANCH_NUM: 4\nGAN:\n    GF_DIM: 32\n")
    config.cfg_from_file(str(p))
    assert config.cfg.TREE.BRANCH_NUM == 4 and config.cfg.GAN.GF_DIM == 32
    p.write_text("NOPE: 1\n")
    with pytest.raises(KeyError):
        config.cfg_from_file(str(p))
    p.write_text("GAN:\n    GF_DIM: 'x'\n")
    with pytest.raises(ValueError):
        config.cfg_from_file(str(p))


def test_glu_odd_channels_error():
    from tgsr_b200 import ops
    if torch.cuda.is_available():
        pytest.skip('checked on the GPU path')
    with pytest.raises(RuntimeError):
        ops.conv2d_fused(torch.zeros(1, 3, 4, 4), torch.zeros(5, 3, 3, 3), None, None, 1)


def test_pyramid_host_logic_and_argument_errors():
    """Host-only parts of the input-pyramid entry points: the fused kernel's envelope, the resize workspace size, and
    argument validation that returns before any CUDA call (so it runs without a GPU)."""
    lib = _lib.load()
    # every x8 / x16 branch from a 256-row source fits the fused kernel, with and without blur
    for (ih, oh) in [(256, 256), (256, 128), (256, 64), (256, 32), (32, 256), (32, 128), (32, 64), (128, 128), (16, 128)]:
        assert lib.tgsr_pyramid_supported(ih, ih, oh, oh, 1) == 1 and lib.tgsr_pyramid_supported(ih, ih, oh, oh, 0) == 1
    assert lib.tgsr_pyramid_supported(100, 100, 50, 50, 1) == 0          # width not a multiple of 4 -> generic kernels
    assert lib.tgsr_pyramid_supported(0, 256, 32, 32, 1) == 0
    # workspace: two tap tables (ksize + 2 ints per output sample) + the horizontally resized intermediate
    assert lib.tgsr_resize_u8_ws_bytes(2, 256, 256, 32, 32, 3) >= 2 * 32 * (17 + 2) * 4 + 2 * 256 * 32 * 3
    assert lib.tgsr_resize_u8_ws_bytes(0, 256, 256, 32, 32, 3) == 0
    err_null, err_shape = -3, -1
    assert lib.tgsr_error_string(err_null) and lib.tgsr_error_string(err_shape)
    assert lib.tgsr_resize_u8(None, None, None, 1, 8, 8, 4, 4, 3, None) != 0
    assert lib.tgsr_gaussian_blur_u8(None, None, None, 1, 8, 8, 3, 2.0, 3, None) != 0
    assert lib.tgsr_normalize_u8_f32(None, None, 1, 8, 8, 3, None) != 0
    assert lib.tgsr_attn_word_stats_f32(None, None, None, None, 1, 4, 16, 0.1, 3, None) != 0
    assert lib.tgsr_kl_loss_f32(None, None, None, 4, None) != 0


def test_pyramid_module_has_no_cpu_path():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from tgsr_b200 import pyramid
    with pytest.raises(RuntimeError, match='CUDA'):
        pyramid.get_imgs_blur(torch.zeros(1, 64, 64, 3, dtype=torch.uint8), [16, 32, 64])
    with pytest.raises(RuntimeError, match='CUDA'):
        pyramid.resize(torch.zeros(1, 64, 64, 3, dtype=torch.uint8), 32)


def test_host_side_size_helpers_agree_with_the_library():
    """Pure host functions of the C ABI (no device work): workspace / operand-image sizes the Python side computes itself must
    equal what the library reports."""
    from tgsr_b200 import build, fast
    build.build()
    lib = _lib.load()
    for B in (1, 7, 32):
        for T in (1, 18, 24, 25, 32, 33, 48):
            assert fast._sfrag_words(B, T) * 4 == lib.tgsr_attn_sfrag_bytes_t(B, T)
        assert lib.tgsr_attn_sfrag_bytes(B) == lib.tgsr_attn_sfrag_bytes_t(B, 24)
    # packed conv image = tap blocks + shift block + ones block (conv_tc.cu TcCfg::WBYTES / 2)
    assert lib.tgsr_packed_conv3x3_elems(128, 64, 0) == 9 * 128 * 64 + 128 * 16 + 128 * 16
    assert lib.tgsr_packed_conv3x3_elems(64, 64, 1) == 16 * 64 * 64 + 4 * 64 * 16 + 128 * 16
    assert fast.attention_blocked_supported(32, 48, 1024) and not fast.attention_blocked_supported(32, 49, 1024)
    assert not fast.attention_blocked_supported(64, 18, 1024) and not fast.attention_blocked_supported(32, 18, 51)
