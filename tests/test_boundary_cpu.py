"""Host-side checks that need no GPU: the C ABI library loads and exports what include/kbnet_b200.h declares, the
module mirrors have the reference's signatures and parameter trees, and there is no CPU fallback."""
import ctypes
import inspect
import os
import re
import sys
import types

import pytest
import torch

from conftest import ROOT, load_golden
import kbnet_b200
from kbnet_b200 import _lib


def _declared_functions():
    src = open(os.path.join(ROOT, 'include', 'kbnet_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(kbnet_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    kbnet_b200.build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_functions()
    assert set(names) == set(_lib.EXPORTS)
    for name in names:
        assert hasattr(lib, name), name
    assert _lib.load().kbnet_version() == 100


def test_library_is_sm100a_and_has_no_torch_dependency():
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert 'sm_100a' in out
    ldd = subprocess.run(['ldd', _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert 'torch' not in ldd and 'c10' not in ldd


def test_bad_arguments_return_error_codes_not_crashes():
    lib = _lib.load()
    rc = lib.kbnet_s2d_forward(None, 1, 2, 8, 8, None, 0, None, 0, None, 1, 8, None, None, 0.2, None, None, None)
    assert rc != 0 and b'null' in lib.kbnet_last_error()
    with pytest.raises(RuntimeError, match='kbnet_b200 .* failed'):
        _lib.check(rc, 'kbnet_s2d_forward')
    # the entry points added around the path: invalid arguments are rejected before any CUDA call
    BADARG = 10001
    assert lib.kbnet_upconv2x_forward(None, 1, 8, 4, 4, 128, None, 8, 1, 0.2, None, 512, None) == BADARG
    assert lib.kbnet_outlier_removal_forward(None, None, 1, 8, 8, 7, 1.5, None, None, None, None) == BADARG
    assert b'outlier_removal' in lib.kbnet_last_error()
    assert lib.kbnet_conv3x3_depth_head_forward(None, 0, 1, 12, 8, 8, None, 1, 1.5, 100.0, None, 0, None) == BADARG
    assert lib.kbnet_conv2d_forward(None, 1, 1, 8, 8, None, None, 8, 3, 1, 1, 0.2, None, 512, None) == BADARG
    assert lib.kbnet_kb_forward(None, 0, None, 0, None, 0, None, None, 1, 16, 16, 0, 8, 8, None, 16, None, 16, None, None, 16, None, None, None,
                                0.2, None, 0, None, 0, None, 0, None, 0, None) == BADARG
    assert lib.kbnet_intrinsics_inverse(None, 1, 0.5, 0.5, None, None, None) == BADARG
    # empty batches are a no-op everywhere
    assert lib.kbnet_upconv2x_forward(None, 0, 8, 4, 4, 128, None, 8, 1, 0.2, None, 512, None) == 0
    assert lib.kbnet_outlier_removal_forward(None, None, 0, 8, 8, 7, 1.5, None, None, None, None) == 0
    lib.kbnet_set_autotune(1)
    lib.kbnet_set_autotune(0)


def test_operand_format_host_side_contract():
    """the host-only parts of the operand-format side channel (include/kbnet_b200.h): sizes, shape support, argument checks"""
    import ctypes
    lib = _lib.load()
    BADARG, ALIGN = 10001, 10003
    # bytes of ONE half: [N][C/8][4 parity planes][H/2][W/2] x 16 B -- the same bytes per value as fp32 halved (2 of 4 bytes)
    assert lib.kbnet_operand_format_bytes(8, 48, 352, 1216, 0) == 8 * 48 * 352 * 1216 * 2
    assert lib.kbnet_operand_format_bytes(8, 48, 352, 1216, 1) == 8 * 48 * 352 * 1216 * 2 // 4     # even pixels only
    assert lib.kbnet_operand_format_bytes(1, 12, 8, 8, 0) == 0                                     # C % 8 != 0
    assert lib.kbnet_operand_format_bytes(1, 8, 7, 8, 0) == 0 and lib.kbnet_operand_format_bytes(1, 8, 8, 6 + 1, 0) == 0   # odd H / W
    # which block shapes run as the fused one-launch program (= may use the side channel): the four KITTI / VOID levels do
    for (ci, cd, cf, fi, fd, ff), (h, w) in zip([(48, 16, 0, 48, 16, 48), (48, 16, 48, 96, 32, 96), (96, 32, 96, 192, 64, 192),
                                                 (192, 64, 192, 384, 128, 384)], [(352, 1216), (176, 608), (88, 304), (44, 152)]):
        assert lib.kbnet_kb_operand_format_supported(ci, cd, cf, h, w, fi, fd, ff) == 1
        assert lib.kbnet_kb_operand_format_supported(ci, cd, cf, h * 480 // 352 // 2 * 2, w * 640 // 1216 // 4 * 4, fi, fd, ff) == 1
    assert lib.kbnet_kb_operand_format_supported(48, 8, 0, 352, 1216, 48, 16, 48) == 0     # C_d % 16 != 0: unfused launches
    assert lib.kbnet_kb_operand_format_supported(48, 16, 0, 351, 1216, 48, 16, 48) == 0    # odd height: no equal parity planes
    assert lib.kbnet_kb_operand_format_supported(48, 16, 0, 352, 1214, 48, 16, 48) == 0    # W % 4 != 0: no TMA
    # argument checks happen before any CUDA call
    assert lib.kbnet_to_operand_format(None, 0, 1, 8, 8, 8, None, None, 0, None) == BADARG
    assert lib.kbnet_from_operand_format(None, None, 1, 8, 8, 8, None, 0, None) == BADARG
    assert lib.kbnet_to_operand_format(None, 0, 0, 8, 8, 8, None, None, 0, None) == 0      # empty batch
    assert lib.kbnet_conv3x3_small_forward_of(None, 0, 1, 3, 8, 8, None, 48, 1, 0.2, None, 0, None, None, None) == BADARG
    of = _lib.KBNetKbOperandFormat()
    of.image_hi = 256                                                                       # hi without lo
    args = [256, 16 * 64, 256, 16 * 64, None, 0, None, 256, 1, 16, 16, 0, 8, 8, 256, 16, 256, 16, 256, 256, 16, None, None, None, 0.2,
            None, 0, 256, 16 * 16, 256, 16 * 16, None, 0]
    assert lib.kbnet_kb_forward_of(*args, ctypes.byref(of), None) == BADARG and b'hi and its lo' in lib.kbnet_last_error()
    of.image_lo = 264                                                                       # 8-byte aligned only
    assert lib.kbnet_kb_forward_of(*args, ctypes.byref(of), None) == ALIGN


def _ref_signature_strings():
    # reference src/networks.py:2100-2107, src/net_utils.py:1297-1308, src/kbnet_model.py:63-82
    return {
        'SparseToDensePool': ['input_channels', 'min_pool_sizes', 'max_pool_sizes', 'n_filter', 'n_convolution',
                              'weight_initializer', 'activation_func'],
        'CalibratedBackprojectionBlock': ['in_channels_image', 'in_channels_depth', 'in_channels_fused', 'n_filter_image',
                                          'n_filter_depth', 'n_filter_fused', 'n_convolution_image', 'n_convolution_depth',
                                          'n_convolution_fused', 'weight_initializer', 'activation_func'],
    }


def test_constructor_signatures_match_reference():
    for name, params in _ref_signature_strings().items():
        sig = inspect.signature(getattr(kbnet_b200, name).__init__)
        assert list(sig.parameters)[1:] == params, name
    s = inspect.signature(kbnet_b200.SparseToDensePool.__init__).parameters
    assert s['min_pool_sizes'].default == [3, 5, 7, 9] and s['n_filter'].default == 8 and s['n_convolution'].default == 3
    assert s['activation_func'].default == 'leaky_relu' and s['weight_initializer'].default == 'kaiming_uniform'
    k = inspect.signature(kbnet_b200.CalibratedBackprojectionBlock.__init__).parameters
    assert (k['n_filter_image'].default, k['n_filter_depth'].default, k['n_filter_fused'].default) == (48, 16, 48)
    f = list(inspect.signature(kbnet_b200.CalibratedBackprojectionBlock.forward).parameters)
    assert f[1:5] == ['image', 'depth', 'coordinates', 'fused']


def test_state_dict_trees_match_reference_fixtures():
    for c in load_golden('s2d_cases.pt'):
        m = kbnet_b200.SparseToDensePool(c['x'].shape[1], c['min_pool_sizes'], c['max_pool_sizes'], 8, c['n_convolution'],
                                         'xavier_normal', 'leaky_relu')
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in c['state_dict'].items()}
        m.load_state_dict(c['state_dict'], strict=True)
    for c in load_golden('kb_cases.pt'):
        ci, cd = c['image'].shape[1], c['depth'].shape[1]
        cf = c['fused'].shape[1] if c['fused'] is not None else 0
        fi, fd, ff = c['conv_image'].shape[1], c['conv_depth'].shape[1], c['conv_fused'].shape[1]
        m = kbnet_b200.CalibratedBackprojectionBlock(ci, cd, ci + cf, fi, fd, ff, 1, 1, 1, 'xavier_normal',
                                                     torch.nn.LeakyReLU(0.2, inplace=True))
        m.load_state_dict(c['state_dict'], strict=True)
        assert m.negative_slope == pytest.approx(0.2)
    g = load_golden('model_small.pt')
    model = kbnet_b200.KBNetModel(3, 2, g['min_pool_sizes'], g['max_pool_sizes'], 3, 8, g['f_img'], g['f_dep'], [0, 1, 2, 3],
                                  g['f_dec'], device=torch.device('cpu'))
    # the fixture keeps the DataParallel 'module.' prefix list too: strict load must work from reference checkpoints
    assert all(k.startswith('module.') for k in g['raw_state_dict_keys'])
    model.load_state_dicts({'module.' + k: v for k, v in g['s2d_state_dict'].items()}, g['enc_state_dict'], g['dec_state_dict'])


def test_full_size_parameter_count_is_the_reference_one():
    model = kbnet_b200.KBNetModel(3, 2, [5, 7, 9, 11, 13], [15, 17], 3, 8, [48, 96, 192, 384, 384], [16, 32, 64, 128, 128],
                                  [0, 1, 2, 3], [256, 128, 128, 64, 12], device=torch.device('cpu'))
    assert sum(p.numel() for p in model.parameters()) == 6957780


def test_no_cpu_fallback():
    c = load_golden('s2d_cases.pt')[0]
    m = kbnet_b200.SparseToDensePool(2, c['min_pool_sizes'], c['max_pool_sizes'], 8, 3, 'xavier_normal', 'leaky_relu')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(c['x'])
    k = load_golden('kb_cases.pt')[0]
    blk = kbnet_b200.CalibratedBackprojectionBlock(4, 2, 8, 6, 4, 6, 1, 1, 1, 'xavier_normal', torch.nn.LeakyReLU(0.2))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        blk(k['image'], k['depth'], k['coordinates'], k['fused'])


def test_activation_and_initializer_errors_follow_reference():
    with pytest.raises(ValueError, match='Unsupported activation function'):
        kbnet_b200.SparseToDensePool(2, activation_func='swish')
    with pytest.raises(ValueError, match='Unsupported weight initializer'):
        kbnet_b200.SparseToDensePool(2, weight_initializer='orthogonal')
    m = kbnet_b200.SparseToDensePool(2, [1, 3, 5], [1, 7])           # sizes <= 1 dropped, src/networks.py:2112-2118
    assert m.min_pool_sizes == [3, 5] and m.max_pool_sizes == [7] and m.pool_convs[0].conv.weight.shape[1] == 3


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'calibrated-backprojection-network_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in txt.replace('the oracle', '').replace('oracle/', ''), os.path.join(dirpath, f)


@pytest.mark.skipif(not os.path.isdir('/root/reference/src'), reason='needs the reference checkout (build container only)')
def test_install_drops_into_reference_kbnet_model():
    """reference kbnet_model.KBNetModel, unchanged, builds the B200 layers after install() and strict-loads a
    checkpoint produced by the unpatched reference; forward then refuses to run on CPU (no fallback)."""
    sys.path.insert(0, '/root/reference/src')
    for m in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    import networks, net_utils, kbnet_model  # noqa: E401
    args = (3, 2, [5, 7, 9], [11, 13], 3, 8, [8, 16, 16, 32, 32], [4, 8, 8, 16, 16], [0, 1, 2, 3], [32, 16, 16, 8, 4])
    kw = dict(deconv_type='up', weight_initializer='xavier_normal', activation_func='leaky_relu', device=torch.device('cpu'))
    ref = kbnet_model.KBNetModel(*args, **kw)
    orig = kbnet_b200.install(networks, net_utils)
    try:
        new = kbnet_model.KBNetModel(*args, **kw)
        assert isinstance(new.sparse_to_dense_pool.module, kbnet_b200.SparseToDensePool)
        assert isinstance(new.encoder.module.calibrated_backprojection3, kbnet_b200.CalibratedBackprojectionBlock)
        new.sparse_to_dense_pool.load_state_dict(ref.sparse_to_dense_pool.state_dict(), strict=True)
        new.encoder.load_state_dict(ref.encoder.state_dict(), strict=True)
        new.decoder.load_state_dict(ref.decoder.state_dict(), strict=True)
        new.eval()
        g = load_golden('model_small.pt')
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            new.forward(g['image'], g['sparse_depth'], g['validity_map'], g['intrinsics'])
    finally:
        networks.SparseToDensePool, net_utils.CalibratedBackprojectionBlock = orig
