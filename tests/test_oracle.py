"""Pin the oracle: restatement vs fixtures produced by the unmodified reference
(tests/golden/make_golden.py), and vs the live reference when it is present."""
import hashlib
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import kbnet_oracle as K
from oracle.pools_np import pool_pyramid_np

torch.set_grad_enabled(False)


def eq_nan(a, b):
    """bit-exact on non-NaN values, identical NaN positions"""
    return torch.equal(a.isnan(), b.isnan()) and torch.equal(torch.nan_to_num(a, nan=0.0), torch.nan_to_num(b, nan=0.0))


@pytest.mark.parametrize('idx', range(4))
def test_s2d_oracle_matches_reference_fixture(idx):
    c = load_golden('s2d_cases.pt')[idx]
    out, pyr = K.s2d_forward(c['x'], c['state_dict'], c['min_pool_sizes'], c['max_pool_sizes'], return_pyramid=True)
    assert eq_nan(pyr, c['pyramid']), c['name']
    m = ~c['out'].isnan()
    assert torch.equal(out.isnan(), c['out'].isnan())
    assert (out[m] - c['out'][m]).abs().max() <= 1e-6 * max(1.0, float(c['out'][m].abs().max()))


@pytest.mark.parametrize('idx', range(4))
def test_numpy_pool_pyramid_bit_exact(idx):
    c = load_golden('s2d_cases.pt')[idx]
    pyr = pool_pyramid_np(c['x'][:, 0].numpy(), c['min_pool_sizes'], c['max_pool_sizes'])
    assert eq_nan(torch.from_numpy(pyr), c['pyramid']), c['name']


def test_s2d_quirks_are_in_the_fixture():
    c = load_golden('s2d_cases.pt')[0]
    pyr, x = c['pyramid'], c['x']
    assert pyr[0, 0, 20, 30] != 999          # exactly-999 point is erased
    assert pyr[0, 0, 30, 40] == -2.0         # negative participates in the min
    assert pyr[1].isnan().any()              # NaN propagated
    assert x[0, 0, 10, 10] == 1000.0 and pyr[0, 0, 10, 10] == 0.0   # >= 999 loses to empties in a 3x3 of empties


@pytest.mark.parametrize('idx', range(3))
def test_kb_oracle_matches_reference_fixture(idx):
    c = load_golden('kb_cases.pt')[idx]
    coords = K.camera_coordinates(c['k'], *c['image'].shape[-2:])
    assert (coords - c['coordinates']).abs().max() <= 1e-6
    o = K.kb_forward(c['image'], c['depth'], c['coordinates'], c['fused'], c['state_dict'])
    for got, ref in zip(o, (c['conv_image'], c['conv_depth'], c['conv_fused'])):
        assert got.shape == ref.shape
        assert (got - ref).abs().max() <= 1e-6 * max(1.0, float(ref.abs().max()))


def test_kb_even_pixel_shortcut_is_equivalent():
    """proj_depth / xyz / conv_fused only matter at even pixels (1x1, stride 2)."""
    c = load_golden('kb_cases.pt')[2]
    sd = c['state_dict']
    ev = lambda t: t[:, :, ::2, ::2]
    z = K.conv2d(ev(c['depth']), sd['proj_depth.conv.weight'])
    cat = torch.cat([ev(c['image']), ev(c['coordinates']) * z, ev(c['fused'])], 1)
    got = K.conv2d(cat, sd['conv_fused.conv.weight'])
    assert (got - c['conv_fused']).abs().max() <= 1e-6


def test_model_oracle_matches_reference_fixture():
    g = load_golden('model_small.pt')
    x = torch.cat([g['sparse_depth'], g['validity_map']], 1)
    s2d = K.s2d_forward(x, g['s2d_state_dict'], g['min_pool_sizes'], g['max_pool_sizes'])
    assert (s2d - g['s2d_out']).abs().max() <= 1e-5
    latent, skips, coords = K.encoder_forward(g['image'], g['s2d_out'], g['intrinsics'], g['enc_state_dict'], return_coords=True)
    for a, b in zip(coords, g['coordinates']):     # includes the level>=2 scale bug
        assert a.shape == b.shape and (a - b).abs().max() <= 1e-5
    assert (latent - g['latent']).abs().max() <= 1e-5
    for a, b in zip(skips, g['skips']):
        assert (a - b).abs().max() <= 1e-5
    out = K.model_forward(g['image'], g['sparse_depth'], g['validity_map'], g['intrinsics'],
                          g['s2d_state_dict'], g['enc_state_dict'], g['dec_state_dict'],
                          g['min_pool_sizes'], g['max_pool_sizes'], g['min_predict_depth'], g['max_predict_depth'])
    assert (out - g['output_depth']).abs().max() <= 1e-4 * float(g['output_depth'].abs().max())


def test_fullsize_config1_pyramid_sha256():
    g = load_golden('s2d_config1_fullsize.pt')
    gen = torch.Generator().manual_seed(1)
    mask = (torch.rand(1, 1, 352, 1216, generator=gen) < 0.05).float()
    z = (torch.rand(1, 1, 352, 1216, generator=gen) * 79.0 + 1.0) * mask
    sha = lambda t: hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()
    assert sha(z) == g['x_sha256']
    pyr = K.s2d_pool_pyramid(z, [5, 7, 9, 11, 13], [15, 17])
    assert sha(pyr) == g['pyramid_sha256']
    out = K.s2d_forward(z, g['state_dict'], [5, 7, 9, 11, 13], [15, 17])
    assert (out[:, :, ::16, ::16] - g['out_sample']).abs().max() <= 1e-5 * g['out_absmax']


def test_weight_factories_have_reference_shapes():
    g = load_golden('model_small.pt')
    gen = torch.Generator().manual_seed(0)
    s, e, d = K.make_model_weights(gen, 2, 5, 8, 3, g['f_img'], g['f_dep'], g['f_dec'])
    for mine, ref in ((s, g['s2d_state_dict']), (e, g['enc_state_dict']), (d, g['dec_state_dict'])):
        assert set(mine) == set(ref)
        for k in mine:
            assert tuple(mine[k].shape) == tuple(ref[k].shape), k


@pytest.mark.skipif(not os.path.isdir('/root/reference/src'), reason='live reference only exists in the build container')
def test_oracle_vs_live_reference_random_shapes():
    sys.path.insert(0, '/root/reference/src')
    for m in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    import networks, net_utils  # noqa: E401
    gen = torch.Generator().manual_seed(7)
    torch.manual_seed(3)
    mod = networks.SparseToDensePool(2, [3, 9], [5, 21], 8, 2, 'xavier_normal', 'leaky_relu').eval()
    z = (torch.rand(2, 1, 33, 41, generator=gen) < 0.1).float() * torch.rand(2, 1, 33, 41, generator=gen) * 50
    x = torch.cat([z, (z > 0).float()], 1)
    ref = mod(x)
    got = K.s2d_forward(x, mod.state_dict(), [3, 9], [5, 21])
    assert (ref - got).abs().max() <= 1e-6 * float(ref.abs().max())
    blk = net_utils.CalibratedBackprojectionBlock(5, 3, 9, 7, 6, 5, 1, 1, 1, 'xavier_normal', torch.nn.LeakyReLU(0.2)).eval()
    im, de, fu = torch.rand(2, 5, 11, 16, generator=gen), torch.rand(2, 3, 11, 16, generator=gen), torch.rand(2, 4, 11, 16, generator=gen)
    k = torch.tensor([[20.0, 0, 8], [0, 21.0, 5], [0, 0, 1]]).view(1, 3, 3).repeat(2, 1, 1)
    co = K.camera_coordinates(k, 11, 16)
    for a, b in zip(blk(im, de, co, fu), K.kb_forward(im, de, co, fu, blk.state_dict())):
        assert (a - b).abs().max() <= 1e-6


@pytest.mark.parametrize('idx', range(4))
def test_outlier_removal_oracle_matches_reference_fixture(idx):
    """OutlierRemoval.remove_outliers (src/net_utils.py:1761-1806): the oracle restatement is bit-exact against outputs of
    the reference itself (tests/golden/make_golden.py::write_outlier_cases)."""
    c = load_golden('outlier_cases.pt')[idx]
    d, v = K.outlier_removal(c['sparse_depth'], c['validity_map'], c['kernel_size'], c['threshold'])
    assert torch.equal(d, c['out_depth']) and torch.equal(v, c['out_validity']), c['name']
    d2, v2 = K.outlier_removal(c['sparse_depth'], None, c['kernel_size'], c['threshold'])
    assert torch.equal(d2, c['out_depth']) and torch.equal(v2, c['out_validity'])
