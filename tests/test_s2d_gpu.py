"""GPU parity of the fused SparseToDensePool kernel: pool pyramid bit-exact, output within 1e-3 (all calls go
through the C ABI via kbnet_b200.SparseToDensePool)."""
import hashlib

import pytest
import torch

from conftest import load_golden
from util import assert_close, eq_nan
from oracle import kbnet_oracle as K
import kbnet_b200

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = 'cuda:0'


def _module(cin, mins, maxs, nconv, sd):
    m = kbnet_b200.SparseToDensePool(cin, mins, maxs, 8, nconv, 'xavier_normal', 'leaky_relu')
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize('idx', range(4))
def test_golden_cases(idx):
    c = load_golden('s2d_cases.pt')[idx]
    m = _module(c['x'].shape[1], c['min_pool_sizes'], c['max_pool_sizes'], c['n_convolution'], c['state_dict'])
    out, pyr = m(c['x'].to(DEV), return_pyramid=True)
    assert eq_nan(pyr.cpu(), c['pyramid']), c['name'] + ': pyramid not bit-exact'
    assert_close(out, c['out'], what=c['name'])


def test_config1_fullsize_pyramid_sha256_and_output():
    g = load_golden('s2d_config1_fullsize.pt')
    gen = torch.Generator().manual_seed(1)
    mask = (torch.rand(1, 1, 352, 1216, generator=gen) < 0.05).float()
    z = (torch.rand(1, 1, 352, 1216, generator=gen) * 79.0 + 1.0) * mask
    m = _module(1, [5, 7, 9, 11, 13], [15, 17], 3, g['state_dict'])
    out, pyr = m(z.to(DEV), return_pyramid=True)
    sha = hashlib.sha256(pyr.cpu().contiguous().numpy().tobytes()).hexdigest()
    assert sha == g['pyramid_sha256']
    assert_close(out[:, :, ::16, ::16], g['out_sample'], what='config1 sample')


@pytest.mark.parametrize('cfg,n,p', [('KITTI', 2, 0.05), ('KITTI', 1, 0.0005), ('VOID', 2, 0.005)])
def test_fullsize_vs_oracle(cfg, n, p):
    cfg = getattr(K, cfg)
    gen = torch.Generator().manual_seed(1)
    depth, valid = K.synth_sparse_depth(gen, n, cfg['h'], cfg['w'], p, cfg['dmin'], cfg['dmax'])
    x = torch.cat([depth, valid], 1)
    w = K.make_s2d_weights(torch.Generator().manual_seed(0), 2, len(cfg['min_pools']) + len(cfg['max_pools']))
    ref, ref_pyr = K.s2d_forward(x, w, cfg['min_pools'], cfg['max_pools'], return_pyramid=True)
    m = _module(2, cfg['min_pools'], cfg['max_pools'], 3, w)
    out, pyr = m(x.to(DEV), return_pyramid=True)
    assert torch.equal(pyr.cpu(), ref_pyr)
    assert_close(out, ref, what='s2d full size')


@pytest.mark.parametrize('shape', [(1, 1, 5, 7), (2, 2, 17, 63), (1, 2, 16, 64), (3, 1, 33, 130), (1, 2, 1, 1)])
def test_ragged_shapes_vs_oracle(shape):
    gen = torch.Generator().manual_seed(5)
    n, cin, h, w_ = shape
    z = (torch.rand(n, 1, h, w_, generator=gen) < 0.2).float() * (torch.rand(n, 1, h, w_, generator=gen) * 50 + 1)
    x = torch.cat([z] + [torch.rand(n, 1, h, w_, generator=gen) for _ in range(cin - 1)], 1)
    mins, maxs = [3, 7, 13], [5, 17]
    w = K.make_s2d_weights(torch.Generator().manual_seed(2), cin, 5, 8, 2)
    ref, ref_pyr = K.s2d_forward(x, w, mins, maxs, return_pyramid=True)
    out, pyr = _module(cin, mins, maxs, 2, w)(x.to(DEV), return_pyramid=True)
    assert torch.equal(pyr.cpu(), ref_pyr)
    assert_close(out, ref, what=str(shape))


def test_dense_and_empty_inputs():
    w = K.make_s2d_weights(torch.Generator().manual_seed(2), 2, 7)
    m = _module(2, [5, 7, 9, 11, 13], [15, 17], 3, w)
    for z in (torch.zeros(1, 1, 40, 80), torch.rand(1, 1, 40, 80) * 90 + 0.5, -torch.rand(1, 1, 40, 80)):
        x = torch.cat([z, (z > 0).float()], 1)
        ref, ref_pyr = K.s2d_forward(x, w, [5, 7, 9, 11, 13], [15, 17], return_pyramid=True)
        out, pyr = m(x.to(DEV), return_pyramid=True)
        assert torch.equal(pyr.cpu(), ref_pyr)
        assert_close(out, ref, what='dense/empty')


def test_empty_batch_and_noncontiguous_input():
    w = K.make_s2d_weights(torch.Generator().manual_seed(2), 2, 2, 8, 1)
    m = _module(2, [3], [5], 1, w)
    assert m(torch.zeros(0, 2, 8, 8, device=DEV)).shape == (0, 8, 8, 8)
    xt = torch.rand(1, 24, 2, 20).permute(0, 2, 1, 3)           # non-contiguous [1,2,24,20]
    assert_close(m(xt.to(DEV)), K.s2d_forward(xt.contiguous(), w, [3], [5]), what='non-contiguous')


def test_idempotent_and_input_unmodified():
    w = K.make_s2d_weights(torch.Generator().manual_seed(2), 2, 7)
    m = _module(2, [5, 7, 9, 11, 13], [15, 17], 3, w)
    x = torch.rand(2, 2, 64, 96, device=DEV)
    x0 = x.clone()
    a, b = m(x), m(x)
    assert torch.equal(a, b) and torch.equal(x, x0)
