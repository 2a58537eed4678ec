"""GPU parity of the Calibrated Backprojection layer (through the C ABI via kbnet_b200.CalibratedBackprojectionBlock)."""
import pytest
import torch

from conftest import load_golden
from util import assert_close
from oracle import kbnet_oracle as K
import kbnet_b200

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = 'cuda:0'


def _block(ci, cd, cf, fi, fd, ff, sd):
    m = kbnet_b200.CalibratedBackprojectionBlock(ci, cd, ci + cf, fi, fd, ff, 1, 1, 1, 'xavier_normal',
                                                 torch.nn.LeakyReLU(negative_slope=0.20, inplace=True))
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval()


def _dims(c):
    ci, cd = c['image'].shape[1], c['depth'].shape[1]
    cf = c['fused'].shape[1] if c['fused'] is not None else 0
    return ci, cd, cf, c['conv_image'].shape[1], c['conv_depth'].shape[1], c['conv_fused'].shape[1]


@pytest.fixture(autouse=True)
def _restore_precision():
    yield
    kbnet_b200.set_precision('bf16x3')


@pytest.mark.parametrize('idx', range(3))
@pytest.mark.parametrize('path', ['coordinates', 'kinv'])
@pytest.mark.parametrize('precision', ['bf16x3', 'fp32'])
def test_golden_cases(idx, path, precision):
    kbnet_b200.set_precision(precision)
    c = load_golden('kb_cases.pt')[idx]
    m = _block(*_dims(c), c['state_dict'])
    coords = c['coordinates'].to(DEV) if path == 'coordinates' else torch.inverse(c['k']).to(DEV)
    fused = c['fused'].to(DEV) if c['fused'] is not None else None
    got = m(image=c['image'].to(DEV), depth=c['depth'].to(DEV), coordinates=coords, fused=fused)
    for g, name in zip(got, ('conv_image', 'conv_depth', 'conv_fused')):
        assert_close(g, c[name], what='%s/%s/%s' % (c['name'], path, name))


def test_config2_fullsize_vs_oracle():
    """BASELINE config 2: KB block 2, N=8, image/fused 48ch, depth 16ch @176x608, KITTI K scaled by 0.5."""
    gen = torch.Generator().manual_seed(1)
    n, h, w = 8, 176, 608
    image = torch.rand(n, 48, h, w, generator=gen)
    depth = torch.rand(n, 16, h, w, generator=gen)
    fused = torch.rand(n, 48, h, w, generator=gen)
    k = torch.tensor(K.KITTI['k']).view(1, 3, 3).repeat(n, 1, 1)
    k = K.scale_intrinsics_level1(k, 352, 1216, 176, 608)
    coords = K.camera_coordinates(k, h, w)
    wts = K.make_kb_weights(torch.Generator().manual_seed(0), 48, 16, 96, 96, 32, 96)
    ref = K.kb_forward(image, depth, coords, fused, wts)
    m = _block(48, 16, 48, 96, 32, 96, wts)
    got = m(image.to(DEV), depth.to(DEV), torch.inverse(k).to(DEV), fused.to(DEV))
    for g, r, name in zip(got, ref, ('conv_image', 'conv_depth', 'conv_fused')):
        assert g.shape == r.shape
        assert_close(g, r, what='config2/' + name)
    got2 = m(image.to(DEV), depth.to(DEV), coords.to(DEV), fused.to(DEV))
    for g, r, name in zip(got2, ref, ('conv_image', 'conv_depth', 'conv_fused')):
        assert_close(g, r, what='config2/coords-tensor/' + name)


@pytest.mark.parametrize('lvl,n', [(1, 1), (3, 2), (4, 2)])
def test_kitti_levels_vs_oracle(lvl, n):
    """all four KB levels of the KITTI encoder, incl. fused=None at level 1"""
    fi, fd = [48, 96, 192, 384], [16, 32, 64, 128]
    ci, cd = ([48] + fi)[lvl - 1], ([16] + fd)[lvl - 1]
    cf = 0 if lvl == 1 else ci
    h, w = 352 >> (lvl - 1), 1216 >> (lvl - 1)
    gen = torch.Generator().manual_seed(lvl)
    image, depth = torch.rand(n, ci, h, w, generator=gen), torch.rand(n, cd, h, w, generator=gen)
    fused = torch.rand(n, cf, h, w, generator=gen) if cf else None
    k = torch.tensor(K.KITTI['k']).view(1, 3, 3).repeat(n, 1, 1)
    if lvl > 1:
        k = K.scale_intrinsics_level1(k, 352, 1216, 176, 608)
    coords = K.camera_coordinates(k, h, w)
    wts = K.make_kb_weights(torch.Generator().manual_seed(0), ci, cd, ci + cf, fi[lvl - 1], fd[lvl - 1], fi[lvl - 1])
    ref = K.kb_forward(image, depth, coords, fused, wts)
    m = _block(ci, cd, cf, fi[lvl - 1], fd[lvl - 1], fi[lvl - 1], wts)
    got = m(image.to(DEV), depth.to(DEV), torch.inverse(k).to(DEV), fused.to(DEV) if cf else None)
    for g, r, name in zip(got, ref, ('conv_image', 'conv_depth', 'conv_fused')):
        assert_close(g, r, what='L%d/%s' % (lvl, name))


def test_outputs_into_concat_slices_and_strided_inputs():
    c = load_golden('kb_cases.pt')[0]
    ci, cd, cf, fi, fd, ff = _dims(c)
    m = _block(ci, cd, cf, fi, fd, ff, c['state_dict'])
    n, _, h, w = c['image'].shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    skip = torch.full((n, ff + fd, ho, wo), -7.0, device=DEV)
    big = torch.zeros(n, ci + cd + 3, h, w, device=DEV)
    big[:, :ci] = c['image'].to(DEV)
    big[:, ci + 3:] = c['depth'].to(DEV)
    o_i, o_d, o_f = m(big[:, :ci], big[:, ci + 3:], c['coordinates'].to(DEV), c['fused'].to(DEV),
                      out_fused=skip[:, :ff], out_depth=skip[:, ff:])
    assert o_f.data_ptr() == skip.data_ptr()
    assert_close(skip, torch.cat([c['conv_fused'], c['conv_depth']], 1), what='skip concat in place')
    assert_close(o_i, c['conv_image'], what='conv_image')


def test_empty_batch():
    c = load_golden('kb_cases.pt')[0]
    m = _block(*_dims(c), c['state_dict'])
    z = lambda ch: torch.zeros(0, ch, 14, 18, device=DEV)
    o = m(z(4), z(2), z(3), z(4))
    assert [tuple(t.shape) for t in o] == [(0, 6, 7, 9), (0, 4, 7, 9), (0, 6, 7, 9)]
