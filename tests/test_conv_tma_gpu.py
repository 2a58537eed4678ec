"""Persistent TMA-fed tcgen05 engine (csrc/conv_tma.cu) vs torch fp64 on the same inputs: every geometry (3x3 s1, 1x1 s1,
3x3 s2 parity planes, 1x1 s2), multi-source concat, exact-2x nearest upsample, filter splits, partial tiles, many tiles per
CTA (persistent loop + TMEM double buffering), output slices -- and the fused KB block program with both coordinate forms.
Every case asserts that the launch really went through this engine (kbnet_tma_engine_launch_count)."""
import pytest
import torch
import torch.nn.functional as F

from util import assert_close
from oracle import kbnet_oracle as K
import kbnet_b200

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = 'cuda:0'


def _ref(srcs, w, stride, slope, size=None):
    for s in srcs:
        if not isinstance(s, tuple):
            size = tuple(s.shape[-2:])
    parts = [F.interpolate(s[0], size=size, mode='nearest') if isinstance(s, tuple) else s for s in srcs]
    y = F.conv2d(torch.cat(parts, 1).double(), w.double(), stride=stride, padding=w.shape[-1] // 2)
    return (F.leaky_relu(y, slope) if slope is not None else y).float()


def _v2_count():
    return kbnet_b200._lib.load().kbnet_tma_engine_launch_count()


CASES = [
    # name, N, [C...], H, W, F, k, stride
    ('s1_small', 2, [16], 12, 20, 16, 3, 1),
    ('s1_c3_f48', 2, [3], 33, 48, 48, 3, 1),
    ('s1_wide', 1, [8], 16, 300, 16, 3, 1),
    ('s1_two_src', 2, [20, 13], 19, 36, 24, 3, 1),
    ('s1_f12', 1, [64], 24, 40, 12, 3, 1),
    ('s1_c200_f256', 1, [200], 11, 40, 256, 3, 1),
    ('s1_many_tiles', 2, [32, 32], 88, 304, 64, 3, 1),
    ('s1_fullres_f12', 1, [12], 352, 1216, 12, 3, 1),
    ('s1_c40_f1', 1, [40], 21, 32, 1, 3, 1),
    ('s2_small', 1, [16], 12, 20, 16, 3, 2),
    ('s2_odd_h', 2, [19], 13, 20, 32, 3, 2),
    ('s2_c48_f96', 2, [48], 44, 76, 96, 3, 2),
    ('s2_f384', 1, [40], 22, 40, 384, 3, 2),
    ('s2_many_tiles', 2, [48], 176, 608, 48, 3, 2),
    ('k1s2', 2, [48, 3, 48], 22, 40, 96, 1, 2),
    ('k1s2_odd', 1, [7], 13, 20, 20, 1, 2),
    ('k1s1', 1, [24], 9, 16, 8, 1, 1),
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_tma_engine_vs_fp64(case):
    name, n, chans, h, w, f, k, stride = case
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    srcs = [torch.randn(n, c, h, w, generator=gen) for c in chans]
    wt = torch.randn(f, sum(chans), k, k, generator=gen) / (sum(chans) * k * k) ** 0.5
    ref = _ref(srcs, wt, stride, 0.2)
    c0 = _v2_count()
    got = kbnet_b200.ops.conv2d_fused([s.to(DEV) for s in srcs], wt.to(DEV), stride, 0.2, engine='tcgen05')
    assert _v2_count() == c0 + 1, 'the TMA engine did not take this shape'
    assert_close(got, ref, tol=1e-4, what=name)


@pytest.mark.parametrize('shape', [(2, 24, 7, 8, 16), (1, 64, 22, 76, 128), (2, 64, 88, 304, 12)])
def test_tma_engine_upsample_2x(shape):
    n, c, h, w, f = shape
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(n, c, h, w, generator=gen)
    wt = torch.randn(f, c, 3, 3, generator=gen) / (9 * c) ** 0.5
    ref = _ref([(x,)], wt, 1, 0.2, size=(2 * h, 2 * w))
    c0 = _v2_count()
    got = kbnet_b200.ops.conv2d_fused([(x.to(DEV), 'up', (2 * h, 2 * w))], wt.to(DEV), 1, 0.2)
    assert _v2_count() == c0 + 1
    assert_close(got, ref, tol=1e-4, what='upsample 2x')


def test_tma_engine_concat_linear_output_and_slice():
    gen = torch.Generator().manual_seed(22)
    d = torch.randn(2, 16, 14, 20, generator=gen)
    skip = torch.randn(2, 10, 14, 20, generator=gen)
    w = torch.randn(20, 26, 3, 3, generator=gen) * 0.1
    buf = torch.full((2, 30, 14, 20), 5.0, device=DEV)
    wide = torch.randn(2, 40, 14, 20, generator=gen).to(DEV)
    wide[:, 8:24] = d.to(DEV)
    c0 = _v2_count()
    kbnet_b200.ops.conv2d_fused([wide[:, 8:24], skip.to(DEV)], w.to(DEV), 1, None, out=buf[:, 4:24])
    assert _v2_count() == c0 + 1
    assert_close(buf[:, 4:24], _ref([d, skip], w, 1, None), tol=1e-4, what='concat, linear, slices')
    assert bool((buf[:, :4] == 5).all()) and bool((buf[:, 24:] == 5).all())


@pytest.mark.parametrize('form', ['kinv', 'coordinates'])
@pytest.mark.parametrize('dims', [(16, 16, 0, 16, 16, 16, 24, 40), (16, 16, 16, 32, 16, 32, 21, 36), (32, 16, 32, 48, 32, 48, 44, 152)])
def test_fused_kb_program_vs_oracle(form, dims):
    ci, cd, cf, fi, fd, ff, h, w = dims
    gen = torch.Generator().manual_seed(31)
    n = 2
    image = torch.rand(n, ci, h, w, generator=gen)
    depth = torch.rand(n, cd, h, w, generator=gen)
    fused = torch.rand(n, cf, h, w, generator=gen) if cf else None
    k = torch.tensor([[300.0, 0.0, w / 2.0], [0.0, 310.0, h / 2.0], [0.0, 0.0, 1.0]]).view(1, 3, 3).repeat(n, 1, 1)
    k[1, 0, 2] += 3.0
    coords = K.camera_coordinates(k, h, w)
    wts = K.make_kb_weights(torch.Generator().manual_seed(0), ci, cd, ci + cf, fi, fd, ff)
    ref = K.kb_forward(image, depth, coords, fused, wts)
    m = kbnet_b200.CalibratedBackprojectionBlock(ci, cd, ci + cf, fi, fd, ff, 1, 1, 1, 'xavier_normal', torch.nn.LeakyReLU(0.2))
    m.load_state_dict(wts, strict=True)
    m.to(DEV)
    c0 = _v2_count()
    got = m(image=image.to(DEV), depth=depth.to(DEV), coordinates=(torch.inverse(k) if form == 'kinv' else coords).to(DEV),
            fused=fused.to(DEV) if fused is not None else None)
    assert _v2_count() == c0 + 1, 'the KB block did not run as one fused launch'
    for g, r, name in zip(got, ref, ('conv_image', 'conv_depth', 'conv_fused')):
        assert_close(g, r, what='fused KB %s/%s' % (form, name))


@pytest.mark.parametrize('shape', [(2, 24, 7, 8, 16), (1, 64, 22, 76, 128), (2, 64, 88, 304, 12), (1, 128, 44, 152, 64), (1, 20, 5, 12, 256)])
def test_upconv2x_as_four_parity_convs(shape):
    """nearest-2x upsample + 3x3 conv computed as four 2x2 convolutions at the source resolution (kbnet_upconv2x_forward)"""
    n, c, h, w, f = shape
    gen = torch.Generator().manual_seed(41)
    x = torch.randn(n, c, h, w, generator=gen)
    wt = torch.randn(f, c, 3, 3, generator=gen) / (9 * c) ** 0.5
    for slope in (0.2, None):
        ref = _ref([(x,)], wt, 1, slope, size=(2 * h, 2 * w))
        c0 = _v2_count()
        got = kbnet_b200.ops.upconv2x(x.to(DEV), wt.to(DEV), slope)
        assert got is not None and _v2_count() == c0 + 1
        assert_close(got, ref, tol=1e-4, what='upconv2x')
