"""tcgen05 conv engine (bf16x3 split) vs the exact-fp32 CUDA-core path of the same library and vs torch fp64 on the
same inputs -- every geometry the hot path uses: 3x3 s1, 3x3 s2 (parity planes), 1x1 s2, multi-source concat,
nearest upsample, ragged sizes, F not a multiple of 16, F > 256 (two N parts)."""
import pytest
import torch
import torch.nn.functional as F

from util import assert_close
import kbnet_b200

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = 'cuda:0'


@pytest.fixture(autouse=True)
def _restore_precision():
    yield
    kbnet_b200.set_precision('bf16x3')


def _ref(srcs, w, stride, slope, size=None):
    for s in srcs:
        if not isinstance(s, tuple):
            size = tuple(s.shape[-2:])
    parts = [F.interpolate(s[0], size=size, mode='nearest') if isinstance(s, tuple) else s for s in srcs]
    y = F.conv2d(torch.cat(parts, 1).double(), w.double(), stride=stride, padding=w.shape[-1] // 2)
    return (F.leaky_relu(y, slope) if slope is not None else y).float()


CASES = [
    # name, N, [C...], H, W, F, k, stride
    ('s1_small', 1, [16], 12, 20, 16, 3, 1),
    ('s1_c3_f48', 2, [3], 33, 47, 48, 3, 1),
    ('s1_wide', 1, [8], 16, 300, 16, 3, 1),
    ('s1_two_src', 2, [20, 13], 19, 37, 24, 3, 1),
    ('s1_f12', 1, [64], 24, 40, 12, 3, 1),
    ('s1_f1', 1, [12], 21, 33, 1, 3, 1),
    ('s1_c200_f256', 1, [200], 11, 38, 256, 3, 1),
    ('s2_small', 1, [16], 12, 20, 16, 3, 2),
    ('s2_odd', 2, [19], 13, 19, 32, 3, 2),
    ('s2_c48_f96', 2, [48], 44, 76, 96, 3, 2),
    ('s2_f384', 1, [40], 22, 38, 384, 3, 2),
    ('k1s2', 2, [48, 3, 48], 22, 38, 96, 1, 2),
    ('k1s2_odd', 1, [7], 13, 19, 20, 1, 2),
    ('k1s1', 1, [24], 9, 15, 8, 1, 1),
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_umma_vs_fp32_paths(case):
    name, n, chans, h, w, f, k, stride = case
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    srcs = [torch.randn(n, c, h, w, generator=gen) for c in chans]
    wt = torch.randn(f, sum(chans), k, k, generator=gen) / (sum(chans) * k * k) ** 0.5
    ref = _ref(srcs, wt, stride, 0.2)
    dsrcs = [s.to(DEV) for s in srcs]
    kbnet_b200.set_precision('fp32')
    exact = kbnet_b200.ops.conv2d_fused(dsrcs, wt.to(DEV), stride, 0.2)
    assert_close(exact, ref, tol=1e-5, what=name + ' fp32 path')
    kbnet_b200.set_precision('bf16x3')
    got = kbnet_b200.ops.conv2d_fused(dsrcs, wt.to(DEV), stride, 0.2)
    assert_close(got, ref, tol=1e-4, what=name + ' tensor-core path')


def test_umma_upsample_concat_and_linear_output():
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(2, 24, 7, 9, generator=gen)
    skip = torch.randn(2, 10, 14, 18, generator=gen)
    w1 = torch.randn(16, 24, 3, 3, generator=gen) * 0.1
    w2 = torch.randn(20, 26, 3, 3, generator=gen) * 0.1
    d_ref = _ref([(x,)], w1, 1, 0.2, size=(14, 18))
    d = kbnet_b200.ops.conv2d_fused([(x.to(DEV), 'up', (14, 18))], w1.to(DEV), 1, 0.2)
    assert_close(d, d_ref, tol=1e-4, what='upsample conv')
    o_ref = _ref([d_ref, skip], w2, 1, None)
    o = kbnet_b200.ops.conv2d_fused([d, skip.to(DEV)], w2.to(DEV), 1, None)
    assert_close(o, o_ref, tol=1e-4, what='concat conv, linear output')


def test_umma_output_into_channel_slice_and_weight_repack_on_update():
    gen = torch.Generator().manual_seed(12)
    x = torch.randn(2, 16, 10, 14, generator=gen).to(DEV)
    w = (torch.randn(8, 16, 3, 3, generator=gen) * 0.1).to(DEV)
    buf = torch.full((2, 20, 10, 14), 5.0, device=DEV)
    kbnet_b200.ops.conv2d_fused([x], w, 1, 0.2, out=buf[:, 4:12])
    assert_close(buf[:, 4:12], _ref([x.cpu()], w.cpu(), 1, 0.2), tol=1e-4, what='slice output')
    assert bool((buf[:, :4] == 5).all()) and bool((buf[:, 12:] == 5).all())
    w.mul_(2.0)                                    # in-place update bumps the version counter -> re-pack
    y = kbnet_b200.ops.conv2d_fused([x], w, 1, 0.2)
    assert_close(y, _ref([x.cpu()], w.cpu(), 1, 0.2), tol=1e-4, what='after weight update')


def test_bf16x3_is_fp32_grade_on_large_k():
    """K = 9 * 192 = 1728 (KB level 4 conv_image): error stays ~1e-5, two orders inside the parity budget"""
    gen = torch.Generator().manual_seed(13)
    x = torch.rand(1, 192, 44, 76, generator=gen)
    w = torch.randn(384, 192, 3, 3, generator=gen) * (2.0 / (9 * 192 + 9 * 384)) ** 0.5
    ref = _ref([x], w, 2, 0.2)
    got = kbnet_b200.ops.conv2d_fused([x.to(DEV)], w.to(DEV), 2, 0.2)
    assert_close(got, ref, tol=1e-4, what='K=1728')
