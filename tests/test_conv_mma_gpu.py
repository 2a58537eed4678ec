"""Warp-level tensor kernel for the full-resolution convolutions with few channels (csrc/conv_mma.cu, C ABI
kbnet_conv3x3_small_forward) vs torch fp64: every operand packing (C <= 4, C <= 8, C <= 16), filter counts that are not
multiples of 8, partial tiles, odd widths (no TMA: predicated loads), linear / LeakyReLU output, channel-slice input and output."""
import pytest
import torch
import torch.nn.functional as F

from util import assert_close
import kbnet_b200

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = 'cuda:0'

CASES = [
    # name, N, C, H, W, F
    ('c3_f48', 2, 3, 33, 48, 48),
    ('c3_f48_full_tiles', 1, 3, 32, 128, 48),
    ('c1_f8', 2, 1, 16, 64, 8),
    ('c4_f30', 1, 4, 17, 70, 30),
    ('c2_f16_odd_w', 1, 2, 9, 37, 16),
    ('c8_f16', 2, 8, 16, 300, 16),
    ('c8_f16_fullres', 1, 8, 352, 1216, 16),
    ('c5_f12', 2, 5, 21, 66, 12),
    ('c8_f5', 1, 8, 8, 8, 5),
    ('c3_f48_fullres', 1, 3, 352, 1216, 48),
    ('c7_f9_tiny', 3, 7, 1, 1, 9),
    ('c6_f16_w_not_multiple_of_4', 2, 6, 20, 70, 16),
    ('c3_f20_w_odd', 1, 3, 18, 131, 20),
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize('slope', [0.2, None])
def test_small_conv_vs_fp64(case, slope):
    name, n, c, h, w, f = case
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(n, c, h, w, generator=gen)
    wt = torch.randn(f, c, 3, 3, generator=gen) / (9 * c) ** 0.5
    y = F.conv2d(x.double(), wt.double(), padding=1)
    ref = (F.leaky_relu(y, slope) if slope is not None else y).float()
    lib = kbnet_b200._lib.load()
    assert lib.kbnet_conv3x3_small_supported(c, f, 3, 1) == 1
    c0, l0 = lib.kbnet_tma_engine_launch_count(), lib.kbnet_launch_count()
    got = kbnet_b200.ops.conv2d_fused([x.to(DEV)], wt.to(DEV), 1, slope)
    if f > 4:
        assert lib.kbnet_tma_engine_launch_count() == c0 and lib.kbnet_launch_count() == l0 + 1, 'expected the mma.sync kernel'
    assert_close(got, ref, tol=1e-4, what=name)


def test_small_conv_channel_slices_and_untouched_neighbours():
    gen = torch.Generator().manual_seed(5)
    wide = torch.randn(2, 20, 18, 66, generator=gen).to(DEV)
    wt = torch.randn(12, 8, 3, 3, generator=gen) * 0.1
    buf = torch.full((2, 20, 18, 66), 7.0, device=DEV)
    kbnet_b200.ops.conv2d_fused([wide[:, 4:12]], wt.to(DEV), 1, 0.2, out=buf[:, 3:15])
    ref = F.leaky_relu(F.conv2d(wide[:, 4:12].cpu().double(), wt.double(), padding=1), 0.2).float()
    assert_close(buf[:, 3:15], ref, tol=1e-4, what='slices')
    assert bool((buf[:, :3] == 7).all()) and bool((buf[:, 15:] == 7).all())


def test_small_conv_agrees_with_the_tcgen05_engine():
    gen = torch.Generator().manual_seed(6)
    x = torch.rand(1, 8, 64, 192, generator=gen).to(DEV)
    wt = (torch.randn(16, 8, 3, 3, generator=gen) * 0.1).to(DEV)
    a = kbnet_b200.ops.conv2d_fused([x], wt, 1, 0.2)
    b = kbnet_b200.ops.conv2d_fused([x], wt, 1, 0.2, engine='tcgen05')
    assert_close(a, b.cpu(), tol=1e-4, what='mma.sync vs tcgen05')


def test_small_conv_rejects_unsupported_shapes():
    lib = kbnet_b200._lib.load()
    assert lib.kbnet_conv3x3_small_supported(64, 12, 3, 1) == 0
    assert lib.kbnet_conv3x3_small_supported(12, 12, 3, 1) == 0      # two chunk planes: slower than the tcgen05 engine, not offered
    assert lib.kbnet_conv3x3_small_supported(8, 16, 3, 2) == 0
    assert lib.kbnet_conv3x3_small_supported(8, 16, 1, 1) == 0
    assert lib.kbnet_conv3x3_small_forward(None, 0, 1, 64, 8, 8, None, 12, 1, 0.2, None, 0, None) != 0
