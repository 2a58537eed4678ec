"""Parity at the bench configurations themselves (BASELINE configs 3 and 4 at the per-GPU batch of 8), the CUDA-graph
runner, and kbnet_intrinsics_inverse against torch.inverse."""
import pytest
import torch

from util import assert_close
from oracle import kbnet_oracle as K
import kbnet_b200

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = torch.device('cuda:0')
F_IMG, F_DEP, F_DEC = [48, 96, 192, 384, 384], [16, 32, 64, 128, 128], [256, 128, 128, 64, 12]


def _model_and_ref(c, n, seed):
    gen = torch.Generator().manual_seed(seed)
    args = K.synth_model_inputs(gen, n, c)
    n_pool = len(c['min_pools']) + len(c['max_pools'])
    ws, we, wd = K.make_model_weights(torch.Generator().manual_seed(0), 2, n_pool)
    ref = K.model_forward(*args, ws, we, wd, c['min_pools'], c['max_pools'], c['min_pred'], c['max_pred'])
    m = kbnet_b200.KBNetModel(3, 2, c['min_pools'], c['max_pools'], 3, 8, F_IMG, F_DEP, [0, 1, 2, 3], F_DEC,
                              min_predict_depth=c['min_pred'], max_predict_depth=c['max_pred'], device=DEV)
    m.load_state_dicts(ws, we, wd)
    return m, [t.to(DEV) for t in args], ref


@pytest.mark.parametrize('cfg', ['KITTI', 'VOID'])
def test_bench_configuration_n8_vs_oracle(cfg):
    """the exact per-GPU workload bench.py times: 8 frames, full resolution, autotuned tiles, eager and CUDA-graph replay"""
    c = K.KITTI if cfg == 'KITTI' else K.VOID
    m, args, ref = _model_and_ref(c, 8, 21)
    kbnet_b200.set_autotune(True)
    try:
        out = m.forward(*args)
        assert_close(out, ref, what='%s N=8 forward' % cfg)
        runner = m.capture(*args)
        assert runner.launches_per_step >= 10
        got = runner.replay().clone()
        assert torch.equal(got, out), 'graph replay differs from the eager forward'
        # new data through the static buffers
        gen = torch.Generator().manual_seed(22)
        args2 = K.synth_model_inputs(gen, 8, c)
        for dst, src in zip(runner.static_inputs, args2):
            dst.copy_(src.to(DEV))
        got2 = runner.replay().clone()
        assert torch.equal(got2, m.forward(*[t.to(DEV) for t in args2]))
    finally:
        kbnet_b200.set_autotune(False)


def test_intrinsics_inverse_skewed_per_sample_and_singular():
    """kbnet_intrinsics_inverse (general 3x3 inverse in fp64) vs torch.inverse, incl. the level-1 scale of
    src/networks.py:333-352; a singular K must give non-finite values, not a crash"""
    gen = torch.Generator().manual_seed(9)
    n = 6
    k = torch.zeros(n, 3, 3)
    k[:, 0, 0] = 500 + 300 * torch.rand(n, generator=gen)
    k[:, 1, 1] = 500 + 300 * torch.rand(n, generator=gen)
    k[:, 0, 2] = 600 * torch.rand(n, generator=gen)
    k[:, 1, 2] = 200 * torch.rand(n, generator=gen)
    k[:, 0, 1] = torch.rand(n, generator=gen) * 3 - 1.5      # skew
    k[:, 2, 2] = 1.0
    k[3] = torch.rand(3, 3, generator=gen) + torch.eye(3) * 2   # a fully general matrix
    sx, sy = 608 / 1216, 176 / 352
    k0, k1 = kbnet_b200.ops.intrinsics_inverse(k.to(DEV), sx, sy)
    scale = torch.tensor([[sx, 1.0, sx], [1.0, sy, sy], [1.0, 1.0, 1.0]])
    ref0 = torch.inverse(k.double()).float()
    ref1 = torch.inverse((k * scale).double()).float()
    assert_close(k0, ref0, tol=1e-6, what='K^-1')
    assert_close(k1, ref1, tol=1e-6, what='(K*scale)^-1')
    # ... and through the KB layer: coordinates generated in-kernel from K^-1 == the reference's tensor path
    c = K.camera_coordinates(k, 10, 12)
    c2 = torch.matmul(k0.cpu(), K.meshgrid_h(n, 10, 12).reshape(n, 3, -1)).view(n, 3, 10, 12)
    assert_close(c2, c, tol=1e-5, what='coordinates from K^-1')
    ks = k.clone()
    ks[1] = 0.0
    ks[1, 0, 0] = 1.0                                           # singular
    s0, _ = kbnet_b200.ops.intrinsics_inverse(ks.to(DEV), sx, sy)
    assert not torch.isfinite(s0[1]).all()
    assert torch.isfinite(s0[0]).all() and torch.isfinite(s0[2:]).all()
