"""GPU parity of the callers either side of the path: encoder wiring (incl. the reference's intrinsics-scale quirk),
decoder, output head -- against reference-generated fixtures and the oracle."""
import pytest
import torch

from conftest import load_golden
from util import assert_close
from oracle import kbnet_oracle as K
import kbnet_b200

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = torch.device('cuda:0')


def _small_model(g):
    m = kbnet_b200.KBNetModel(3, 2, g['min_pool_sizes'], g['max_pool_sizes'], 3, 8, g['f_img'], g['f_dep'], [0, 1, 2, 3],
                              g['f_dec'], min_predict_depth=g['min_predict_depth'], max_predict_depth=g['max_predict_depth'],
                              device=DEV)
    m.load_state_dicts(g['s2d_state_dict'], g['enc_state_dict'], g['dec_state_dict'])
    return m


@pytest.mark.parametrize('backend', ['kbnet', 'torch'])
def test_small_model_matches_reference_fixture(backend, monkeypatch):
    import kbnet_b200.model as M
    monkeypatch.setattr(M, 'CONV_BACKEND', backend)
    g = load_golden('model_small.pt')
    m = _small_model(g)
    dev = lambda t: t.to(DEV)
    s2d = m.sparse_to_dense_pool(torch.cat([dev(g['sparse_depth']), dev(g['validity_map'])], 1))
    assert_close(s2d, g['s2d_out'], what='s2d')
    latent, skips = m.encoder(dev(g['image']), dev(g['s2d_out']), dev(g['intrinsics']))
    assert_close(latent, g['latent'], what='latent')
    for i, (a, b) in enumerate(zip(skips, g['skips'])):
        assert_close(a, b, what='skip%d' % i)
    out = m.forward(dev(g['image']), dev(g['sparse_depth']), dev(g['validity_map']), dev(g['intrinsics']))
    assert_close(out, g['output_depth'], what='output_depth')


def test_decoder_and_head_vs_oracle():
    g = load_golden('model_small.pt')
    m = _small_model(g)
    out = m.decoder(g['latent'].to(DEV), [s.to(DEV) for s in g['skips']], g['s2d_out'].shape[-2:])[-1]
    ref = K.decoder_forward(g['latent'], g['skips'], g['s2d_out'].shape[-2:], g['dec_state_dict'])
    assert_close(out, ref, what='decoder')
    head = kbnet_b200.ops.depth_head(out, 1.5, 100.0)
    assert_close(head, 1.5 / (torch.sigmoid(ref) + 1.5 / 100.0), what='head')


def test_fused_conv_upsample_non_integer_ratio_and_concat():
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(2, 5, 7, 9, generator=gen)
    skip = torch.rand(2, 3, 16, 21, generator=gen)
    w = torch.randn(6, 8, 3, 3, generator=gen) * 0.2
    import torch.nn.functional as F
    ref = F.leaky_relu(F.conv2d(torch.cat([F.interpolate(x, size=(16, 21), mode='nearest'), skip], 1), w, padding=1), 0.2)
    got = kbnet_b200.ops.conv2d_fused([(x.to(DEV), 'up'), skip.to(DEV)], w.to(DEV), 1, 0.2)
    assert_close(got, ref, what='upsample+concat conv')
    ref2 = F.conv2d(x, w[:, :5], stride=2, padding=1)
    got2 = kbnet_b200.ops.conv2d_fused([x.to(DEV)], w[:, :5].contiguous().to(DEV), 2, None)
    assert_close(got2, ref2, what='stride-2 linear conv')


def test_batch_slices_equal_full_batch():
    """sharding property: every op is per-sample, so a rank's slice equals the same rows of the full batch"""
    g = load_golden('model_small.pt')
    m = _small_model(g)
    args = [g[k].to(DEV) for k in ('image', 'sparse_depth', 'validity_map', 'intrinsics')]
    full = m.forward(*args)
    for r in range(2):
        part = m.forward(*kbnet_b200.dist.shard_batch(args, r, 2))
        assert torch.equal(part, full[r:r + 1])


@pytest.mark.parametrize('cfg', ['KITTI', 'VOID'])
def test_full_size_forward_vs_oracle(cfg):
    """BASELINE configs 3 / 4 at full resolution (one frame): the whole forward -- S2D, four fused KB blocks, decoder with the
    parity up-convolutions, head -- against the CPU oracle; also with measured tile selection switched on."""
    c = K.KITTI if cfg == 'KITTI' else K.VOID
    gen = torch.Generator().manual_seed(5)
    n = 1
    image = torch.rand(n, 3, c['h'], c['w'], generator=gen)
    depth, valid = K.synth_sparse_depth(gen, n, c['h'], c['w'], c['p'], c['dmin'], c['dmax'])
    k = torch.tensor(c['k']).view(1, 3, 3).repeat(n, 1, 1)
    n_pool = len(c['min_pools']) + len(c['max_pools'])
    ws, we, wd = K.make_model_weights(torch.Generator().manual_seed(0), 2, n_pool)
    ref = K.model_forward(image, depth, valid, k, ws, we, wd, c['min_pools'], c['max_pools'], c['min_pred'], c['max_pred'])
    m = kbnet_b200.KBNetModel(3, 2, c['min_pools'], c['max_pools'], 3, 8, [48, 96, 192, 384, 384], [16, 32, 64, 128, 128],
                              [0, 1, 2, 3], [256, 128, 128, 64, 12], min_predict_depth=c['min_pred'],
                              max_predict_depth=c['max_pred'], device=DEV)
    m.load_state_dicts(ws, we, wd)
    args = [t.to(DEV) for t in (image, depth, valid, k)]
    out = m.forward(*args)
    assert_close(out, ref, what='%s full-size forward' % cfg)
    kbnet_b200.set_autotune(True)
    try:
        out2 = m.forward(*args)
    finally:
        kbnet_b200.set_autotune(False)
    assert_close(out2, ref, what='%s full-size forward, autotuned tiles' % cfg)


@pytest.mark.parametrize('shape', [(2, 12, 21, 33, 1), (1, 12, 352, 1216, 1), (2, 7, 16, 130, 3), (1, 32, 9, 64, 4)])
def test_small_direct_conv_and_fused_depth_head(shape):
    """conv_small.cu: exact-fp32 direct 3x3 convolution for <= 4 filters (KBNet's output0), alone and fused with the depth head"""
    import torch.nn.functional as F
    n, c, h, w, f = shape
    gen = torch.Generator().manual_seed(17)
    x = torch.randn(n, c, h, w, generator=gen)
    wt = torch.randn(f, c, 3, 3, generator=gen) * 0.2
    ref = F.conv2d(x.double(), wt.double(), padding=1).float()
    lib = kbnet_b200._lib.load()
    c0 = lib.kbnet_tma_engine_launch_count()
    got = kbnet_b200.ops.conv2d_fused([x.to(DEV)], wt.to(DEV), 1, None)
    assert lib.kbnet_tma_engine_launch_count() == c0, 'expected the CUDA-core kernel for a <= 4-filter layer'
    assert_close(got, ref, tol=1e-5, what='small conv, linear')
    assert_close(kbnet_b200.ops.conv2d_fused([x.to(DEV)], wt.to(DEV), 1, 0.2), F.leaky_relu(ref, 0.2), tol=1e-5, what='small conv, leaky')
    head = kbnet_b200.ops.conv3x3_depth_head(x.to(DEV), wt.to(DEV), 1.5, 100.0)
    assert_close(head, 1.5 / (torch.sigmoid(ref) + 1.5 / 100.0), tol=1e-5, what='fused depth head')
