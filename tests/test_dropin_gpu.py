"""The real drop-in on hardware: the UNMODIFIED reference kbnet_model.KBNetModel (baseline/_ref, installed by
tools/install_reference.py; /root/reference in the build container) with kbnet_b200.install() applied, running on the GPU:
reference encoder wiring -> [N,3,H,W] coordinates path of the KB kernels, torch.nn.DataParallel wrappers
(src/kbnet_model.py:408-415), reference decoder on stock cuDNN.  Compared with the unpatched reference on the CPU."""
import pytest
import torch

from util import assert_close
from oracle import refload
import kbnet_b200

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

MODS = refload.load()
needs_ref = pytest.mark.skipif(MODS is None, reason='no copy of the reference (baseline/_ref) on this box')

SMALL = dict(h=64, w=96, min_pools=[5, 7, 9], max_pools=[11, 13], min_pred=1.5, max_pred=100.0,
             f_img=[16, 32, 32, 64, 64], f_dep=[8, 16, 16, 32, 32], f_dec=[64, 32, 32, 16, 8])
KITTI = dict(h=352, w=1216, min_pools=[5, 7, 9, 11, 13], max_pools=[15, 17], min_pred=1.5, max_pred=100.0,
             f_img=[48, 96, 192, 384, 384], f_dep=[16, 32, 64, 128, 128], f_dec=[256, 128, 128, 64, 12])


def _inputs(n, h, w, seed, skew=False):
    gen = torch.Generator().manual_seed(seed)
    image = torch.rand(n, 3, h, w, generator=gen)
    mask = (torch.rand(n, 1, h, w, generator=gen) < 0.05).float()
    depth = (torch.rand(n, 1, h, w, generator=gen) * 79.0 + 1.0) * mask
    valid = (depth > 0).float()
    k = torch.tensor([[0.6 * w, 0.0, 0.5 * w], [0.0, 0.6 * w, 0.5 * h], [0.0, 0.0, 1.0]]).view(1, 3, 3).repeat(n, 1, 1)
    if skew:
        k[:, 0, 1] = 0.7
        k = k * (1.0 + 0.05 * torch.arange(n).view(n, 1, 1))      # per-sample K
        k[:, 2, 2] = 1.0
    return image, depth, valid, k


def _pair(cfg, device):
    """(unpatched reference model on the CPU, patched reference model on `device`) with identical weights"""
    # The reference calls torch.inverse inside the DataParallel replicas (src/networks.py:328), i.e. from one thread per GPU;
    # torch's lazy linalg initialisation is not thread-safe on first use ("lazy wrapper should be called at most once"), so
    # initialise it on the main thread for every visible device first.  (Stock-PyTorch behaviour, nothing of this library.)
    for d in range(torch.cuda.device_count()):
        torch.inverse(torch.eye(3, device='cuda:%d' % d).unsqueeze(0))
    torch.cuda.synchronize()
    networks, net_utils, kbnet_model, _ = MODS
    ref = refload.build_reference_model(cfg, torch.device('cpu'), cfg['f_img'], cfg['f_dep'], cfg['f_dec'], seed=3)
    orig = kbnet_b200.install(networks, net_utils)
    try:
        new = refload.build_reference_model(cfg, device, cfg['f_img'], cfg['f_dep'], cfg['f_dec'], seed=4)
    finally:
        networks.SparseToDensePool, net_utils.CalibratedBackprojectionBlock = orig[0], orig[1]
    assert isinstance(new.sparse_to_dense_pool.module, kbnet_b200.SparseToDensePool)
    assert isinstance(new.encoder.module.calibrated_backprojection2, kbnet_b200.CalibratedBackprojectionBlock)
    new.sparse_to_dense_pool.load_state_dict(ref.sparse_to_dense_pool.state_dict(), strict=True)
    new.encoder.load_state_dict(ref.encoder.state_dict(), strict=True)
    new.decoder.load_state_dict(ref.decoder.state_dict(), strict=True)
    ref.eval()
    new.eval()
    return ref, new


@needs_ref
@pytest.mark.parametrize('cfg,n,skew', [(SMALL, 3, False), (SMALL, 2, True), (KITTI, 1, False)])
def test_patched_reference_model_on_gpu_matches_unpatched_reference(cfg, n, skew):
    dev = torch.device('cuda:0')
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False          # the reference's own layers (decoder, conv0) stay exact fp32 on cuDNN
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref, new = _pair(cfg, dev)
        args = _inputs(n, cfg['h'], cfg['w'], 11, skew)
        want = ref.forward(*args)
        lib = kbnet_b200._lib.load()
        n0 = lib.kbnet_launch_count()
        got = new.forward(*[t.to(dev) for t in args])
        assert lib.kbnet_launch_count() - n0 >= 5, 'the patched model did not run the B200 kernels'
        assert_close(got, want, what='patched reference KBNetModel.forward (n=%d skew=%s)' % (n, skew))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@needs_ref
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='DataParallel replicas need 2 visible GPUs')
def test_dataparallel_replicas_on_two_gpus():
    """src/kbnet_model.py:408-415 wraps every network in torch.nn.DataParallel: with two visible devices forward() scatters
    the batch and calls the replicas from one thread per device -- the C ABI must be safe under that."""
    dev = torch.device('cuda:0')
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref, new = _pair(SMALL, dev)
        assert len(new.encoder.device_ids) >= 2
        args = _inputs(4, SMALL['h'], SMALL['w'], 12, True)
        want = ref.forward(*args)
        for _ in range(3):                           # repeated: replicas re-pack weights, threads race on first use
            got = new.forward(*[t.to(dev) for t in args])
            assert_close(got, want, what='DataParallel x2')
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
