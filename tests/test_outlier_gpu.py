"""GPU parity of OutlierRemoval.remove_outliers (through the C ABI): bit-exact against reference-generated fixtures and, at
full KITTI / VOID size, against the oracle."""
import pytest
import torch

from conftest import load_golden
from oracle import kbnet_oracle as K
import kbnet_b200

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = 'cuda:0'


@pytest.mark.parametrize('idx', range(4))
@pytest.mark.parametrize('with_validity', [True, False])
def test_golden_cases(idx, with_validity):
    c = load_golden('outlier_cases.pt')[idx]
    m = kbnet_b200.OutlierRemoval(kernel_size=c['kernel_size'], threshold=c['threshold'])
    d, v = m.remove_outliers(c['sparse_depth'].to(DEV), c['validity_map'].to(DEV) if with_validity else None)
    assert torch.equal(d.cpu(), c['out_depth']) and torch.equal(v.cpu(), c['out_validity']), c['name']


@pytest.mark.parametrize('cfg', ['KITTI', 'VOID'])
def test_fullsize_vs_oracle_and_idempotence(cfg):
    c = K.KITTI if cfg == 'KITTI' else K.VOID
    gen = torch.Generator().manual_seed(9)
    z, valid = K.synth_sparse_depth(gen, 4, c['h'], c['w'], c['p'] * 2, c['dmin'], c['dmax'])
    ref_d, ref_v = K.outlier_removal(z, valid)
    m = kbnet_b200.OutlierRemoval()
    d, v = m.remove_outliers(z.to(DEV), valid.to(DEV))
    assert torch.equal(d.cpu(), ref_d) and torch.equal(v.cpu(), ref_v)
    assert int((ref_v > 0).sum()) < int((valid > 0).sum()), 'the synthetic input should contain outliers'
    # size-independent property: surviving points are subsets and the input is not modified
    assert bool(((v.cpu() > 0) <= (valid > 0)).all())
    d2, v2 = m.remove_outliers(z.to(DEV), valid.to(DEV))
    assert torch.equal(d2, d) and torch.equal(v2, v)


def test_empty_batch_and_errors():
    m = kbnet_b200.OutlierRemoval()
    d, v = m.remove_outliers(torch.zeros(0, 1, 8, 8, device=DEV))
    assert d.shape == (0, 1, 8, 8) and v.shape == (0, 1, 8, 8)
    with pytest.raises(RuntimeError):
        m.remove_outliers(torch.zeros(1, 1, 8, 8))          # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        kbnet_b200.OutlierRemoval(kernel_size=19).remove_outliers(torch.zeros(1, 1, 8, 8, device=DEV))
