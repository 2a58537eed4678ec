"""Shared test helpers: the parity metric (tolerance written once, SURVEY 7 'hard parts')."""
import torch

TOL = 1e-3  # north-star tolerance for floating-point outputs (relative, fp32)


def eq_nan(a, b):
    """bit-exact on non-NaN values (torch.equal treats -0.0 == 0.0), identical NaN positions"""
    return torch.equal(a.isnan(), b.isnan()) and torch.equal(torch.nan_to_num(a, nan=0.0), torch.nan_to_num(b, nan=0.0))


def rel_err(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    m = ~(ref.isnan() | got.isnan())
    return float((got[m] - ref[m]).abs().max() / ref[m].abs().max().clamp_min(1e-30))


def assert_close(got, ref, tol=TOL, what=''):
    """max|got-ref| <= tol * max|ref|   AND   |got-ref| <= tol*|ref| + tol*rms(ref) element-wise."""
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, '%s: shape %s vs %s' % (what, tuple(got.shape), tuple(ref.shape))
    assert torch.equal(got.isnan(), ref.isnan()), '%s: NaN positions differ' % what
    m = ~ref.isnan()
    g, r = got[m].double(), ref[m].double()
    if r.numel() == 0:
        return
    err = (g - r).abs()
    scale = r.abs().max().clamp_min(1e-30)
    rms = r.pow(2).mean().sqrt()
    assert float(err.max() / scale) <= tol, '%s: max-norm error %.3e > %.1e' % (what, float(err.max() / scale), tol)
    bad = err > tol * r.abs() + tol * rms
    assert not bool(bad.any()), '%s: %d of %d elements outside rtol=atol/rms=%.1e (worst %.3e)' % (
        what, int(bad.sum()), r.numel(), tol, float((err / (r.abs() + rms)).max()))
