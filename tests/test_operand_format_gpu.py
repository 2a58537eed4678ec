"""GPU parity of the operand-format side channel (include/kbnet_b200.h: kbnet_to/from_operand_format, kbnet_kb_forward_of,
kbnet_conv3x3_small_forward_of): an activation handed from one of our kernels to the next as bf16 hi/lo 16-byte entries must
give the same results as the NCHW fp32 path, and those match the oracle (reference src/net_utils.py:1343-1371,
src/networks.py:364-475)."""
import pytest
import torch

from util import assert_close
from oracle import kbnet_oracle as K
import kbnet_b200
from kbnet_b200.modules import OperandFormat

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = 'cuda:0'


def _block(ci, cd, cf, fi, fd, ff, sd):
    m = kbnet_b200.CalibratedBackprojectionBlock(ci, cd, ci + cf, fi, fd, ff, 1, 1, 1, 'xavier_normal',
                                                 torch.nn.LeakyReLU(negative_slope=0.20, inplace=True))
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval()


@pytest.mark.parametrize('shape', [(2, 8, 6, 12), (1, 48, 34, 64), (3, 16, 16, 20)])
def test_round_trip(shape):
    """to_operand_format -> from_operand_format gives hi + lo: within 2^-16 relative of the input, exact for bf16-representable values"""
    gen = torch.Generator().manual_seed(3)
    x = (torch.rand(*shape, generator=gen) - 0.3) * 50.0
    x[0, 0, 0, :4] = torch.tensor([0.0, 1.0, -2.5, 1e-20])
    of = OperandFormat.from_nchw(x.to(DEV))
    y = of.to_nchw().cpu()
    assert float(((y - x).abs() / x.abs().clamp_min(1e-30)).max()) <= 2.0 ** -15
    assert torch.equal(y[0, 0, 0, :3], x[0, 0, 0, :3])
    # entry layout: [n, c/8, parity plane (y&1)*2 + (x&1), y/2, x/2, 8 consecutive channels], hi = round-to-nearest bf16
    n, c, h, w = shape
    hi = of.hi.float().cpu()
    xb = x.to(torch.bfloat16).float().view(n, c // 8, 8, h // 2, 2, w // 2, 2)      # [n, cg, ch, yy, py, xx, px]
    want = xb.permute(0, 1, 4, 6, 3, 5, 2).reshape(n, c // 8, 4, h // 2, w // 2, 8)
    assert torch.equal(hi, want)
    # the even-pixel form is plane 0
    ev = OperandFormat.from_nchw(x.to(DEV), even_only=True)
    assert torch.equal(ev.hi, of.hi[:, :, :1]) and torch.equal(ev.lo, of.lo[:, :, :1])
    # a channel slice of a wider buffer as the source
    wide = torch.rand(shape[0], shape[1] + 8, shape[2], shape[3], generator=gen).to(DEV)
    of2 = OperandFormat.from_nchw(wide[:, 8:])
    assert torch.equal(of2.to_nchw(), OperandFormat.from_nchw(wide[:, 8:].contiguous()).to_nchw())
    # odd sizes have no parity planes of equal size
    with pytest.raises(RuntimeError):
        OperandFormat.from_nchw(torch.rand(1, 8, 5, 12, device=DEV))


@pytest.mark.parametrize('cfg', [
    # (n, C_i, C_d, C_f, F_i, F_d, F_f, h, w)
    (2, 16, 16, 0, 16, 16, 16, 32, 80),          # block-1 form: no incoming fused
    (2, 48, 16, 48, 96, 32, 96, 44, 152),        # block-2 channels
    (1, 96, 32, 96, 192, 64, 192, 22, 76),       # block-3 channels: filter split
    (2, 32, 16, 8, 24, 16, 40, 18, 36),          # ragged tiles, odd output height, F_i not a multiple of 16
])
def test_kb_operand_format_matches_fp32_path_and_oracle(cfg):
    n, ci, cd, cf, fi, fd, ff, h, w = cfg
    gen = torch.Generator().manual_seed(sum(cfg))
    image, depth = torch.rand(n, ci, h, w, generator=gen) - 0.25, torch.rand(n, cd, h, w, generator=gen)
    fused = (torch.rand(n, cf, h, w, generator=gen) - 0.5) if cf else None
    k = torch.tensor([[60.0, 0.5, w / 2.0], [0, 61.0, h / 2.0], [0, 0, 1]]).view(1, 3, 3).repeat(n, 1, 1)
    wts = K.make_kb_weights(torch.Generator().manual_seed(1), ci, cd, ci + cf, fi, fd, ff)
    ref = K.kb_forward(image, depth, K.camera_coordinates(k, h, w), fused, wts)
    m = _block(ci, cd, cf, fi, fd, ff, wts)
    assert m.operand_format_supported(h, w)
    kinv = torch.inverse(k).to(DEV)
    d_image, d_depth, d_fused = image.to(DEV), depth.to(DEV), (fused.to(DEV) if cf else None)
    plain = m(d_image, d_depth, kinv, d_fused)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    # (a) image in, conv_image out in operand format
    of_out_ok = fi % 8 == 0 and ho % 2 == 0 and wo % 2 == 0
    out_of = OperandFormat(n, fi, ho, wo, DEV) if of_out_ok else None
    got = m(OperandFormat.from_nchw(d_image), d_depth, kinv, d_fused, out_image=out_of)
    conv_image = got[0].to_nchw() if out_of is not None else got[0]
    for g, p, r, name in zip((conv_image, got[1], got[2]), plain, ref, ('conv_image', 'conv_depth', 'conv_fused')):
        assert_close(g, r, what='of/%s' % name)
        # same split of the same values -> the two paths agree far below the parity tolerance (conv_image: one more bf16 hi/lo
        # rounding on the way out)
        assert_close(g, p, tol=2e-5, what='of-vs-fp32-path/%s' % name)
    # (b) fused in / conv_fused out in operand format as well
    if cf:
        of_f = OperandFormat(n, ff, ho, wo, DEV, even_only=True) if (ff % 8 == 0 and ho % 2 == 0 and wo % 2 == 0) else None
        got = m(OperandFormat.from_nchw(d_image), d_depth, kinv, None, fused_of=OperandFormat.from_nchw(d_fused, even_only=True),
                out_fused_of=of_f)
        for g, r, name in zip(got, ref, ('conv_image', 'conv_depth', 'conv_fused')):
            assert_close(g, r, what='of-fused/%s' % name)
        if of_f is not None:   # the even-pixel copy == the even-pixel form of the fp32 output, bit for bit
            want = OperandFormat.from_nchw(got[2], even_only=True)
            assert torch.equal(of_f.hi, want.hi) and torch.equal(of_f.lo, want.lo)
    # (c) conv_image not wanted: the other two outputs are unchanged, from fp32 and from operand-format sources
    for img in (d_image, OperandFormat.from_nchw(d_image)):
        got = m(img, d_depth, kinv, d_fused, need_image=False)
        assert got[0] is None
        assert_close(got[1], ref[1], what='skip-image/conv_depth')
        assert_close(got[2], ref[2], what='skip-image/conv_fused')
        assert_close(got[1], plain[1], tol=2e-5, what='skip-image/conv_depth vs the full block')   # (another tile plan: not bit-equal)


def test_kb_operand_format_coordinates_tensor_path():
    """the reference's [N,3,H,W] coordinates tensor together with operand-format image"""
    n, ci, cd, cf, fi, fd, ff, h, w = 2, 16, 16, 16, 32, 16, 32, 24, 40
    gen = torch.Generator().manual_seed(11)
    image, depth, fused = (torch.rand(n, c, h, w, generator=gen) for c in (ci, cd, cf))
    k = torch.tensor([[50.0, 0, 20], [0, 50.0, 12], [0, 0, 1]]).view(1, 3, 3).repeat(n, 1, 1)
    coords = K.camera_coordinates(k, h, w)
    wts = K.make_kb_weights(torch.Generator().manual_seed(2), ci, cd, ci + cf, fi, fd, ff)
    ref = K.kb_forward(image, depth, coords, fused, wts)
    m = _block(ci, cd, cf, fi, fd, ff, wts)
    got = m(OperandFormat.from_nchw(image.to(DEV)), depth.to(DEV), coords.to(DEV), fused.to(DEV))
    for g, r, name in zip(got, ref, ('conv_image', 'conv_depth', 'conv_fused')):
        assert_close(g, r, what='of-coords/%s' % name)


def test_kb_operand_format_rejected_outside_the_fused_program():
    """C_d = 8 runs the unfused launches: the side channel must raise, not be ignored"""
    n, ci, cd, fi, fd, ff, h, w = 1, 16, 8, 16, 16, 16, 16, 32
    wts = K.make_kb_weights(torch.Generator().manual_seed(2), ci, cd, ci, fi, fd, ff)
    m = _block(ci, cd, 0, fi, fd, ff, wts)
    assert not m.operand_format_supported(h, w)
    image, depth = torch.rand(n, ci, h, w, device=DEV), torch.rand(n, cd, h, w, device=DEV)
    kinv = torch.eye(3, device=DEV).view(1, 3, 3)
    with pytest.raises(RuntimeError, match='operand-format'):
        m(OperandFormat.from_nchw(image), depth, kinv)
    got = m(image, depth, kinv, need_image=False)      # unfused path without conv_image
    full = m(image, depth, kinv)
    assert got[0] is None and torch.equal(got[1], full[1]) and torch.equal(got[2], full[2])   # separate launches: bit-equal


@pytest.mark.parametrize('c,f,h,w', [(3, 48, 40, 136), (3, 48, 20, 70), (8, 16, 32, 64)])
def test_small_conv_writes_operand_format(c, f, h, w):
    """conv0_image (src/networks.py:364) writing its result in operand format only"""
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(2, c, h, w, generator=gen)
    wt = torch.randn(f, c, 3, 3, generator=gen) * 0.2
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, wt, padding=1), 0.2)
    wd = torch.nn.Parameter(wt.to(DEV), requires_grad=False)
    of = kbnet_b200.ops.conv2d_fused([x.to(DEV)], wd, 1, 0.2, out_of=OperandFormat(2, f, h, w, DEV))
    assert_close(of.to_nchw(), ref, what='conv_mma operand-format output')
    plain = kbnet_b200.ops.conv2d_fused([x.to(DEV)], wd, 1, 0.2)
    assert_close(of.to_nchw(), plain, tol=2e-5, what='operand-format output vs NCHW output')


def test_model_with_and_without_operand_format():
    """the whole model at KBNet's channel widths (src/networks.py:317-533 wiring) with conv_image travelling in operand format ==
    the NCHW fp32 model, and both match the oracle; dropping block 4's dead conv_image changes nothing"""
    from kbnet_b200 import model as M
    c = K.KITTI
    gen = torch.Generator().manual_seed(9)
    n, h, w = 2, 96, 256
    image = torch.rand(n, 3, h, w, generator=gen)
    depth, valid = K.synth_sparse_depth(gen, n, h, w, c['p'], c['dmin'], c['dmax'])
    k = torch.tensor([[150.0, 0.0, 128.0], [0.0, 150.0, 48.0], [0.0, 0.0, 1.0]]).view(1, 3, 3).repeat(n, 1, 1)
    ws, we, wd = K.make_model_weights(torch.Generator().manual_seed(0), 2, 7)
    ref = K.model_forward(image, depth, valid, k, ws, we, wd, c['min_pools'], c['max_pools'], c['min_pred'], c['max_pred'])
    mdl = kbnet_b200.KBNetModel(3, 2, c['min_pools'], c['max_pools'], 3, 8, [48, 96, 192, 384, 384], [16, 32, 64, 128, 128],
                                [0, 1, 2, 3], [256, 128, 128, 64, 12], min_predict_depth=c['min_pred'],
                                max_predict_depth=c['max_pred'], device=torch.device(DEV))
    mdl.load_state_dicts(ws, we, wd)
    args = [t.to(DEV) for t in (image, depth, valid, k)]
    outs = {}
    of0, skip0 = M.OPERAND_FORMAT, M.SKIP_DEAD_CONV_IMAGE
    lib = kbnet_b200._lib.load()
    launches = {}
    mdl.forward(*args)   # weight packs are built (and counted as launches) on first use
    for of_on, skip in ((True, True), (False, False), (True, False), (False, True)):
        M.OPERAND_FORMAT, M.SKIP_DEAD_CONV_IMAGE = of_on, skip
        try:
            n0 = lib.kbnet_launch_count()
            outs[(of_on, skip)] = mdl.forward(*args).clone()
            launches[(of_on, skip)] = lib.kbnet_launch_count() - n0
        finally:
            M.OPERAND_FORMAT, M.SKIP_DEAD_CONV_IMAGE = of0, skip0
    base = outs[(False, False)]
    assert len(set(launches.values())) == 1, 'operand format must not add launches: %r' % (launches,)
    # (not bit-equal in general: the program without conv_image may get another tile plan, e.g. the dual-N split mode)
    assert_close(outs[(False, True)], base, tol=2e-5, what='dropping the dead conv_image of block 4 must not change the output')
    assert_close(outs[(True, True)], outs[(True, False)], tol=2e-5, what='... nor in operand-format mode')
    assert_close(outs[(True, True)], base, tol=2e-5, what='model operand format vs NCHW')
    assert_close(outs[(True, True)], ref, what='model operand format vs oracle')
    assert_close(base, ref, what='model NCHW vs oracle')
