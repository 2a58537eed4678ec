"""Generate golden fixtures by running the UNMODIFIED reference modules (CPU, fp32).

Run in the build container (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes small ``*.pt`` fixtures next to this file: seeded inputs, the reference
modules' own ``state_dict`` and the reference outputs.  Tests load the fixtures
only; nothing at test/bench time reads /root/reference.

Import recipe: SURVEY.md Appendix C (flat modules under src/, matplotlib stubbed).
"""
import hashlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('KBNET_REF', '/root/reference/src')


def import_reference():
    sys.path.insert(0, REF)
    for m in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    import networks, net_utils, kbnet_model  # noqa: E401
    return networks, net_utils, kbnet_model


def sparse_depth(gen, n, h, w, p, lo=1.0, hi=80.0):
    mask = (torch.rand(n, 1, h, w, generator=gen) < p).float()
    return (torch.rand(n, 1, h, w, generator=gen) * (hi - lo) + lo) * mask


def sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def pyramid_of(s2d, x):
    grabbed = {}
    h = s2d.pool_convs.register_forward_pre_hook(lambda m, a: grabbed.__setitem__('p', a[0].clone()))
    out = s2d(x)
    h.remove()
    return out, grabbed['p']


def write_outlier_cases(net_utils):
    gen = torch.Generator().manual_seed(7)
    cases = []
    for name, (n, h, w, p, ks, thr) in {'kitti_like': (2, 48, 80, 0.08, 7, 1.5), 'dense_k5': (1, 33, 41, 0.6, 5, 0.5),
                                        'sparse_k9': (2, 40, 52, 0.01, 9, 2.0), 'empty': (1, 16, 20, 0.0, 7, 1.5)}.items():
        z = sparse_depth(gen, n, h, w, p)
        if name == 'kitti_like':                         # plant near/far pairs so that points really get removed, also at borders
            z[0, 0, 0, 0], z[0, 0, 0, 2] = 5.0, 30.0
            z[0, 0, 20, 30], z[0, 0, 22, 33] = 4.0, 5.4
            z[1, 0, 47, 79], z[1, 0, 45, 77] = 60.0, 3.0
        valid = torch.where(z > 0, torch.ones_like(z), z)
        d, v = net_utils.OutlierRemoval(kernel_size=ks, threshold=thr).remove_outliers(sparse_depth=z, validity_map=valid)
        cases.append(dict(name=name, kernel_size=ks, threshold=thr, sparse_depth=z, validity_map=valid, out_depth=d, out_validity=v))
    torch.save(cases, os.path.join(HERE, 'outlier_cases.pt'))


def main():
    networks, net_utils, kbnet_model = import_reference()
    torch.set_grad_enabled(False)
    gen = torch.Generator().manual_seed(1)

    # ---- S2D, small, with the quirk cases of SURVEY 8a-1 ---------------------------------
    cases = []
    for name, (n, cin, h, w), mins, maxs, nconv in [
        ('s2d_small_quirks', (2, 2, 48, 80), [3, 5], [3], 3),
        ('s2d_kitti_sizes', (1, 2, 40, 72), [5, 7, 9, 11, 13], [15, 17], 3),
        ('s2d_void_sizes_c1', (1, 1, 45, 67), [15, 17], [23, 27, 29], 2),
        ('s2d_dropped_size1', (1, 3, 17, 23), [1, 3], [1, 5, 7], 1),
    ]:
        torch.manual_seed(0)
        mod = networks.SparseToDensePool(cin, mins, maxs, 8, nconv, 'xavier_normal', 'leaky_relu').eval()
        z = sparse_depth(gen, n, h, w, 0.05)
        if 'quirks' in name:
            z[0, 0, 0, 0] = 7.5            # corner point
            z[0, 0, 47, 79] = 3.25         # opposite corner
            z[0, 0, 10, 10] = 1000.0       # >= 999 loses to empties
            z[0, 0, 20, 30] = 999.0        # exactly 999 -> 0
            z[0, 0, 30, 40] = -2.0         # negatives participate
            z[0, 0, 30, 41] = -0.0         # -0.0 is empty
            z[1, 0, 5, 70] = float('nan')  # NaN propagates
            z[1, 0, 40, 3] = 1e-30
        chans = [z]
        for c in range(1, cin):
            chans.append((z > 0).float() if c == 1 else torch.rand(n, 1, h, w, generator=gen))
        x = torch.cat(chans, dim=1)
        out, pyr = pyramid_of(mod, x)
        cases.append(dict(name=name, x=x, min_pool_sizes=mins, max_pool_sizes=maxs, n_convolution=nconv,
                          state_dict={k: v.clone() for k, v in mod.state_dict().items()},
                          pyramid=pyr, out=out))
    torch.save(cases, os.path.join(HERE, 's2d_cases.pt'))

    # ---- KB block, small: odd/even dims, fused None / present, skewed per-sample K ----------
    act = torch.nn.LeakyReLU(negative_slope=0.20, inplace=True)
    kb_cases = []
    for name, (n, ci, cd, cf, h, w), (fi, fd, ff), with_fused in [
        ('kb_even_fused', (2, 4, 2, 4, 14, 18), (6, 4, 6), True),
        ('kb_odd_nofused', (2, 6, 4, 0, 13, 19), (8, 4, 8), False),
        ('kb_odd_fused_wide', (1, 16, 8, 16, 21, 37), (32, 16, 32), True),
    ]:
        torch.manual_seed(0)
        in_fused = ci + cf
        mod = net_utils.CalibratedBackprojectionBlock(ci, cd, in_fused, fi, fd, ff, 1, 1, 1, 'xavier_normal', act).eval()
        image = torch.rand(n, ci, h, w, generator=gen)
        depth = torch.rand(n, cd, h, w, generator=gen) - 0.3
        fused = torch.rand(n, cf, h, w, generator=gen) if with_fused else None
        k = torch.tensor([[40.0, 0.7, w / 2 - 0.3], [0.0, 38.0, h / 2 + 0.2], [0.0, 0.0, 1.0]]).view(1, 3, 3).repeat(n, 1, 1)
        k[-1, 0, 0] += 3.0
        xy = net_utils.meshgrid(n, h, w, k.device, True).view(n, 3, -1)
        coords = torch.matmul(torch.inverse(k), xy).view(n, 3, h, w)
        o_img, o_dep, o_fus = mod(image=image, depth=depth, coordinates=coords, fused=fused)
        kb_cases.append(dict(name=name, image=image, depth=depth, fused=fused, k=k, coordinates=coords,
                             state_dict={kk: v.clone() for kk, v in mod.state_dict().items()},
                             conv_image=o_img, conv_depth=o_dep, conv_fused=o_fus))
    torch.save(kb_cases, os.path.join(HERE, 'kb_cases.pt'))

    # ---- whole model, reduced widths (full widths = 27.8 MB of weights) ---------------------
    torch.manual_seed(0)
    f_img, f_dep, f_dec = [8, 16, 16, 32, 32], [4, 8, 8, 16, 16], [32, 16, 16, 8, 4]
    model = kbnet_model.KBNetModel(3, 2, [5, 7, 9], [11, 13], 3, 8, f_img, f_dep, [0, 1, 2, 3], f_dec,
                                   deconv_type='up', weight_initializer='xavier_normal', activation_func='leaky_relu',
                                   min_predict_depth=1.5, max_predict_depth=100.0, device=torch.device('cpu'))
    model.eval()
    n, h, w = 2, 64, 96
    image = torch.rand(n, 3, h, w, generator=gen)
    depth = sparse_depth(gen, n, h, w, 0.05)
    valid = (depth > 0).float()
    k = torch.tensor([[60.0, 0.0, 47.5], [0.0, 60.0, 31.5], [0.0, 0.0, 1.0]]).view(1, 3, 3).repeat(n, 1, 1)
    enc = model.encoder.module
    coords = {}
    hooks = []
    for i in (1, 2, 3, 4):
        blk = getattr(enc, 'calibrated_backprojection%d' % i)
        hooks.append(blk.register_forward_pre_hook(
            (lambda idx: lambda m, a, kw: coords.__setitem__(idx, kw['coordinates'].clone()))(i), with_kwargs=True))
    s2d_out = model.sparse_to_dense_pool(torch.cat([depth, valid], 1))
    latent, skips = model.encoder(image, s2d_out, k)
    out = model.forward(image, depth, valid, k)
    for hk in hooks:
        hk.remove()
    strip = lambda sd: {kk.replace('module.', '', 1): v.clone() for kk, v in sd.items()}
    torch.save(dict(image=image, sparse_depth=depth, validity_map=valid, intrinsics=k,
                    f_img=f_img, f_dep=f_dep, f_dec=f_dec, min_pool_sizes=[5, 7, 9], max_pool_sizes=[11, 13],
                    min_predict_depth=1.5, max_predict_depth=100.0,
                    s2d_state_dict=strip(model.sparse_to_dense_pool.state_dict()),
                    enc_state_dict=strip(model.encoder.state_dict()),
                    dec_state_dict=strip(model.decoder.state_dict()),
                    raw_state_dict_keys=list(model.encoder.state_dict().keys()),
                    coordinates=[coords[i] for i in (1, 2, 3, 4)],
                    s2d_out=s2d_out, latent=latent, skips=skips, output_depth=out),
               os.path.join(HERE, 'model_small.pt'))

    # ---- full-size config 1: pyramid SHA-256 (bit-exact target) + strided samples of the output ----
    gen2 = torch.Generator().manual_seed(1)
    torch.manual_seed(0)
    mod = networks.SparseToDensePool(1, [5, 7, 9, 11, 13], [15, 17], 8, 3, 'xavier_normal', 'leaky_relu').eval()
    mask = (torch.rand(1, 1, 352, 1216, generator=gen2) < 0.05).float()
    z = (torch.rand(1, 1, 352, 1216, generator=gen2) * 79.0 + 1.0) * mask
    out, pyr = pyramid_of(mod, z)
    torch.save(dict(state_dict={k: v.clone() for k, v in mod.state_dict().items()},
                    x_sha256=sha(z), pyramid_sha256=sha(pyr),
                    out_sample=out[:, :, ::16, ::16].clone(), out_absmax=float(out.abs().max()),
                    out_mean=float(out.mean())),
               os.path.join(HERE, 's2d_config1_fullsize.pt'))
    # ---- OutlierRemoval.remove_outliers (src/net_utils.py:1761-1806) with the validity map of src/kbnet.py:899-902 ----
    write_outlier_cases(net_utils)
    print('golden fixtures written to', HERE)
    for f in sorted(os.listdir(HERE)):
        print('  %-32s %8d bytes' % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == '__main__':
    main()
