"""CPU oracle for the KBNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``kbnet_b200``) never imports it and fails loudly when the CUDA library is
missing.

This is an independent functional restatement (torch CPU ops on plain tensors,
weights passed as a flat ``dict`` keyed like the reference ``state_dict``) of the
reference algorithm for the path named in BASELINE.json:north_star.  Each function
cites the reference file:line it follows (paths relative to the reference repo
alexklwong/calibrated-backprojection-network @ 7995b27).

Parity status: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the reference itself:
``tests/golden/make_golden.py`` imports the unmodified reference modules in the
build container, runs them on seeded inputs and commits the input/output tensors
under ``tests/golden/``; ``tests/test_oracle.py`` checks this file against those
fixtures (pool pyramid bit-exact, conv outputs <= 1e-6) and, when
``/root/reference`` is present, against the live reference modules.
"""
import math

import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.20  # src/net_utils.py:37  LeakyReLU(negative_slope=0.20)


def _act(x, slope):
    # src/net_utils.py:120-141  Conv2d.forward: conv -> activation
    return x if slope is None else F.leaky_relu(x, slope)


def conv2d(x, w, stride=1, slope=LEAKY_SLOPE):
    """src/net_utils.py:85-93,120-141: bias-free conv, padding = k // 2, then activation."""
    k = w.shape[-1]
    return _act(F.conv2d(x, w, None, stride=stride, padding=k // 2), slope)


# --------------------------------------------------------------------------------------
# SparseToDensePool
# --------------------------------------------------------------------------------------
def s2d_pool_pyramid(z, min_pool_sizes, max_pool_sizes):
    """src/networks.py:2172-2189.  z: [N,1,H,W].  Returns [N, n_min + n_max, H, W].

    min pool  = -maxpool_s(where(z == 0, -999, -z)), then 999 -> 0   (:2177-2179)
    max pool  =  maxpool_s(z)  (implicit -inf padding)                 (:2184)
    Sizes <= 1 are dropped (src/networks.py:2112-2118).
    """
    mins = [s for s in min_pool_sizes if s > 1]
    maxs = [s for s in max_pool_sizes if s > 1]
    pyr = []
    for s in mins:
        flagged = torch.where(z == 0, torch.full_like(z, -999.0), -z)
        m = -F.max_pool2d(flagged, kernel_size=s, stride=1, padding=s // 2)
        pyr.append(torch.where(m == 999, torch.zeros_like(m), m))
    for s in maxs:
        pyr.append(F.max_pool2d(z, kernel_size=s, stride=1, padding=s // 2))
    return torch.cat(pyr, dim=1)


def s2d_forward(x, weights, min_pool_sizes, max_pool_sizes, slope=LEAKY_SLOPE, return_pyramid=False):
    """src/networks.py:2168-2196.

    weights: {'pool_convs.{i}.conv.weight': [F, C, 1, 1], 'conv.conv.weight': [F, F + C_in, 3, 3]}
    """
    z = x[:, 0:1]
    pyr = s2d_pool_pyramid(z, min_pool_sizes, max_pool_sizes)
    q = pyr
    i = 0
    while 'pool_convs.%d.conv.weight' % i in weights:
        q = conv2d(q, weights['pool_convs.%d.conv.weight' % i], 1, slope)
        i += 1
    out = conv2d(torch.cat([q, x], dim=1), weights['conv.conv.weight'], 1, slope)
    return (out, pyr) if return_pyramid else out


# --------------------------------------------------------------------------------------
# Camera coordinates / intrinsics scaling
# --------------------------------------------------------------------------------------
def meshgrid_h(n, h, w, device=None):
    """src/net_utils.py:1601-1636: [N,3,H,W] homogeneous pixel grid (x = col, y = row, 1)."""
    xs = torch.arange(w, dtype=torch.float32, device=device).view(1, 1, 1, w).expand(n, 1, h, w)
    ys = torch.arange(h, dtype=torch.float32, device=device).view(1, 1, h, 1).expand(n, 1, h, w)
    return torch.cat([xs, ys, torch.ones(n, 1, h, w, device=device)], dim=1)


def camera_coordinates(k, h, w):
    """src/networks.py:317-331: coordinates = inverse(K) @ [x, y, 1]  ->  [N,3,H,W]."""
    n = k.shape[0]
    xy_h = meshgrid_h(n, h, w, k.device).reshape(n, 3, -1)
    return torch.matmul(torch.inverse(k), xy_h).view(n, 3, h, w)


def scale_intrinsics_level1(k, h0, w0, h1, w1):
    """src/networks.py:333-352.  The reference closure ignores its own arguments and
    always uses the level-1 / level-0 ratio (:342-343), so levels 1, 2 and 3 all get
    K * [[sx,1,sx],[1,sy,sy],[1,1,1]] with sx = W1/W0, sy = H1/H0."""
    sx = w1 / w0
    sy = h1 / h0
    scale = torch.tensor([[sx, 1.0, sx], [1.0, sy, sy], [1.0, 1.0, 1.0]], dtype=torch.float32, device=k.device)
    return k * scale.view(1, 3, 3)


# --------------------------------------------------------------------------------------
# Calibrated backprojection block
# --------------------------------------------------------------------------------------
def kb_forward(image, depth, coordinates, fused, weights, slope=LEAKY_SLOPE, prefix=''):
    """src/net_utils.py:1343-1371.  Returns (conv_image, conv_depth, conv_fused)."""
    w_img = weights[prefix + 'conv_image.conv_block.0.conv.weight']
    w_dep = weights[prefix + 'conv_depth.conv_block.0.conv.weight']
    w_prj = weights[prefix + 'proj_depth.conv.weight']
    w_fus = weights[prefix + 'conv_fused.conv.weight']
    conv_image = conv2d(image, w_img, 2, slope)                                   # :1348
    conv_depth = conv2d(torch.cat([depth, coordinates], dim=1), w_dep, 2, slope)  # :1351
    z = conv2d(depth, w_prj, 1, slope)                                            # :1357
    xyz = coordinates * z                                                         # :1360
    parts = [image, xyz] + ([fused] if fused is not None else [])                # :1354-1365
    conv_fused = conv2d(torch.cat(parts, dim=1), w_fus, 2, slope)                 # :1368-1369 (1x1, stride 2)
    return conv_image, conv_depth, conv_fused


# --------------------------------------------------------------------------------------
# Encoder / decoder / whole model  (callers either side of the path: SURVEY 8 rows a-8, a-9, f-1, f-2)
# --------------------------------------------------------------------------------------
def encoder_forward(image, depth, intrinsics, weights, slope=LEAKY_SLOPE, prefix='', return_coords=False):
    """src/networks.py:301-533 with resolutions_backprojection = [0,1,2,3]."""
    p = prefix
    n, _, h0, w0 = image.shape
    coords_all = []
    coords0 = camera_coordinates(intrinsics, h0, w0)                               # :361
    coords_all.append(coords0)
    c0_img = conv2d(image, weights[p + 'conv0_image.conv.weight'], 1, slope)       # :364
    c0_dep = conv2d(depth, weights[p + 'conv0_depth.conv.weight'], 1, slope)       # :365
    ci, cd, cf = kb_forward(c0_img, c0_dep, coords0, None, weights, slope, p + 'calibrated_backprojection1.')
    skips = [torch.cat([cf, cd], dim=1)]                                           # :374,383
    h1, w1 = ci.shape[-2:]
    k_scaled = scale_intrinsics_level1(intrinsics, h0, w0, h1, w1)                 # same K for levels 1..3
    for lvl in (2, 3, 4):
        hh, ww = ci.shape[-2:]
        coords = camera_coordinates(k_scaled, hh, ww)                              # :392, :429, :466
        coords_all.append(coords)
        ci, cd, cf = kb_forward(ci, cd, coords, cf, weights, slope, p + 'calibrated_backprojection%d.' % lvl)
        skips.append(torch.cat([cf, cd], dim=1))
    c5_img = conv2d(cf, weights[p + 'conv5_image.conv_block.0.conv.weight'], 2, slope)   # :520-523 (input = conv4_fused)
    c5_dep = conv2d(cd, weights[p + 'conv5_depth.conv_block.0.conv.weight'], 2, slope)   # :527
    latent = torch.cat([c5_img, c5_dep], dim=1)
    if return_coords:
        return latent, skips, coords_all
    return latent, skips


def decoder_forward(latent, skips, shape, weights, slope=LEAKY_SLOPE, prefix=''):
    """src/networks.py:1855-1989 with n_resolution = 1, deconv_type = 'up', output_func = 'linear';
    DecoderBlock src/net_utils.py:1453-1487, UpConv2d :484-499."""
    p = prefix
    x = latent
    for name, skip in (('deconv4', skips[3]), ('deconv3', skips[2]), ('deconv2', skips[1]),
                       ('deconv1', skips[0]), ('deconv0', None)):
        size = tuple(skip.shape[-2:]) if skip is not None else tuple(shape)
        up = F.interpolate(x, size=size, mode='nearest')                            # net_utils.py:497
        d = conv2d(up, weights[p + name + '.deconv.conv.conv.weight'], 1, slope)
        cat = torch.cat([d, skip], dim=1) if skip is not None else d               # net_utils.py:1482-1485
        x = conv2d(cat, weights[p + name + '.conv.conv.weight'], 1, slope)
    return conv2d(x, weights[p + 'output0.conv.weight'], 1, None)                  # linear output


def model_forward(image, sparse_depth, validity_map, intrinsics, w_s2d, w_enc, w_dec,
                  min_pool_sizes, max_pool_sizes, min_predict_depth, max_predict_depth, slope=LEAKY_SLOPE):
    """src/kbnet_model.py:143-186."""
    x = torch.cat([sparse_depth, validity_map], dim=1)                              # :166-171
    s2d = s2d_forward(x, w_s2d, min_pool_sizes, max_pool_sizes, slope)              # :173
    latent, skips = encoder_forward(image, s2d, intrinsics, w_enc, slope)           # :177
    out = decoder_forward(latent, skips, s2d.shape[-2:], w_dec, slope)              # :179
    sig = torch.sigmoid(out)                                                        # :181
    return min_predict_depth / (sig + min_predict_depth / max_predict_depth)        # :183-184


# --------------------------------------------------------------------------------------
# Random-weight construction with the reference's key names / shapes
# --------------------------------------------------------------------------------------
def _xavier(gen, *shape):
    fan_in = shape[1] * shape[2] * shape[3]
    fan_out = shape[0] * shape[2] * shape[3]
    std = math.sqrt(2.0 / (fan_in + fan_out))
    return torch.randn(*shape, generator=gen) * std


def make_s2d_weights(gen, c_in, n_pool, n_filter=8, n_convolution=3):
    w = {}
    c = n_pool
    for i in range(n_convolution):
        w['pool_convs.%d.conv.weight' % i] = _xavier(gen, n_filter, c, 1, 1)
        c = n_filter
    w['conv.conv.weight'] = _xavier(gen, n_filter, n_filter + c_in, 3, 3)
    return w


def make_kb_weights(gen, c_i, c_d, c_fused_in, f_i, f_d, f_f, prefix=''):
    """c_fused_in = in_channels_fused of the reference ctor (already includes image channels)."""
    return {
        prefix + 'conv_image.conv_block.0.conv.weight': _xavier(gen, f_i, c_i, 3, 3),
        prefix + 'conv_depth.conv_block.0.conv.weight': _xavier(gen, f_d, c_d + 3, 3, 3),
        prefix + 'proj_depth.conv.weight': _xavier(gen, 1, c_d, 1, 1),
        prefix + 'conv_fused.conv.weight': _xavier(gen, f_f, c_fused_in + 3, 1, 1),
    }


def outlier_removal(sparse_depth, validity_map=None, kernel_size=7, threshold=1.5):
    """src/net_utils.py:1761-1806 (and, for validity_map=None, the validity map of src/kbnet.py:899-902)."""
    if validity_map is None:
        validity_map = torch.where(sparse_depth > 0, torch.ones_like(sparse_depth), sparse_depth)
    max_value = 10 * torch.max(sparse_depth)                                          # :1776
    filled = torch.where(validity_map <= 0, torch.full_like(sparse_depth, float(max_value)), sparse_depth)
    pad = kernel_size // 2
    filled = F.pad(filled, (pad, pad, pad, pad), mode='constant', value=float(max_value))   # :1783-1787
    min_values = -F.max_pool2d(-filled, kernel_size=kernel_size, stride=1, padding=0)      # :1789-1793
    keep = torch.where(min_values < sparse_depth - threshold, torch.zeros_like(validity_map), torch.ones_like(validity_map))
    validity_clean = validity_map * keep                                              # :1802
    return sparse_depth * validity_clean, validity_clean


def make_model_weights(gen, c_in_depth=2, n_pool=7, n_filter_s2d=8, n_conv_s2d=3,
                       f_img=(48, 96, 192, 384, 384), f_dep=(16, 32, 64, 128, 128),
                       f_dec=(256, 128, 128, 64, 12)):
    """Weight dicts with the reference's state_dict keys (no 'module.' prefix)."""
    w_s2d = make_s2d_weights(gen, c_in_depth, n_pool, n_filter_s2d, n_conv_s2d)
    enc = {'conv0_image.conv.weight': _xavier(gen, f_img[0], 3, 3, 3),
           'conv0_depth.conv.weight': _xavier(gen, f_dep[0], n_filter_s2d, 3, 3)}
    ci, cd, cfz = f_img[0], f_dep[0], f_img[0]
    for lvl in range(4):
        enc.update(make_kb_weights(gen, ci, cd, cfz, f_img[lvl], f_dep[lvl], f_img[lvl],
                                   'calibrated_backprojection%d.' % (lvl + 1)))
        ci, cd = f_img[lvl], f_dep[lvl]
        cfz = f_img[lvl] + f_img[lvl]  # src/networks.py:142: in_channels_fused = image + previous fused
    enc['conv5_image.conv_block.0.conv.weight'] = _xavier(gen, f_img[4], f_img[3], 3, 3)
    enc['conv5_depth.conv_block.0.conv.weight'] = _xavier(gen, f_dep[4], f_dep[3], 3, 3)
    dec = {}
    n_enc = [a + b for a, b in zip(f_img, f_dep)]
    n_skips = n_enc[:-1][::-1] + [0]
    c = n_enc[-1]
    for i, name in enumerate(('deconv4', 'deconv3', 'deconv2', 'deconv1', 'deconv0')):
        dec[name + '.deconv.conv.conv.weight'] = _xavier(gen, f_dec[i], c, 3, 3)
        dec[name + '.conv.conv.weight'] = _xavier(gen, f_dec[i], f_dec[i] + n_skips[i], 3, 3)
        c = f_dec[i]
    dec['output0.conv.weight'] = _xavier(gen, 1, c, 3, 3)
    return w_s2d, enc, dec


# --------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------
KITTI = dict(h=352, w=1216, min_pools=[5, 7, 9, 11, 13], max_pools=[15, 17], dmin=1.0, dmax=80.0, p=0.05,
             k=[[721.5377, 0.0, 609.5593], [0.0, 721.5377, 172.854], [0.0, 0.0, 1.0]], min_pred=1.5, max_pred=100.0)
VOID = dict(h=480, w=640, min_pools=[15, 17], max_pools=[23, 27, 29], dmin=0.2, dmax=7.0, p=0.005,
            k=[[514.6, 0.0, 319.5], [0.0, 514.6, 239.5], [0.0, 0.0, 1.0]], min_pred=0.1, max_pred=8.0)


def synth_sparse_depth(gen, n, h, w, p, dmin, dmax):
    mask = (torch.rand(n, 1, h, w, generator=gen) < p).float()
    depth = (torch.rand(n, 1, h, w, generator=gen) * (dmax - dmin) + dmin) * mask
    return depth, (depth > 0).float()


def synth_model_inputs(gen, n, cfg):
    image = torch.rand(n, 3, cfg['h'], cfg['w'], generator=gen)
    depth, valid = synth_sparse_depth(gen, n, cfg['h'], cfg['w'], cfg['p'], cfg['dmin'], cfg['dmax'])
    k = torch.tensor(cfg['k'], dtype=torch.float32).view(1, 3, 3).repeat(n, 1, 1)
    return image, depth, valid, k
