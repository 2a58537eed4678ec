"""Functional wrappers over the C ABI for the convolutions either side of the hot path."""
import ctypes

import torch

from . import _lib
from .modules import _view4, _w, _stream, packed_weight


def conv2d_fused(sources, weight, stride=1, slope=0.2, out=None, engine=None, out_of=None):
    """out = act(conv_k(cat_c(sources), weight)), padding k//2, no bias.

    engine=None picks the kernel by shape: a single full-resolution source with few channels and filters (3x3, stride 1)
    goes to the warp-level tensor kernel (kbnet_conv3x3_small_forward), everything else to the tcgen05 engine behind
    kbnet_conv2d_forward; engine='tcgen05' forces the latter (used by the engine's own tests).

    sources: list of 4-D CUDA float32 tensors or ``(tensor, 'up')`` pairs; an ``'up'`` source is nearest-upsampled to
    the spatial size of the first non-upsampled source (or ``size`` given as ``(tensor, 'up', (H, W))``) inside the kernel.
    slope=None means no activation.  ``out`` may be a channel slice of a wider NCHW buffer.
    ``out_of``: an OperandFormat that receives the result INSTEAD of an NCHW tensor (returned; only the few-channel kernel
    writes it -- callers check ``conv3x3_small_takes`` first).
    Replaces Conv2d.forward + torch.cat + F.interpolate(mode='nearest') of the reference
    (src/net_utils.py:120-141, 484-499, 1453-1487; src/networks.py:364-365, 520-531).
    """
    lib = _lib.load()
    parsed, h_in, w_in = [], None, None
    for s in sources:
        if isinstance(s, (tuple, list)):
            t, sn = _view4('source', s[0])
            parsed.append((t, sn, 1))
            if len(s) > 2 and h_in is None:
                h_in, w_in = int(s[2][0]), int(s[2][1])
        else:
            t, sn = _view4('source', s)
            parsed.append((t, sn, 0))
            h_in, w_in = t.shape[2], t.shape[3]
    if h_in is None:
        raise RuntimeError('kbnet_b200: conv2d_fused needs a size for upsample-only inputs')
    dev = parsed[0][0].device
    n = parsed[0][0].shape[0]
    for t, _, _ in parsed[1:]:
        if t.shape[0] != n or t.device != dev:
            raise RuntimeError('kbnet_b200: conv2d_fused sources disagree on batch size / device: %s on %s vs N=%d on %s'
                               % (tuple(t.shape), t.device, n, dev))
    w = _w(weight, dev)
    f, ctot, k, _ = w.shape
    if sum(t.shape[1] for t, _, _ in parsed) != ctot:
        raise RuntimeError('kbnet_b200: conv weight expects %d input channels, sources have %d'
                           % (ctot, sum(t.shape[1] for t, _, _ in parsed)))
    pad = k // 2
    ho, wo = (h_in + 2 * pad - k) // stride + 1, (w_in + 2 * pad - k) // stride + 1
    if out_of is not None:
        if out is not None or tuple(out_of.shape) != (n, f, ho, wo) or out_of.device != dev:
            raise RuntimeError('kbnet_b200: bad operand-format conv output')
        if not conv3x3_small_takes(ctot, f, k, stride, slope) or len(parsed) != 1 or parsed[0][2] != 0:
            raise RuntimeError('kbnet_b200: only the few-channel 3x3 kernel writes operand format')
        t, sn, _ = parsed[0]
        with torch.cuda.device(dev):
            rc = lib.kbnet_conv3x3_small_forward_of(t.data_ptr(), sn, n, ctot, h_in, w_in, w.data_ptr(), f, 0 if slope is None else 1,
                                                    0.0 if slope is None else float(slope), None, f * ho * wo,
                                                    out_of.hi.data_ptr(), out_of.lo.data_ptr(), _stream(dev))
        _lib.check(rc, 'kbnet_conv3x3_small_forward_of')
        return out_of
    if out is None:
        out = torch.empty((n, f, ho, wo), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (n, f, ho, wo) or out.stride()[1:] != (ho * wo, wo, 1) or out.dtype != torch.float32 or out.device != dev:
        raise RuntimeError('kbnet_b200: bad preallocated conv output')
    osn = out.stride(0) if n > 1 else f * ho * wo
    if engine is None and len(parsed) == 1 and parsed[0][2] == 0 and conv3x3_small_takes(ctot, f, k, stride, slope):
        t, sn, _ = parsed[0]
        with torch.cuda.device(dev):
            rc = lib.kbnet_conv3x3_small_forward(t.data_ptr(), sn, n, ctot, h_in, w_in, w.data_ptr(), f, 0 if slope is None else 1,
                                                 0.0 if slope is None else float(slope), out.data_ptr(), osn, _stream(dev))
        _lib.check(rc, 'kbnet_conv3x3_small_forward')
        return out
    arr = (_lib.KBNetConvSrc * len(parsed))()
    for i, (t, sn, mode) in enumerate(parsed):
        arr[i].ptr, arr[i].C, arr[i].H, arr[i].W, arr[i].stride_n, arr[i].mode = t.data_ptr(), t.shape[1], t.shape[2], t.shape[3], sn, mode
    pk = packed_weight(weight, [t.shape[1] for t, _, _ in parsed])
    with torch.cuda.device(dev):
        rc = lib.kbnet_conv2d_forward(arr, len(parsed), n, h_in, w_in, w.data_ptr(), pk.data_ptr() if pk is not None else None,
                                      f, k, stride,
                                      0 if slope is None else 1, 0.0 if slope is None else float(slope),
                                      out.data_ptr(), osn, _stream(dev))
    _lib.check(rc, 'kbnet_conv2d_forward')
    return out


def conv3x3_small_takes(c, f, k, stride, slope):
    """True when conv2d_fused routes a single full-resolution source of this shape to the few-channel mma.sync kernel."""
    from .modules import get_precision
    return get_precision() == 'bf16x3' and (slope is None or 0.0 <= float(slope) <= 1.0) and \
        bool(_lib.load().kbnet_conv3x3_small_supported(c, f, k, stride)) and \
        not (f <= 4 and c <= 32)        # <= 4 filters: the exact-fp32 CUDA-core kernel inside kbnet_conv2d_forward


def _upconv_packs(weight):
    """Four packed parity weights of an up-convolution (cached on the weight tensor like packed_weight)."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), 'up2x')
    cache = weight.__dict__.get('_kbnet_up2x') if hasattr(weight, '__dict__') else None
    if cache is not None and cache[0] == key:
        return cache[1]
    lib = _lib.load()
    wc = weight.detach().contiguous()
    f, c = wc.shape[0], wc.shape[1]
    dev = wc.device
    wpar = torch.empty((4, f, c, 2, 2), dtype=torch.float32, device=dev)
    arr = (ctypes.c_int * 1)(c)
    nbytes = lib.kbnet_conv_wpack_bytes(arr, 1, f, 2)
    packs = [torch.empty((nbytes,), dtype=torch.uint8, device=dev) for _ in range(4)]
    with torch.cuda.device(dev):
        _lib.check(lib.kbnet_upconv2x_prepare_weights(wc.data_ptr(), f, c, wpar.data_ptr(), _stream(dev)), 'kbnet_upconv2x_prepare_weights')
        for j in range(4):
            _lib.check(lib.kbnet_conv_pack_weights(wpar[j].data_ptr(), arr, 1, f, 2, packs[j].data_ptr(), _stream(dev)), 'kbnet_conv_pack_weights')
    try:
        weight.__dict__['_kbnet_up2x'] = (key, packs)
    except Exception:
        pass
    return packs


def upconv2x(x, weight, slope=0.2, out=None):
    """act(conv3x3(nearest_upsample_2x(x), weight)) as four source-resolution 2x2 convolutions (kbnet_upconv2x_forward);
    returns None when the tensor-core path cannot take the tensor (caller falls back to conv2d_fused with an 'up' source)."""
    from .modules import get_precision
    if get_precision() != 'bf16x3':
        return None
    lib = _lib.load()
    t, sn = _view4('source', x)
    n, c, hs, ws = t.shape
    f = weight.shape[0]
    if tuple(weight.shape[1:]) != (c, 3, 3) or not lib.kbnet_upconv2x_supported(t.data_ptr(), c, hs, ws, sn, f):
        return None
    dev = t.device
    _w(weight, dev)
    if out is None:
        out = torch.empty((n, f, 2 * hs, 2 * ws), dtype=torch.float32, device=dev)
    elif tuple(out.shape) != (n, f, 2 * hs, 2 * ws) or out.stride()[1:] != (4 * hs * ws, 2 * ws, 1) or out.dtype != torch.float32 \
            or out.device != dev:
        raise RuntimeError('kbnet_b200: bad preallocated upconv2x output')
    packs = _upconv_packs(weight)
    arr = (ctypes.c_void_p * 4)(*[pk.data_ptr() for pk in packs])
    osn = out.stride(0) if n > 1 else f * 4 * hs * ws
    with torch.cuda.device(dev):
        rc = lib.kbnet_upconv2x_forward(t.data_ptr(), n, c, hs, ws, sn, arr, f, 0 if slope is None else 1,
                                        0.0 if slope is None else float(slope), out.data_ptr(), osn, _stream(dev))
    _lib.check(rc, 'kbnet_upconv2x_forward')
    return out


def conv3x3_depth_head(x, weight, min_predict_depth, max_predict_depth):
    """min_d / (sigmoid(conv3x3(x, weight)) + min_d / max_d): `output0` fused with the head (reference src/networks.py:1989 +
    src/kbnet_model.py:181-184); None when the layer is not small enough for the fused kernel."""
    lib = _lib.load()
    t, sn = _view4('source', x)
    n, c, h, w_ = t.shape
    f = weight.shape[0]
    if tuple(weight.shape[1:]) != (c, 3, 3) or c > 32 or f > 4:
        return None
    dev = t.device
    wt = _w(weight, dev)
    y = torch.empty((n, f, h, w_), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.kbnet_conv3x3_depth_head_forward(t.data_ptr(), sn, n, c, h, w_, wt.data_ptr(), f, float(min_predict_depth),
                                                  float(max_predict_depth), y.data_ptr(), f * h * w_, _stream(dev))
    _lib.check(rc, 'kbnet_conv3x3_depth_head_forward')
    return y


def depth_head(x, min_predict_depth, max_predict_depth):
    """min_d / (sigmoid(x) + min_d / max_d)   (reference src/kbnet_model.py:181-184)."""
    lib = _lib.load()
    x = x.contiguous()
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = lib.kbnet_depth_head_forward(x.data_ptr(), x.numel(), float(min_predict_depth), float(max_predict_depth),
                                          y.data_ptr(), _stream(x.device))
    _lib.check(rc, 'kbnet_depth_head_forward')
    return y


def intrinsics_inverse(intrinsics, scale_x=1.0, scale_y=1.0):
    """(inverse(K), inverse(K * [[sx,1,sx],[1,sy,sy],[1,1,1]])) in one launch (reference src/networks.py:317-352)."""
    lib = _lib.load()
    k = intrinsics.contiguous()
    if not k.is_cuda or k.dtype != torch.float32 or k.dim() != 3 or tuple(k.shape[1:]) != (3, 3):
        raise RuntimeError('kbnet_b200: intrinsics must be a CUDA float32 [N,3,3] tensor')
    k0, k1 = torch.empty_like(k), torch.empty_like(k)
    with torch.cuda.device(k.device):
        rc = lib.kbnet_intrinsics_inverse(k.data_ptr(), k.shape[0], float(scale_x), float(scale_y), k0.data_ptr(), k1.data_ptr(),
                                          _stream(k.device))
    _lib.check(rc, 'kbnet_intrinsics_inverse')
    return k0, k1
