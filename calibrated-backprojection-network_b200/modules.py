"""Drop-in torch.nn.Module mirrors of the reference's two hot-path layers, backed by libkbnet_b200.so.

Same constructor / forward signatures and the same parameter tree (state_dict keys and shapes) as

  * networks.SparseToDensePool              reference src/networks.py:2078-2196
  * net_utils.CalibratedBackprojectionBlock reference src/net_utils.py:1269-1371

so the reference's kbnet_model.py runs unchanged after ``kbnet_b200.install()`` and its checkpoints load with
``strict=True``.  Forward is inference-only and CUDA-only: anything else raises (there is no CPU fallback).
"""
import ctypes

import torch

from . import _lib


# ---------------------------------------------------------------------------------------------------------
# parameter containers reproducing the reference key names
# ---------------------------------------------------------------------------------------------------------
def activation_slope(activation_func):
    """Map the reference's activation spec (net_utils.activation_func, src/net_utils.py:23-45; a string for S2D,
    an nn.Module instance for the KB block) to the negative slope the kernels apply."""
    if activation_func is None:
        raise NotImplementedError('kbnet_b200: linear (no) activation is not part of the hot path')
    if isinstance(activation_func, str):
        if 'linear' in activation_func:
            raise NotImplementedError('kbnet_b200: linear (no) activation is not part of the hot path')
        if 'leaky_relu' in activation_func:
            return 0.20  # src/net_utils.py:37
        if 'relu' in activation_func:
            return 0.0
        if 'elu' in activation_func or 'sigmoid' in activation_func:
            raise NotImplementedError('kbnet_b200: activation %r has no sm_100a kernel' % activation_func)
        raise ValueError('Unsupported activation function: {}'.format(activation_func))  # src/net_utils.py:45
    if isinstance(activation_func, torch.nn.LeakyReLU):
        return float(activation_func.negative_slope)
    if isinstance(activation_func, torch.nn.ReLU):
        return 0.0
    raise NotImplementedError('kbnet_b200: activation %r has no sm_100a kernel' % (activation_func,))


def _init_weight(weight, weight_initializer):
    # src/net_utils.py:95-105
    if weight_initializer == 'kaiming_normal':
        torch.nn.init.kaiming_normal_(weight)
    elif weight_initializer == 'xavier_normal':
        torch.nn.init.xavier_normal_(weight)
    elif weight_initializer == 'xavier_uniform':
        torch.nn.init.xavier_uniform_(weight)
    elif weight_initializer == 'kaiming_uniform':
        pass
    else:
        raise ValueError('Unsupported weight initializer: {}'.format(weight_initializer))


class ConvWeight(torch.nn.Module):
    """Holds ``conv.weight`` [out, in, k, k] exactly like net_utils.Conv2d (bias-free, src/net_utils.py:87-93)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, weight_initializer='kaiming_uniform'):
        super().__init__()
        self.conv = torch.nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                    padding=kernel_size // 2, bias=False)
        _init_weight(self.conv.weight, weight_initializer)
        self.kernel_size = kernel_size
        self.stride = stride

    @property
    def weight(self):
        return self.conv.weight


class ConvBlockWeight(torch.nn.Module):
    """Key layout of net_utils.VGGNetBlock with n_convolution == 1: ``conv_block.0.conv.weight``."""

    def __init__(self, in_channels, out_channels, n_convolution=1, stride=1, weight_initializer='kaiming_uniform'):
        super().__init__()
        if n_convolution != 1:
            raise NotImplementedError('kbnet_b200: n_convolution=%d (KBNet always uses 1, src/kbnet_model.py:107-109)'
                                      % n_convolution)
        self.conv_block = torch.nn.Sequential(ConvWeight(in_channels, out_channels, 3, stride, weight_initializer))

    @property
    def weight(self):
        return self.conv_block[0].conv.weight


def _check_input(name, t, ndim=4):
    if not isinstance(t, torch.Tensor):
        raise TypeError('kbnet_b200: %s must be a torch.Tensor' % name)
    if not t.is_cuda:
        raise RuntimeError('kbnet_b200: %s is on %s; this path runs on CUDA (sm_100a) only, there is no CPU fallback'
                           % (name, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError('kbnet_b200: %s must be float32 (got %s)' % (name, t.dtype))
    if t.dim() != ndim:
        raise RuntimeError('kbnet_b200: %s must be %d-D (got shape %s)' % (name, ndim, tuple(t.shape)))
    if torch.is_grad_enabled() and t.requires_grad:
        raise RuntimeError('kbnet_b200: inference-only kernels; call under torch.no_grad() (no backward is provided)')
    return t.contiguous()


def _view4(name, t):
    """Validate a 4-D input and return (tensor, elements-between-samples).  A channel slice of a wider NCHW buffer
    (dense C,H,W strides, larger batch stride) is passed through without a copy."""
    if isinstance(t, torch.Tensor) and t.dim() == 4 and t.is_cuda and t.dtype == torch.float32:
        c, h, w = t.shape[1:]
        if t.stride()[1:] == (h * w, w, 1) and (t.shape[0] == 1 or t.stride(0) >= c * h * w):
            if torch.is_grad_enabled() and t.requires_grad:
                raise RuntimeError('kbnet_b200: inference-only kernels; call under torch.no_grad()')
            return t, (t.stride(0) if t.shape[0] > 1 else c * h * w)
    t = _check_input(name, t)
    return t, t.shape[1] * t.shape[2] * t.shape[3]


def _w(param, device):
    w = param.detach()
    if w.device != device:
        raise RuntimeError('kbnet_b200: weights on %s but input on %s' % (w.device, device))
    return w.contiguous()


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# ---------------------------------------------------------------------------------------------------------
# arithmetic mode of the convolutions + packed-weight cache for the tensor-core engine
# ---------------------------------------------------------------------------------------------------------
_PRECISION = 'bf16x3'


def set_precision(mode):
    """'bf16x3' (default): convolutions run on the tcgen05 tensor cores with the 3-term bf16 operand split and fp32
    accumulation (~1e-5 relative, inside the 1e-3 parity budget).  'fp32': exact fp32 FFMA on the CUDA cores
    (reference-grade, slow) -- the on-GPU cross-check of the tensor path."""
    global _PRECISION
    if mode not in ('bf16x3', 'fp32'):
        raise ValueError('kbnet_b200: precision must be "bf16x3" or "fp32"')
    _PRECISION = mode


def get_precision():
    return _PRECISION


def packed_weight(weight, src_channels):
    """Device buffer with `weight` ([F, sum(src_channels), k, k]) re-laid for the tensor-core engine; None in fp32 mode or
    when the layer is outside the engine's envelope (the exact-fp32 kernel then takes it).
    The buffer is cached ON the weight tensor object (so it dies with it -- a cache keyed on data_ptr alone would be fooled
    by the allocator handing a freed weight's address to another layer) and is re-packed when the version counter or the
    storage changes: load_state_dict, optimizer steps, ``p.copy_()``, ``.to()``.  Writes THROUGH ``p.data`` (``p.data.mul_()``,
    ``p.data.copy_()``) do not bump the version counter: call ``invalidate_packs(module)`` after such surgery.
    Under torch.nn.DataParallel the replicas are fresh tensors every forward, so the pack is rebuilt per forward there
    (DataParallel re-broadcasts all weights each forward anyway, src/kbnet_model.py:408-415)."""
    if _PRECISION != 'bf16x3':
        return None
    key = (weight.data_ptr(), weight._version, tuple(int(c) for c in src_channels), tuple(weight.shape))
    cache = weight.__dict__.get('_kbnet_pack') if hasattr(weight, '__dict__') else None
    if cache is not None and cache[0] == key:
        return cache[1]
    lib = _lib.load()
    wc = weight.detach().contiguous()
    f, ctot, k, _ = wc.shape
    assert ctot == sum(src_channels), (ctot, src_channels)
    arr = (ctypes.c_int * len(src_channels))(*[int(c) for c in src_channels])
    nbytes = lib.kbnet_conv_wpack_bytes(arr, len(src_channels), f, k)
    if nbytes == 0:
        return None
    buf = torch.empty((nbytes,), dtype=torch.uint8, device=wc.device)
    with torch.cuda.device(wc.device):
        rc = lib.kbnet_conv_pack_weights(wc.data_ptr(), arr, len(src_channels), f, k, buf.data_ptr(), _stream(wc.device))
    _lib.check(rc, 'kbnet_conv_pack_weights')
    try:
        weight.__dict__['_kbnet_pack'] = (key, buf)
    except Exception:
        pass
    return buf


def invalidate_packs(module):
    """Drop every cached packed-weight image under `module` (needed only after writes through ``param.data``, which do
    not bump the tensor version the caches are keyed on)."""
    for m in module.modules():
        m.__dict__.pop('_kbnet_s2d_pack', None)
    for p in module.parameters():
        for k in ('_kbnet_pack', '_kbnet_up2x'):
            p.__dict__.pop(k, None)


# ---------------------------------------------------------------------------------------------------------
# operand format: activations stored the way the tensor-core engine consumes them (include/kbnet_b200.h)
# ---------------------------------------------------------------------------------------------------------
class OperandFormat(object):
    """An [N, C, H, W] activation held as two bf16 tensors [N, C/8, 4, H/2, W/2, 8] -- the hi and the lo half of every fp32 value
    (x ~= hi + lo, 16 mantissa bits), 8 consecutive channels of a pixel in one 16-byte entry, pixels split into the four parity
    planes (y & 1) * 2 + (x & 1) that a stride-2 convolution reads with unit stride.  ``even_only=True`` keeps plane 0 (the
    even pixels) only: [N, C/8, 1, H/2, W/2, 8].  Written by the producing kernel's epilogue and loaded by the next KB block with
    TMA straight into its MMA operand planes.  Exists only between this library's own layers; ``to_nchw()`` gives the fp32
    tensor back (exact for the bits the engine uses)."""

    def __init__(self, n, c, h, w, device, even_only=False):
        if c % 8 or h % 2 or w % 2:
            raise RuntimeError('kbnet_b200: operand format needs a multiple of 8 channels and even H, W (got %d, %d x %d)' % (c, h, w))
        self.shape = (n, c, h, w)
        self.even_only = bool(even_only)
        self.hi = torch.empty((n, c // 8, 1 if even_only else 4, h // 2, w // 2, 8), dtype=torch.bfloat16, device=device)
        self.lo = torch.empty_like(self.hi)
        self.device = self.hi.device

    @classmethod
    def from_nchw(cls, x, even_only=False):
        x, sn = _view4('x', x)
        n, c, h, w = x.shape
        of = cls(n, c, h, w, x.device, even_only)
        with torch.cuda.device(x.device):
            rc = _lib.load().kbnet_to_operand_format(x.data_ptr(), sn, n, c, h, w, of.hi.data_ptr(), of.lo.data_ptr(), int(of.even_only),
                                                     _stream(x.device))
        _lib.check(rc, 'kbnet_to_operand_format')
        return of

    def to_nchw(self):
        if self.even_only:
            raise RuntimeError('kbnet_b200: an even-pixel operand-format tensor holds a quarter of the pixels; no NCHW form')
        n, c, h, w = self.shape
        x = torch.empty((n, c, h, w), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.load().kbnet_from_operand_format(self.hi.data_ptr(), self.lo.data_ptr(), n, c, h, w, x.data_ptr(), c * h * w,
                                                       _stream(self.device))
        _lib.check(rc, 'kbnet_from_operand_format')
        return x


# ---------------------------------------------------------------------------------------------------------
# SparseToDensePool
# ---------------------------------------------------------------------------------------------------------
class SparseToDensePool(torch.nn.Module):
    """Mirror of networks.SparseToDensePool (reference src/networks.py:2078-2196): one fused sm_100a kernel."""

    def __init__(self,
                 input_channels,
                 min_pool_sizes=[3, 5, 7, 9],
                 max_pool_sizes=[3, 5, 7, 9],
                 n_filter=8,
                 n_convolution=3,
                 weight_initializer='kaiming_uniform',
                 activation_func='leaky_relu'):
        super().__init__()
        self.negative_slope = activation_slope(activation_func)
        self.min_pool_sizes = [s for s in min_pool_sizes if s > 1]   # src/networks.py:2112-2118
        self.max_pool_sizes = [s for s in max_pool_sizes if s > 1]
        self.len_pool_sizes = len(self.min_pool_sizes) + len(self.max_pool_sizes)
        self.input_channels = input_channels
        self.n_filter = n_filter
        self.n_convolution = n_convolution
        in_channels = self.len_pool_sizes
        pool_convs = []
        for _ in range(n_convolution):
            pool_convs.append(ConvWeight(in_channels, n_filter, 1, 1, weight_initializer))
            in_channels = n_filter
        self.pool_convs = torch.nn.Sequential(*pool_convs)
        self.conv = ConvWeight(n_filter + input_channels, n_filter, 3, 1, weight_initializer)

    def forward(self, x, return_pyramid=False):
        x = _check_input('x', x)
        for s in self.min_pool_sizes + self.max_pool_sizes:
            if s % 2 == 0:
                # the reference fails at the torch.cat in src/networks.py:2194 (H+1 rows); fail the same way, earlier
                raise RuntimeError('kbnet_b200: even pool size %d changes the spatial size (reference raises in torch.cat)' % s)
        n, c, h, w = x.shape
        if c != self.input_channels:
            raise RuntimeError('kbnet_b200: expected %d input channels, got %d' % (self.input_channels, c))
        lib = _lib.load()
        dev = x.device
        out = torch.empty((n, self.n_filter, h, w), dtype=torch.float32, device=dev)
        pyr = torch.empty((n, self.len_pool_sizes, h, w), dtype=torch.float32, device=dev) if return_pyramid else None
        w1 = [_w(m.conv.weight, dev) for m in self.pool_convs]
        w3 = _w(self.conv.conv.weight, dev)
        n_min, n_max = len(self.min_pool_sizes), len(self.max_pool_sizes)
        mins = (ctypes.c_int * max(n_min, 1))(*self.min_pool_sizes)
        maxs = (ctypes.c_int * max(n_max, 1))(*self.max_pool_sizes)
        w1p = (ctypes.c_void_p * max(len(w1), 1))(*[t.data_ptr() for t in w1])
        # packed weight image for the fused kernel, cached until a weight changes (value, storage or device)
        key = tuple((t.data_ptr(), t._version) for t in w1 + [w3]) + (c, self.negative_slope)
        cache = self.__dict__.get('_kbnet_s2d_pack')
        if cache is None or cache[0] != key:
            nbytes = lib.kbnet_s2d_wpack_bytes(c, mins, n_min, maxs, n_max, len(w1), self.n_filter, self.negative_slope)
            pack = None
            if nbytes:
                pack = torch.empty((int(nbytes),), dtype=torch.uint8, device=dev)
                with torch.cuda.device(dev):
                    rc = lib.kbnet_s2d_pack_weights(w1p, len(w1), w3.data_ptr(), c, mins, n_min, maxs, n_max, pack.data_ptr(), _stream(dev))
                _lib.check(rc, 'kbnet_s2d_pack_weights')
            cache = (key, pack)
            self.__dict__['_kbnet_s2d_pack'] = cache
        pack = cache[1]
        with torch.cuda.device(dev):
            rc = lib.kbnet_s2d_forward(x.data_ptr(), n, c, h, w, mins, n_min, maxs, n_max, w1p, len(w1), self.n_filter,
                                       w3.data_ptr(), pack.data_ptr() if pack is not None else None, self.negative_slope, out.data_ptr(),
                                       pyr.data_ptr() if pyr is not None else None, _stream(dev))
        _lib.check(rc, 'kbnet_s2d_forward')
        return (out, pyr) if return_pyramid else out


# ---------------------------------------------------------------------------------------------------------
# CalibratedBackprojectionBlock
# ---------------------------------------------------------------------------------------------------------
class CalibratedBackprojectionBlock(torch.nn.Module):
    """Mirror of net_utils.CalibratedBackprojectionBlock (reference src/net_utils.py:1269-1371).

    ``forward(image, depth, coordinates, fused=None)`` accepts what the reference passes.  In addition
    ``coordinates`` may be an ``[N, 3, 3]`` tensor holding K^-1 per sample: the kernels then generate the
    per-pixel coordinates themselves and the [N,3,H,W] tensor never exists (fast path used by KBNetEncoder here).
    """

    def __init__(self,
                 in_channels_image,
                 in_channels_depth,
                 in_channels_fused,
                 n_filter_image=48,
                 n_filter_depth=16,
                 n_filter_fused=48,
                 n_convolution_image=1,
                 n_convolution_depth=1,
                 n_convolution_fused=1,
                 weight_initializer='kaiming_uniform',
                 activation_func=torch.nn.LeakyReLU(negative_slope=0.10, inplace=True)):
        super().__init__()
        self.negative_slope = activation_slope(activation_func)
        self.in_channels_image = in_channels_image
        self.in_channels_depth = in_channels_depth
        self.in_channels_fused = in_channels_fused  # already includes the image channels (src/networks.py:102,142)
        self.n_filter_image, self.n_filter_depth, self.n_filter_fused = n_filter_image, n_filter_depth, n_filter_fused
        self.conv_image = ConvBlockWeight(in_channels_image, n_filter_image, n_convolution_image, 2, weight_initializer)
        self.conv_depth = ConvBlockWeight(in_channels_depth + 3, n_filter_depth, n_convolution_depth, 2, weight_initializer)
        self.proj_depth = ConvWeight(in_channels_depth, 1, 1, 1, weight_initializer)
        # n_convolution_fused is ignored by the reference too (src/net_utils.py:1335-1341)
        self.conv_fused = ConvWeight(in_channels_fused + 3, n_filter_fused, 1, 2, weight_initializer)

    def operand_format_supported(self, h, w, c_f=None):
        """True when a block input of spatial size (h, w) runs as the fused one-launch program in the tensor-core mode, i.e.
        ``image`` / ``out_image`` / ``fused_of`` / ``out_fused_of`` may be OperandFormat tensors."""
        c_f = self.in_channels_fused - self.in_channels_image if c_f is None else c_f
        return _PRECISION == 'bf16x3' and bool(_lib.load().kbnet_kb_operand_format_supported(
            self.in_channels_image, self.in_channels_depth, c_f, h, w, self.n_filter_image, self.n_filter_depth, self.n_filter_fused))

    def forward(self, image, depth, coordinates, fused=None, out_image=None, out_depth=None, out_fused=None, fused_of=None,
                out_fused_of=None, need_image=True):
        """The reference's ``forward(image, depth, coordinates, fused=None)`` (src/net_utils.py:1343-1371).  Extras used by
        KBNetEncoder here: ``out_*`` preallocated outputs (channel slices of the skip buffers); ``image`` and ``out_image`` may
        be OperandFormat tensors (then conv_image is returned as that OperandFormat and no NCHW copy of it is written);
        ``fused_of`` = `fused` as an even-pixel OperandFormat (`fused` itself may then be None); ``out_fused_of`` = an even-pixel
        OperandFormat that receives the even pixels of conv_fused in addition to the fp32 output; ``need_image=False`` does not compute conv_image at all (returned
        as None) -- for a caller that, like KBNetEncoder for its last block (src/networks.py:476-523), never reads it."""
        image_of = image if isinstance(image, OperandFormat) else None
        out_image_of = out_image if isinstance(out_image, OperandFormat) else None
        if out_image_of is not None:
            out_image = None
        depth, depth_sn = _view4('depth', depth)
        if image_of is not None:
            if image_of.even_only:
                raise RuntimeError('kbnet_b200: image needs the full (four-plane) operand format')
            n, c_i, h, w = image_of.shape
            image, image_sn = None, c_i * h * w
            if image_of.device != depth.device:
                raise RuntimeError('kbnet_b200: image (operand format) on %s but depth on %s' % (image_of.device, depth.device))
        else:
            image, image_sn = _view4('image', image)
            n, c_i, h, w = image.shape
        c_d = depth.shape[1]
        dev = depth.device
        if depth.shape[0] != n or tuple(depth.shape[-2:]) != (h, w):
            raise RuntimeError('kbnet_b200: depth shape %s does not match image %s' % (tuple(depth.shape), (n, c_i, h, w)))
        c_f, fused_sn = 0, 0
        if fused_of is not None:
            if tuple(fused_of.shape[0:1] + fused_of.shape[2:]) != (n, h, w) or fused_of.device != dev or not fused_of.even_only:
                raise RuntimeError('kbnet_b200: fused (operand format) shape %s does not match image' % (tuple(fused_of.shape),))
            c_f = fused_of.shape[1]
            fused = None
        elif fused is not None:
            fused, fused_sn = _view4('fused', fused)
            c_f = fused.shape[1]
            if fused.shape[0] != n or tuple(fused.shape[-2:]) != (h, w):
                raise RuntimeError('kbnet_b200: fused shape %s does not match image' % (tuple(fused.shape),))
        if c_i != self.in_channels_image or c_d != self.in_channels_depth or c_i + c_f != self.in_channels_fused:
            raise RuntimeError('kbnet_b200: channel mismatch: image %d (expected %d), depth %d (expected %d), '
                               'image+fused %d (expected %d)' % (c_i, self.in_channels_image, c_d, self.in_channels_depth,
                                                                c_i + c_f, self.in_channels_fused))
        coords_ptr = kinv_ptr = None
        if coordinates.dim() == 3:
            kinv = _check_input('coordinates (K^-1)', coordinates, 3)
            if tuple(kinv.shape) != (n, 3, 3):
                raise RuntimeError('kbnet_b200: K^-1 must be [N,3,3], got %s' % (tuple(kinv.shape),))
            kinv_ptr = kinv.data_ptr()
        else:
            coordinates = _check_input('coordinates', coordinates)
            if tuple(coordinates.shape) != (n, 3, h, w):
                raise RuntimeError('kbnet_b200: coordinates must be [N,3,H,W], got %s' % (tuple(coordinates.shape),))
            coords_ptr = coordinates.data_ptr()
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        f_i, f_d, f_f = self.n_filter_image, self.n_filter_depth, self.n_filter_fused

        def _out(t, f):
            if t is None:
                return torch.empty((n, f, ho, wo), dtype=torch.float32, device=dev)
            # a channel slice of a wider [N, C, ho, wo] buffer (lets the encoder skip the torch.cat of the skips)
            if tuple(t.shape) != (n, f, ho, wo) or t.dtype != torch.float32 or t.device != dev or \
                    t.stride()[1:] != (ho * wo, wo, 1):
                raise RuntimeError('kbnet_b200: preallocated output must be a [N,%d,%d,%d] float32 channel slice' % (f, ho, wo))
            return t

        o_i = _out(out_image, f_i) if (out_image_of is None and need_image) else None
        o_d, o_f = _out(out_depth, f_d), _out(out_fused, f_f)
        lib = _lib.load()
        of = None
        if image_of is not None or out_image_of is not None or fused_of is not None or out_fused_of is not None:
            for t, f, even in ((out_image_of, f_i, False), (out_fused_of, f_f, True)):
                if t is not None and (tuple(t.shape) != (n, f, ho, wo) or t.device != dev or t.even_only != even):
                    raise RuntimeError('kbnet_b200: operand-format output must be [N,%d,%d,%d] on %s (conv_image: full, conv_fused: '
                                       'even pixels only)' % (f, ho, wo, dev))
            of = _lib.KBNetKbOperandFormat()
            for name, t in (('image', image_of), ('fused', fused_of), ('out_image', out_image_of), ('out_fused', out_fused_of)):
                if t is not None:
                    setattr(of, name + '_hi', t.hi.data_ptr())
                    setattr(of, name + '_lo', t.lo.data_ptr())
        ws_bytes = lib.kbnet_kb_workspace_bytes(n, c_i, c_d, c_f, h, w, f_i, f_d, f_f)
        # scratch for the un-fused path (z at even pixels); kept on the module and re-used: no allocation per call
        ws = self.__dict__.get('_kbnet_ws')
        if ws is None or ws.device != dev or ws.numel() < max(int(ws_bytes), 256):
            ws = torch.empty((max(int(ws_bytes), 256),), dtype=torch.uint8, device=dev)
            self.__dict__['_kbnet_ws'] = ws
        wi, wd = _w(self.conv_image.weight, dev), _w(self.conv_depth.weight, dev)
        wp, wf = _w(self.proj_depth.weight, dev), _w(self.conv_fused.weight, dev)
        sn = lambda t: t.stride(0) if n > 1 else t.shape[1] * ho * wo
        pk_i = packed_weight(self.conv_image.weight, [c_i])
        pk_d = packed_weight(self.conv_depth.weight, [c_d, 3])
        pk_f = packed_weight(self.conv_fused.weight, [c_i, 3, c_f] if c_f else [c_i, 3])
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(dev):
            rc = lib.kbnet_kb_forward_of(ptr(image), image_sn, depth.data_ptr(), depth_sn,
                                         ptr(fused), fused_sn, coords_ptr, kinv_ptr, n, c_i, c_d, c_f, h, w,
                                         wi.data_ptr(), f_i, wd.data_ptr(), f_d, wp.data_ptr(), wf.data_ptr(), f_f,
                                         ptr(pk_i), ptr(pk_d), ptr(pk_f), self.negative_slope,
                                         ptr(o_i), sn(o_i) if o_i is not None else f_i * ho * wo, o_d.data_ptr(), sn(o_d), o_f.data_ptr(), sn(o_f),
                                         ws.data_ptr(), ws_bytes, ctypes.byref(of) if of is not None else None, _stream(dev))
        _lib.check(rc, 'kbnet_kb_forward')
        return (o_i if out_image_of is None else out_image_of), o_d, o_f


# ---------------------------------------------------------------------------------------------------------
# OutlierRemoval (pre-processing step immediately before the path)
# ---------------------------------------------------------------------------------------------------------
class OutlierRemoval(object):
    """Mirror of net_utils.OutlierRemoval (reference src/net_utils.py:1745-1806): same constructor and
    ``remove_outliers(sparse_depth, validity_map)``; ``validity_map=None`` derives it as src/kbnet.py:899-902 does."""

    def __init__(self, kernel_size=7, threshold=1.5):
        self.kernel_size = kernel_size
        self.threshold = threshold

    def remove_outliers(self, sparse_depth, validity_map=None):
        sparse_depth = _check_input('sparse_depth', sparse_depth)
        n, c, h, w = sparse_depth.shape
        if c != 1:
            raise RuntimeError('kbnet_b200: sparse_depth must be [N,1,H,W]')
        if validity_map is not None:
            validity_map = _check_input('validity_map', validity_map)
            if validity_map.shape != sparse_depth.shape:
                raise RuntimeError('kbnet_b200: validity_map shape %s does not match sparse_depth' % (tuple(validity_map.shape),))
        dev = sparse_depth.device
        out_d, out_v = torch.empty_like(sparse_depth), torch.empty_like(sparse_depth)
        ws = torch.empty((4,), dtype=torch.int32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            rc = lib.kbnet_outlier_removal_forward(sparse_depth.data_ptr(), validity_map.data_ptr() if validity_map is not None else None,
                                                   n, h, w, int(self.kernel_size), float(self.threshold), out_d.data_ptr(), out_v.data_ptr(),
                                                   ws.data_ptr(), _stream(dev))
        _lib.check(rc, 'kbnet_outlier_removal_forward')
        return out_d, out_v


# ---------------------------------------------------------------------------------------------------------
# install(): swap the two classes into the reference's flat modules
# ---------------------------------------------------------------------------------------------------------
def install(networks_module=None, net_utils_module=None):
    """Monkey-patch ``networks.SparseToDensePool`` and ``net_utils.CalibratedBackprojectionBlock`` (the reference
    resolves them by flat module name: src/kbnet_model.py:18,87; src/networks.py:18,104) so that
    ``kbnet_model.KBNetModel(...)`` builds the B200 layers.  Returns the originals for un-patching."""
    import importlib
    networks_module = networks_module or importlib.import_module('networks')
    net_utils_module = net_utils_module or importlib.import_module('net_utils')
    originals = (networks_module.SparseToDensePool, net_utils_module.CalibratedBackprojectionBlock)
    networks_module.SparseToDensePool = SparseToDensePool
    net_utils_module.CalibratedBackprojectionBlock = CalibratedBackprojectionBlock
    if hasattr(net_utils_module, 'OutlierRemoval'):
        net_utils_module.OutlierRemoval = OutlierRemoval   # src/kbnet.py:134,807 construct it through net_utils
    return originals
