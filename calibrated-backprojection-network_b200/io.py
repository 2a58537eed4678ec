"""Depth / intrinsics I/O in the reference's on-disk formats (reference src/data_utils.py:88-167,
src/datasets.py:275): depth maps are 16-bit PNGs holding depth * 256, intrinsics are 3x3 float .npy files.

The PNG / NPY containers are host work (PIL and numpy, exactly the calls the reference makes); the element-wise part
(/ 256, validity map, * 256 + truncation) runs on the device through the C ABI (kbnet_depth_u16_decode / _encode), so a
sparse depth map crosses PCIe as 2 bytes per pixel and arrives as the (depth, validity) pair KBNetModel.forward takes.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .modules import _stream


def read_depth_png_u16(path):
    """The raw 16-bit samples of a depth PNG as a uint16 [H, W] array (what ``np.array(Image.open(path))`` holds)."""
    from PIL import Image
    a = np.array(Image.open(path))
    if a.ndim != 2:
        raise ValueError('kbnet_b200: %s is not a single-channel 16-bit depth PNG (shape %s)' % (path, a.shape))
    if a.dtype != np.uint16:
        if a.min() < 0 or a.max() > 65535:
            raise ValueError('kbnet_b200: %s holds values outside the 16-bit range' % path)
        a = a.astype(np.uint16)
    return a


def decode_depth_u16(z16, with_validity=True):
    """uint16 tensor [...] on a CUDA device -> (depth, validity) float32 tensors of the same shape (validity None if not asked for):
    depth = z / 256 with non-positive values set to 0, validity = 1 where depth > 0 (src/data_utils.py:100-107, 137-141)."""
    if not isinstance(z16, torch.Tensor) or not z16.is_cuda or z16.dtype != torch.uint16:
        raise RuntimeError('kbnet_b200: decode_depth_u16 needs a CUDA uint16 tensor (there is no CPU fallback)')
    z16 = z16.contiguous()
    dev = z16.device
    depth = torch.empty(z16.shape, dtype=torch.float32, device=dev)
    valid = torch.empty(z16.shape, dtype=torch.float32, device=dev) if with_validity else None
    lib = _lib.load()
    with torch.cuda.device(dev):
        rc = lib.kbnet_depth_u16_decode(z16.data_ptr(), z16.numel(), depth.data_ptr(), valid.data_ptr() if valid is not None else None,
                                        _stream(dev))
    _lib.check(rc, 'kbnet_depth_u16_decode')
    return depth, valid


def load_depth_with_validity_map(path, device, data_format='HW'):
    """Mirror of data_utils.load_depth_with_validity_map (src/data_utils.py:88-121) returning CUDA tensors."""
    z16 = torch.from_numpy(read_depth_png_u16(path)).to(device)
    depth, valid = decode_depth_u16(z16, True)
    return _format(depth, data_format), _format(valid, data_format)


def load_depth(path, device, data_format='HW'):
    """Mirror of data_utils.load_depth (src/data_utils.py:123-152) returning a CUDA tensor."""
    z16 = torch.from_numpy(read_depth_png_u16(path)).to(device)
    return _format(decode_depth_u16(z16, False)[0], data_format)


def encode_depth_u32(depth):
    """float32 CUDA tensor -> uint32 host array of depth * 256 (C truncation), the samples save_depth writes (src/data_utils.py:165)."""
    if not isinstance(depth, torch.Tensor) or not depth.is_cuda or depth.dtype != torch.float32:
        raise RuntimeError('kbnet_b200: encode_depth_u32 needs a CUDA float32 tensor (there is no CPU fallback)')
    depth = depth.contiguous()
    dev = depth.device
    out = torch.empty(depth.shape, dtype=torch.int32, device=dev)     # the bits are uint32
    lib = _lib.load()
    with torch.cuda.device(dev):
        rc = lib.kbnet_depth_u16_encode(depth.data_ptr(), depth.numel(), out.data_ptr(), _stream(dev))
    _lib.check(rc, 'kbnet_depth_u16_encode')
    return out.cpu().numpy().view(np.uint32)


def save_depth(z, path):
    """Mirror of data_utils.save_depth (src/data_utils.py:154-167): a [H, W] depth map -> 16-bit PNG of depth * 256."""
    from PIL import Image
    a = encode_depth_u32(z.reshape(z.shape[-2], z.shape[-1]))
    Image.fromarray(a, mode='I').save(path)


def load_intrinsics(path, device):
    """3x3 intrinsics from a .npy file as a float32 [3, 3] tensor (src/datasets.py:275)."""
    k = np.load(path).astype(np.float32)
    if k.shape != (3, 3):
        raise ValueError('kbnet_b200: %s holds a %s array, expected 3x3 intrinsics' % (path, k.shape))
    return torch.from_numpy(k).to(device)


def _format(t, data_format):
    if data_format == 'HW':
        return t
    if data_format == 'CHW':
        return t.unsqueeze(0)
    if data_format == 'HWC':
        return t.unsqueeze(-1)
    raise ValueError('Unsupported data format: {}'.format(data_format))     # src/data_utils.py:119
