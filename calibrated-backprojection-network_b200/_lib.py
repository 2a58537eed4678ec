"""ctypes binding of libkbnet_b200.so (C ABI declared in include/kbnet_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, the caller gets an exception.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libkbnet_b200.so')

EXPORTS = ('kbnet_version', 'kbnet_last_error', 'kbnet_launch_count', 'kbnet_tma_engine_launch_count', 'kbnet_set_autotune', 'kbnet_upconv2x_supported', 'kbnet_upconv2x_prepare_weights', 'kbnet_upconv2x_forward', 'kbnet_outlier_removal_forward', 'kbnet_conv3x3_depth_head_forward', 'kbnet_conv3x3_small_supported', 'kbnet_conv3x3_small_forward', 'kbnet_depth_u16_decode', 'kbnet_depth_u16_encode', 'kbnet_s2d_forward', 'kbnet_s2d_wpack_bytes', 'kbnet_s2d_pack_weights', 'kbnet_kb_workspace_bytes', 'kbnet_kb_forward', 'kbnet_kb_forward_of', 'kbnet_kb_operand_format_supported',
           'kbnet_operand_format_bytes', 'kbnet_to_operand_format', 'kbnet_from_operand_format', 'kbnet_conv3x3_small_forward_of',
           'kbnet_conv2d_forward', 'kbnet_depth_head_forward', 'kbnet_conv_wpack_bytes', 'kbnet_conv_pack_weights', 'kbnet_intrinsics_inverse')


class KBNetKbOperandFormat(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in ('image_hi', 'image_lo', 'fused_hi', 'fused_lo', 'out_image_hi', 'out_image_lo',
                                               'out_fused_hi', 'out_fused_lo')]


class KBNetConvSrc(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p), ('C', ctypes.c_int), ('H', ctypes.c_int), ('W', ctypes.c_int),
                ('stride_n', ctypes.c_longlong), ('mode', ctypes.c_int)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('kbnet_b200: %s not found -- build it with `python -c "import __graft_entry__ as g; g.build()"` '
                           '(there is no CPU / PyTorch fallback for this path)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i, f, ll, sz = c.c_void_p, c.c_int, c.c_float, c.c_longlong, c.c_size_t
    lib.kbnet_version.restype = i
    lib.kbnet_last_error.restype = c.c_char_p
    lib.kbnet_launch_count.restype = c.c_ulonglong
    lib.kbnet_tma_engine_launch_count.restype = c.c_ulonglong
    lib.kbnet_set_autotune.restype = None
    lib.kbnet_set_autotune.argtypes = [i]
    lib.kbnet_s2d_forward.restype = i
    lib.kbnet_s2d_forward.argtypes = [vp, i, i, i, i, c.POINTER(i), i, c.POINTER(i), i, c.POINTER(vp), i, i, vp, vp, f, vp, vp, vp]
    lib.kbnet_s2d_wpack_bytes.restype = sz
    lib.kbnet_s2d_wpack_bytes.argtypes = [i, c.POINTER(i), i, c.POINTER(i), i, i, i, f]
    lib.kbnet_s2d_pack_weights.restype = i
    lib.kbnet_s2d_pack_weights.argtypes = [c.POINTER(vp), i, vp, i, c.POINTER(i), i, c.POINTER(i), i, vp, vp]
    lib.kbnet_kb_workspace_bytes.restype = sz
    lib.kbnet_kb_workspace_bytes.argtypes = [i] * 9
    lib.kbnet_kb_forward.restype = i
    lib.kbnet_conv_wpack_bytes.restype = sz
    lib.kbnet_conv_wpack_bytes.argtypes = [c.POINTER(i), i, i, i]
    lib.kbnet_conv_pack_weights.restype = i
    lib.kbnet_conv_pack_weights.argtypes = [vp, c.POINTER(i), i, i, i, vp, vp]
    lib.kbnet_kb_forward.argtypes = [vp, ll, vp, ll, vp, ll, vp, vp, i, i, i, i, i, i, vp, i, vp, i, vp, vp, i, vp, vp, vp, f,
                                     vp, ll, vp, ll, vp, ll, vp, sz, vp]
    lib.kbnet_kb_forward_of.restype = i
    lib.kbnet_kb_forward_of.argtypes = lib.kbnet_kb_forward.argtypes[:-1] + [c.POINTER(KBNetKbOperandFormat), vp]
    lib.kbnet_kb_operand_format_supported.restype = i
    lib.kbnet_kb_operand_format_supported.argtypes = [i] * 8
    lib.kbnet_operand_format_bytes.restype = sz
    lib.kbnet_operand_format_bytes.argtypes = [i] * 5
    lib.kbnet_to_operand_format.restype = i
    lib.kbnet_to_operand_format.argtypes = [vp, ll, i, i, i, i, vp, vp, i, vp]
    lib.kbnet_from_operand_format.restype = i
    lib.kbnet_from_operand_format.argtypes = [vp, vp, i, i, i, i, vp, ll, vp]
    lib.kbnet_conv2d_forward.restype = i
    lib.kbnet_conv2d_forward.argtypes = [c.POINTER(KBNetConvSrc), i, i, i, i, vp, vp, i, i, i, i, f, vp, ll, vp]
    lib.kbnet_upconv2x_supported.restype = i
    lib.kbnet_upconv2x_supported.argtypes = [vp, i, i, i, ll, i]
    lib.kbnet_upconv2x_prepare_weights.restype = i
    lib.kbnet_upconv2x_prepare_weights.argtypes = [vp, i, i, vp, vp]
    lib.kbnet_upconv2x_forward.restype = i
    lib.kbnet_upconv2x_forward.argtypes = [vp, i, i, i, i, ll, c.POINTER(vp), i, i, f, vp, ll, vp]
    lib.kbnet_intrinsics_inverse.restype = i
    lib.kbnet_intrinsics_inverse.argtypes = [vp, i, f, f, vp, vp, vp]
    lib.kbnet_outlier_removal_forward.restype = i
    lib.kbnet_outlier_removal_forward.argtypes = [vp, vp, i, i, i, i, f, vp, vp, vp, vp]
    lib.kbnet_conv3x3_depth_head_forward.restype = i
    lib.kbnet_conv3x3_depth_head_forward.argtypes = [vp, ll, i, i, i, i, vp, i, c.c_double, c.c_double, vp, ll, vp]
    lib.kbnet_conv3x3_small_supported.restype = i
    lib.kbnet_conv3x3_small_supported.argtypes = [i, i, i, i]
    lib.kbnet_conv3x3_small_forward.restype = i
    lib.kbnet_conv3x3_small_forward.argtypes = [vp, ll, i, i, i, i, vp, i, i, f, vp, ll, vp]
    lib.kbnet_conv3x3_small_forward_of.restype = i
    lib.kbnet_conv3x3_small_forward_of.argtypes = [vp, ll, i, i, i, i, vp, i, i, f, vp, ll, vp, vp, vp]
    lib.kbnet_depth_u16_decode.restype = i
    lib.kbnet_depth_u16_decode.argtypes = [vp, ll, vp, vp, vp]
    lib.kbnet_depth_u16_encode.restype = i
    lib.kbnet_depth_u16_encode.argtypes = [vp, ll, vp, vp]
    lib.kbnet_depth_head_forward.restype = i
    lib.kbnet_depth_head_forward.argtypes = [vp, ll, c.c_double, c.c_double, vp, vp]
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().kbnet_last_error()
        raise RuntimeError('kbnet_b200 %s failed (code %d): %s' % (what, rc, msg.decode() if msg else ''))
