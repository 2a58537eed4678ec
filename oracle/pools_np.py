"""numpy restatement of the SparseToDensePool pool pyramid -- TEST INFRASTRUCTURE ONLY.

Independent of ATen's max_pool2d: windows are materialised with
``sliding_window_view`` over an explicitly padded array and reduced with the
selection rule spelled out.  Follows src/networks.py:2172-2189 of the reference
(see oracle/kbnet_oracle.py for the parity-pinning statement).  Selection only, so
results are bit-exact by construction; tests compare with ``array_equal``.

Semantics (measured on the reference):
  * a pixel is "empty" iff z == 0 (so -0.0 is empty too);
  * min pool = min over the window of (999 if empty else z); a result of exactly
    999 is replaced by 0;  out-of-image pixels are ignored (-inf padding of the
    negated map);
  * max pool = max over the in-image part of the window (-inf padding);
  * NaN anywhere in the window -> NaN (ATen's max pool propagates NaN).
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view


def _window_reduce(a, s, pad_value, fn):
    r = s // 2
    ap = np.pad(a, ((0, 0), (r, r), (r, r)), mode='constant', constant_values=pad_value)
    win = sliding_window_view(ap, (s, s), axis=(1, 2))  # [N, H, W, s, s]
    return fn(win, axis=(-2, -1))                       # np.min / np.max propagate NaN


def pool_pyramid_np(z, min_pool_sizes, max_pool_sizes):
    """z: float32 [N, H, W] -> float32 [N, n_min + n_max, H, W]."""
    z = np.asarray(z, dtype=np.float32)
    out = []
    for s in [s for s in min_pool_sizes if s > 1]:
        f = np.where(z == 0, np.float32(999.0), z).astype(np.float32)
        m = _window_reduce(f, s, np.float32(np.inf), np.min)
        out.append(np.where(m == np.float32(999.0), np.float32(0.0), m))
    for s in [s for s in max_pool_sizes if s > 1]:
        out.append(_window_reduce(z, s, np.float32(-np.inf), np.max))
    return np.stack(out, axis=1).astype(np.float32)
