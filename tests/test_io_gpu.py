"""Depth I/O in the reference's on-disk format (16-bit PNG of depth * 256, src/data_utils.py:88-167): the device kernels are
bit-exact against the reference's numpy arithmetic, files written by one side are read identically by the other."""
import os

import numpy as np
import pytest
import torch

from oracle import refload
import kbnet_b200

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


def _ref_load(a16):
    z = a16.astype(np.float32) / 256.0                 # src/data_utils.py:137-141
    z[z <= 0] = 0.0
    v = z.astype(np.float32)                           # :104-107
    v[z > 0] = 1.0
    return z, v


def test_decode_encode_bit_exact_vs_numpy():
    rng = np.random.RandomState(0)
    a16 = (rng.randint(0, 65536, size=(3, 37, 53)) * (rng.rand(3, 37, 53) < 0.3)).astype(np.uint16)
    a16[0, 0, :4] = [0, 1, 255, 65535]
    z, v = _ref_load(a16)
    depth, valid = kbnet_b200.io.decode_depth_u16(torch.from_numpy(a16).to(DEV))
    assert np.array_equal(depth.cpu().numpy(), z) and np.array_equal(valid.cpu().numpy(), v)
    zz = (rng.rand(64, 80).astype(np.float32) * 300.0) * (rng.rand(64, 80) < 0.5)
    want = np.uint32(zz * 256.0)                       # src/data_utils.py:165
    got = kbnet_b200.io.encode_depth_u32(torch.from_numpy(zz).to(DEV))
    assert got.dtype == np.uint32 and np.array_equal(got, want)


def test_png_round_trip_and_reference_interop(tmp_path):
    rng = np.random.RandomState(1)
    z = ((rng.rand(48, 64) * 80.0 + 1.0) * (rng.rand(48, 64) < 0.05)).astype(np.float32)
    z = np.floor(z * 256.0) / 256.0                    # representable on disk
    p1 = os.path.join(tmp_path, 'ours.png')
    kbnet_b200.io.save_depth(torch.from_numpy(z).to(DEV), p1)
    d, v = kbnet_b200.io.load_depth_with_validity_map(p1, DEV, data_format='CHW')
    assert tuple(d.shape) == (1, 48, 64) and np.array_equal(d[0].cpu().numpy(), z)
    assert np.array_equal(v[0].cpu().numpy(), (z > 0).astype(np.float32))
    assert np.array_equal(kbnet_b200.io.load_depth(p1, DEV).cpu().numpy(), z)
    k = np.array([[721.5, 0, 609.6], [0, 721.5, 172.9], [0, 0, 1]], dtype=np.float64)
    np.save(os.path.join(tmp_path, 'k.npy'), k)
    assert np.array_equal(kbnet_b200.io.load_intrinsics(os.path.join(tmp_path, 'k.npy'), DEV).cpu().numpy(), k.astype(np.float32))
    mods = refload.load()
    if mods is None:
        pytest.skip('no copy of the reference on this box: interop half skipped')
    import data_utils                                   # the unmodified reference module (baseline/_ref/src)
    zr, vr = data_utils.load_depth_with_validity_map(p1)
    assert np.array_equal(zr, z) and np.array_equal(vr, (z > 0).astype(np.float32))
    p2 = os.path.join(tmp_path, 'ref.png')
    data_utils.save_depth(z, p2)
    assert np.array_equal(kbnet_b200.io.load_depth(p2, DEV).cpu().numpy(), data_utils.load_depth(p2))


def test_io_rejects_cpu_tensors():
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        kbnet_b200.io.decode_depth_u16(torch.zeros(4, 4, dtype=torch.uint16))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        kbnet_b200.io.encode_depth_u32(torch.zeros(4, 4))
