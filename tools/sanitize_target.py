"""One launch of every hot kernel at small sizes, for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
fused S2D (KITTI and VOID pool sets), fused KB block (K^-1 and coordinates paths, operand-format hand-over, no conv_image),
plain conv, few-channel conv with operand-format output, up-convolution, output head.
Run:  compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import kbnet_b200  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
gen = torch.Generator().manual_seed(0)
lib = kbnet_b200._lib.load()
n0 = lib.kbnet_launch_count()
for mins, maxs, hw in (([5, 7, 9, 11, 13], [15, 17], (40, 136)), ([15, 17], [23, 27, 29], (48, 64))):
    m = kbnet_b200.SparseToDensePool(2, mins, maxs, 8, 3, 'xavier_normal', 'leaky_relu').to(dev)
    d = torch.rand(2, 1, *hw, generator=gen) * (torch.rand(2, 1, *hw, generator=gen) < 0.05)
    x = torch.cat([d, (d > 0).float()], 1).to(dev)
    y, p = m(x, return_pyramid=True)
    y2 = m(x)
    assert torch.isfinite(y).all() and torch.equal(y, y2)
kb = kbnet_b200.CalibratedBackprojectionBlock(16, 16, 32, 32, 16, 32, 1, 1, 1, 'xavier_normal', torch.nn.LeakyReLU(0.2)).to(dev)
img, dep, fus = (torch.rand(2, c, 32, 80, generator=gen).to(dev) for c in (16, 16, 16))
k = torch.tensor([[60.0, 0, 40], [0, 60.0, 16], [0, 0, 1]]).view(1, 3, 3).repeat(2, 1, 1)
a = kb(img, dep, torch.inverse(k).to(dev), fus)
xy = torch.stack(torch.meshgrid(torch.arange(32.0), torch.arange(80.0), indexing='ij')[::-1] + (torch.ones(32, 80),), 0)
coords = torch.matmul(torch.inverse(k), xy.reshape(1, 3, -1).repeat(2, 1, 1)).view(2, 3, 32, 80).to(dev)
b = kb(img, dep, coords, fus)
for u, v in zip(a, b):
    assert torch.isfinite(u).all() and (u - v).abs().max() < 1e-3
# operand-format hand-over (separate kernel instantiations) and the block without conv_image
from kbnet_b200.modules import OperandFormat  # noqa: E402
of_i, of_f = OperandFormat(2, 32, 16, 40, dev), OperandFormat(2, 32, 16, 40, dev, even_only=True)
c = kb(OperandFormat.from_nchw(img), dep, torch.inverse(k).to(dev), None, out_image=of_i,
       fused_of=OperandFormat.from_nchw(fus, even_only=True), out_fused_of=of_f)
assert (of_i.to_nchw() - a[0]).abs().max() < 1e-3 and (c[2] - a[2]).abs().max() < 1e-3
d = kb(img, dep, torch.inverse(k).to(dev), fus, need_image=False)
assert d[0] is None and (d[1] - a[1]).abs().max() < 1e-3 and (d[2] - a[2]).abs().max() < 1e-3
wi = torch.randn(48, 3, 3, 3, generator=gen).to(dev) * 0.2
rgb = torch.rand(2, 3, 32, 80, generator=gen).to(dev)
of_c = kbnet_b200.ops.conv2d_fused([rgb], wi, 1, 0.2, out_of=OperandFormat(2, 48, 32, 80, dev))
assert (of_c.to_nchw() - kbnet_b200.ops.conv2d_fused([rgb], wi, 1, 0.2)).abs().max() < 1e-3
w = torch.randn(32, 48, 3, 3, generator=gen).to(dev) * 0.1
skip = torch.rand(2, 32, 32, 80, generator=gen).to(dev)
c1 = kbnet_b200.ops.conv2d_fused([img, skip], w, 1, 0.2)
wu = torch.randn(16, 32, 3, 3, generator=gen).to(dev) * 0.1
u1 = kbnet_b200.ops.upconv2x(c1, wu, 0.2)
wo = torch.randn(1, 16, 3, 3, generator=gen).to(dev) * 0.1
o1 = kbnet_b200.ops.conv3x3_depth_head(u1 if u1 is not None else c1[:, :16].contiguous(), wo, 1.5, 100.0)
torch.cuda.synchronize()
assert torch.isfinite(o1).all()
print('sanitize target ok: %d kernel launches' % (lib.kbnet_launch_count() - n0))
