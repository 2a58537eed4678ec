"""Run individual layer shapes of the KITTI forward (batch 8) through the tensor-core conv engine, for ncu captures.
usage: python tools/profile_convs.py [name ...]   (no names = all)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kbnet_b200
torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
N = 8
LAYERS = {  # name: (sources [(C,H,W,mode)], F, k, stride)
    'conv0_image': ([(3, 352, 1216, 0)], 48, 3, 1),
    'kb1_conv_image': ([(48, 352, 1216, 0)], 48, 3, 2),
    'kb2_conv_image': ([(48, 176, 608, 0)], 96, 3, 2),
    'kb2_conv_fused': ([(48, 176, 608, 0), (3, 176, 608, 0), (48, 176, 608, 0)], 96, 1, 2),
    'deconv2_conv': ([(128, 88, 304, 0), (128, 88, 304, 0)], 128, 3, 1),
    'deconv1_conv': ([(64, 176, 608, 0), (64, 176, 608, 0)], 64, 3, 1),
    'deconv0_deconv': ([(64, 176, 608, 1)], 12, 3, 1),
    'deconv4_conv': ([(256, 22, 76, 0), (512, 22, 76, 0)], 256, 3, 1),
    'conv0_depth': ([(8, 352, 1216, 0)], 16, 3, 1),
    'output0': ([(12, 352, 1216, 0)], 1, 3, 1),
    'deconv0_conv': ([(12, 352, 1216, 0)], 12, 3, 1),
}
names = sys.argv[1:] or list(LAYERS)
for name in names:
    srcs, f, k, stride = LAYERS[name]
    ts, size = [], None
    for c, h, w, mode in srcs:
        t = torch.rand(N, c, h, w, device=dev)
        if mode:
            ts.append((t, 'up', (2 * h, 2 * w)))
        else:
            ts.append(t)
    wt = torch.randn(f, sum(s[0] for s in srcs), k, k, device=dev) * 0.05
    for _ in range(3):
        y = kbnet_b200.ops.conv2d_fused(ts, wt, stride, 0.2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = kbnet_b200.ops.conv2d_fused(ts, wt, stride, 0.2)
    e1.record(); torch.cuda.synchronize()
    print('%-16s out %-22s %8.1f us/launch' % (name, tuple(y.shape), e0.elapsed_time(e1) / 5 * 1e3))
