"""Per-call CUDA-event timing of one full KBNet forward (KITTI 352x1216, batch 8): which layer costs what.
usage: python tools/profile_layers.py [batch]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kbnet_b200
import kbnet_b200.model as M
from kbnet_b200 import ops, modules
from bench import synth_inputs, WORKLOADS, F_IMG, F_DEP, F_DEC
CFG = WORKLOADS['kitti']
MIN_POOLS, MAX_POOLS = CFG['min_pools'], CFG['max_pools']

torch.set_grad_enabled(False)
if os.environ.get('KBNET_AUTOTUNE', '1') != '0':
    kbnet_b200.set_autotune(True)
dev = torch.device('cuda:0')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = kbnet_b200.KBNetModel(3, 2, MIN_POOLS, MAX_POOLS, 3, 8, F_IMG, F_DEP, [0, 1, 2, 3], F_DEC, device=dev)
inp = [t.to(dev) for t in synth_inputs(CFG, N, 1)]
for _ in range(3):
    model.forward(*inp)
torch.cuda.synchronize()

records = []


def timed(name_fn, fn):
    def wrapper(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = fn(*a, **kw)
        e1.record()
        records.append((name_fn(a, kw, y), e0, e1))
        return y
    return wrapper


def conv_name(a, kw, y):
    srcs, w = a[0], a[1]
    a = list(a) + [kw.get('stride', 1)] * 3
    desc = '+'.join(('up' if isinstance(s, (tuple, list)) else '') + str((s[0] if isinstance(s, (tuple, list)) else s).shape[1]) for s in srcs)
    flops = 2.0 * y.numel() * w.shape[1] * w.shape[2] * w.shape[3]
    return 'conv %dx%d s%d  %s -> %d @%dx%d' % (w.shape[2], w.shape[3], a[2], desc, w.shape[0], y.shape[2], y.shape[3]), flops


def kb_name(a, kw, y):
    img = kw['image']
    return 'KB block  img %d @%dx%d' % (img.shape[1], img.shape[2], img.shape[3]), 0.0


ops.conv2d_fused = timed(conv_name, ops.conv2d_fused)


def up_name(a, kw, y):
    x, w = a[0], a[1]
    if y is None:
        return 'upconv2x (declined)', 0.0
    return 'upconv2x (4 parity 2x2 convs)  %d -> %d @%dx%d' % (x.shape[1], w.shape[0], y.shape[2], y.shape[3]), 2.0 * y.numel() * w.shape[1] * 9


ops.upconv2x = timed(up_name, ops.upconv2x)
modules.CalibratedBackprojectionBlock.forward = timed(kb_name, modules.CalibratedBackprojectionBlock.forward)
modules.SparseToDensePool.forward = timed(lambda a, kw, y: ('S2D', 0.0), modules.SparseToDensePool.forward)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
model.forward(*inp)
e1.record()
torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
print('batch %d  total %.3f ms (%.1f frames/s incl. event overhead)' % (N, tot, N / tot * 1e3))
for (name, flops), a, b in records:
    ms = a.elapsed_time(b)
    print('%-52s %8.1f us  %5.1f%%  %s' % (name, ms * 1e3, 100 * ms / tot, ('%.0f TFLOP/s(fp32-equivalent)' % (flops / ms / 1e9)) if flops else ''))
