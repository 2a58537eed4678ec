"""Micro-benchmark of the full-resolution few-channel convolutions (KITTI 352x1216, batch 8): mma.sync kernel vs the tcgen05 engine.
usage: python tools/bench_small_conv.py [C F]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kbnet_b200
torch.set_grad_enabled(False)
kbnet_b200.set_autotune(True)
dev = torch.device('cuda:0')
try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0
N, H, W = 8, 352, 1216
shapes = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(3, 48), (8, 16), (12, 12)]
for c, f in shapes:
    xs = [torch.rand(N, c, H, W, device=dev) for _ in range(4)]
    w = torch.randn(f, c, 3, 3, device=dev) * 0.1
    outs = [torch.empty(N, f, H, W, device=dev) for _ in range(4)]
    for eng in (None, 'tcgen05'):
        for i in range(4):
            kbnet_b200.ops.conv2d_fused([xs[i]], w, 1, 0.2, out=outs[i], engine=eng)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        for i in range(20):
            ev[i][0].record()
            kbnet_b200.ops.conv2d_fused([xs[i % 4]], w, 1, 0.2, out=outs[i % 4], engine=eng)
            ev[i][1].record()
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in ev)
        med = t[len(t) // 2]
        byts = 4.0 * N * H * W * (c + f)
        print('conv3x3 %2d -> %2d  %-8s median %.1f us (min %.1f): %.0f GB/s = %.1f%% of measured HBM peak'
              % (c, f, eng or 'mma.sync', med * 1e3, t[0] * 1e3, byts / med / 1e6, 100 * byts / med / 1e6 / peak))
