"""output0 (12 -> 1, 3x3) fused with the depth head at 352x1216 (kbnet_conv3x3_depth_head_forward): CUDA-event time per launch.
usage: python tools/bench_head.py [N=8]   (KBNET_SMALL_STRIPS=1: the 4-pixel-strip form)"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kbnet_b200
torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0
w = torch.randn(1, 12, 3, 3, device=dev) * 0.1
xs = [torch.randn(N, 12, 352, 1216, device=dev) for _ in range(6)]
for x in xs:
    kbnet_b200.ops.conv3x3_depth_head(x, w, 1.5, 100.0)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
for i, (a, b) in enumerate(ev):
    a.record(); kbnet_b200.ops.conv3x3_depth_head(xs[i % 6], w, 1.5, 100.0); b.record()
torch.cuda.synchronize()
t = sorted(a.elapsed_time(b) for a, b in ev)
med = t[len(t) // 2]
byts = 4.0 * N * 13 * 352 * 1216
print('output0 + head N=%d: median %.1f us (min %.1f); algorithmic %.1f MB -> %.0f GB/s = %.1f%% of measured peak' % (N, med * 1e3, t[0] * 1e3, byts / 1e6, byts / med / 1e6, 100 * byts / med / 1e6 / peak))
