"""Up-convolution micro-benchmark (nearest-2x + 3x3 as four parity 2x2 convolutions, kbnet_upconv2x_forward): CUDA-event time per
launch and algorithmic GB/s.  usage: python tools/bench_upconv.py [C=64] [F=12] [Hs=176] [Ws=608] [N=8]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kbnet_b200
torch.set_grad_enabled(False)
if os.environ.get('KBNET_AUTOTUNE', '1') != '0':
    kbnet_b200.set_autotune(True)
a = [int(v) for v in sys.argv[1:]]
C, F, Hs, Ws, N = (a + [64, 12, 176, 608, 8][len(a):])[:5]
dev = torch.device('cuda:0')
try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0
w = torch.nn.Parameter(torch.randn(F, C, 3, 3, device=dev) * 0.1, requires_grad=False)
NB = 4
xs = [torch.rand(N, C, Hs, Ws, device=dev) for _ in range(NB)]
outs = [torch.empty(N, F, 2 * Hs, 2 * Ws, device=dev) for _ in range(NB)]
for x, o in zip(xs, outs):
    assert kbnet_b200.ops.upconv2x(x, w, 0.2, out=o) is not None
torch.cuda.synchronize()
iters = 20
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
for i in range(iters):
    ev[i][0].record()
    kbnet_b200.ops.upconv2x(xs[i % NB], w, 0.2, out=outs[i % NB])
    ev[i][1].record()
torch.cuda.synchronize()
t = sorted(p.elapsed_time(q) for p, q in ev)
med = t[len(t) // 2]
byts = 4.0 * N * (C * Hs * Ws + F * 4 * Hs * Ws)
flops = 2.0 * N * 4 * Hs * Ws * 9 * C * F
print('upconv2x %d -> %d @%dx%d N=%d: median %.1f us (min %.1f); algorithmic %.1f MB -> %.0f GB/s = %.1f%% of measured peak; %.1f TFLOP/s fp32-equivalent'
      % (C, F, Hs, Ws, N, med * 1e3, t[0] * 1e3, byts / 1e6, byts / med / 1e6, 100 * byts / med / 1e6 / peak, flops / med / 1e9))
