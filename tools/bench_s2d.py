"""S2D kernel micro-benchmark: CUDA-event time per launch, algorithmic GB/s and fraction of the measured HBM peak.
usage: python tools/bench_s2d.py [N] [cfg=KITTI|VOID] [C_in]"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kbnet_b200
torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = sys.argv[2] if len(sys.argv) > 2 else 'KITTI'
C = int(sys.argv[3]) if len(sys.argv) > 3 else 2
H, W, mins, maxs, p, dmax = (352, 1216, [5, 7, 9, 11, 13], [15, 17], 0.05, 80.0) if cfg == 'KITTI' else (480, 640, [15, 17], [23, 27, 29], 0.005, 7.0)
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0
gen = torch.Generator().manual_seed(1)
m = kbnet_b200.SparseToDensePool(C, mins, maxs, 8, 3, 'xavier_normal', 'leaky_relu').to(dev)
NB = 6   # rotate buffer sets so that consecutive launches do not find their input/output in L2 (6 x 137 MB >> 126 MB)
xs = []
for _ in range(NB):
    mask = (torch.rand(N, 1, H, W, generator=gen) < p).float()
    d = (torch.rand(N, 1, H, W, generator=gen) * (dmax - 1) + 1) * mask
    xs.append(torch.cat([d, (d > 0).float()], 1)[:, :C].contiguous().to(dev))
for x in xs:
    m(x)
torch.cuda.synchronize()
iters = 30
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
outs = []
for i in range(iters):
    ev[i][0].record()
    y = m(xs[i % NB])
    ev[i][1].record()
    outs.append(y) if i < NB else None
torch.cuda.synchronize()
t = sorted(a.elapsed_time(b) for a, b in ev)
med = t[len(t) // 2]
byts = 4.0 * H * W * (C + 8) * N
print('S2D %s N=%d C_in=%d: median %.1f us/launch (min %.1f) = %.2f us/frame; algorithmic %.1f MB -> %.0f GB/s = %.1f%% of measured peak %.0f GB/s'
      % (cfg, N, C, med * 1e3, t[0] * 1e3, med * 1e3 / N, byts / 1e6, byts / med / 1e6, 100 * byts / med / 1e6 / peak, peak))
