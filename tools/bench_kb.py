"""KB layer micro-benchmark on BASELINE config 2 (block 2, N=8, image/fused 48ch + depth 16ch @176x608 -> 96/32/96 @88x304):
CUDA-event time per forward, algorithmic GB/s (452.0 MB, SURVEY 8d fast path) and fraction of the measured HBM peak.
usage: python tools/bench_kb.py [N] [level=2]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kbnet_b200
torch.set_grad_enabled(False)
if os.environ.get('KBNET_AUTOTUNE', '1') != '0':
    kbnet_b200.set_autotune(True)
dev = torch.device('cuda:0')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lvl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
# level: (C_i, C_d, C_f, F_i, F_d, F_f, H, W)
CFG = {1: (48, 16, 0, 48, 16, 48, 352, 1216), 2: (48, 16, 48, 96, 32, 96, 176, 608),
       3: (96, 32, 96, 192, 64, 192, 88, 304), 4: (192, 64, 192, 384, 128, 384, 44, 152)}
ci, cd, cf, fi, fd, ff, H, W = CFG[lvl]
try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0
blk = kbnet_b200.CalibratedBackprojectionBlock(ci, cd, ci + cf, fi, fd, ff, 1, 1, 1, 'xavier_normal', torch.nn.LeakyReLU(0.2)).to(dev)
kinv = torch.inverse(torch.tensor([[360.77, 0, 304.78], [0, 360.77, 86.4], [0, 0, 1.0]])).view(1, 3, 3).repeat(N, 1, 1).to(dev)
NB = 4  # rotate buffer sets: 4 x ~450 MB >> 126 MB L2
sets = [(torch.rand(N, ci, H, W, device=dev), torch.rand(N, cd, H, W, device=dev), torch.rand(N, cf, H, W, device=dev) if cf else None)
        for _ in range(NB)]
for s in sets:
    blk(s[0], s[1], kinv, s[2])
torch.cuda.synchronize()
iters = 20
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
for i in range(iters):
    s = sets[i % NB]
    ev[i][0].record()
    blk(s[0], s[1], kinv, s[2])
    ev[i][1].record()
torch.cuda.synchronize()
t = sorted(a.elapsed_time(b) for a, b in ev)
med = t[len(t) // 2]
Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
byts = 4.0 * N * ((ci + cd) * H * W + cf * Ho * Wo + (fi + fd + ff) * Ho * Wo)
flops = 2.0 * N * Ho * Wo * (9 * ci * fi + 9 * (cd + 3) * fd + (ci + 3 + cf) * ff) + 2.0 * N * H * W * cd
print('KB level %d N=%d: median %.1f us/forward (min %.1f); algorithmic %.1f MB -> %.0f GB/s = %.1f%% of measured peak %.0f GB/s; '
      '%.2f GFLOP -> %.1f TFLOP/s fp32-equivalent' % (lvl, N, med * 1e3, t[0] * 1e3, byts / 1e6, byts / med / 1e6, 100 * byts / med / 1e6 / peak, peak,
                                                     flops / 1e9, flops / med / 1e9))
