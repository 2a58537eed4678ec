"""KB layer micro-benchmark at the four KITTI encoder levels (BASELINE config 2 = level 2: N=8, image/fused 48ch + depth 16ch
@176x608 -> 96/32/96 @88x304): CUDA-event time per forward, algorithmic GB/s (SURVEY 8d fast path: every input read once --
`fused` at even pixels only -- and every output written once) and fraction of the measured HBM peak.
Modes: 'nchw' = fp32 NCHW in and out; 'of' = conv_image in and out in operand format (what KBNetEncoder does);
'of+skip' (level 4) = conv_image not computed (dead in the encoder).
usage: python tools/bench_kb.py [N=8] [level=all|1..4] [mode=all|nchw|of]"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kbnet_b200
from kbnet_b200.modules import OperandFormat
torch.set_grad_enabled(False)
if os.environ.get('KBNET_AUTOTUNE', '1') != '0':
    kbnet_b200.set_autotune(True)
dev = torch.device('cuda:0')
N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
lv = sys.argv[2] if len(sys.argv) > 2 else 'all'
md = sys.argv[3] if len(sys.argv) > 3 else 'all'
# level: (C_i, C_d, C_f, F_i, F_d, F_f, H, W)
CFG = {1: (48, 16, 0, 48, 16, 48, 352, 1216), 2: (48, 16, 48, 96, 32, 96, 176, 608),
       3: (96, 32, 96, 192, 64, 192, 88, 304), 4: (192, 64, 192, 384, 128, 384, 44, 152)}
try:
    peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    peak = 6650.0


def run(lvl, mode):
    ci, cd, cf, fi, fd, ff, H, W = CFG[lvl]
    blk = kbnet_b200.CalibratedBackprojectionBlock(ci, cd, ci + cf, fi, fd, ff, 1, 1, 1, 'xavier_normal', torch.nn.LeakyReLU(0.2)).to(dev)
    kinv = torch.inverse(torch.tensor([[360.77, 0, 304.78], [0, 360.77, 86.4], [0, 0, 1.0]])).view(1, 3, 3).repeat(N, 1, 1).to(dev)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    NB = 4 if lvl > 1 else 3  # rotate buffer sets: >> 126 MB L2
    sets = []
    for _ in range(NB):
        img = torch.rand(N, ci, H, W, device=dev)
        fus = torch.rand(N, cf, H, W, device=dev) if cf else None
        fus_of = None
        if mode != 'nchw':
            img = OperandFormat.from_nchw(img)
            if cf:
                fus, fus_of = None, OperandFormat.from_nchw(fus, even_only=True)
        sets.append((img, torch.rand(N, cd, H, W, device=dev), fus, OperandFormat(N, fi, Ho, Wo, dev) if mode == 'of' else None, fus_of,
                     OperandFormat(N, ff, Ho, Wo, dev, even_only=True) if (mode == 'of' and lvl < 4) else None))
    skip = mode == 'of+skip'
    call = lambda s: blk(s[0], s[1], kinv, s[2], out_image=s[3], fused_of=s[4], out_fused_of=s[5], need_image=not skip)
    for s in sets:
        call(s)
    torch.cuda.synchronize()
    iters = 20
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for i in range(iters):
        s = sets[i % NB]
        ev[i][0].record()
        call(s)
        ev[i][1].record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    med = t[len(t) // 2]
    img_in = ci * H * W if not skip else ci * Ho * Wo     # even pixels only when conv_image is not computed
    byts = 4.0 * N * (img_in + cd * H * W + cf * Ho * Wo + ((0 if skip else fi) + fd + ff) * Ho * Wo)
    flops = 2.0 * N * Ho * Wo * ((0 if skip else 9 * ci * fi) + 9 * (cd + 3) * fd + (ci + 3 + cf) * ff) + 2.0 * N * H * W * cd
    print('KB level %d N=%d %-7s: median %.1f us/forward (min %.1f); algorithmic %.1f MB -> %.0f GB/s = %.1f%% of measured peak %.0f GB/s; '
          '%.2f GFLOP -> %.1f TFLOP/s fp32-equivalent' % (lvl, N, mode, med * 1e3, t[0] * 1e3, byts / 1e6, byts / med / 1e6,
                                                         100 * byts / med / 1e6 / peak, peak, flops / 1e9, flops / med / 1e9), flush=True)
    del sets
    torch.cuda.empty_cache()


for lvl in ([1, 2, 3, 4] if lv == 'all' else [int(lv)]):
    for mode in (['nchw', 'of'] if md == 'all' else [md]):
        run(lvl, mode)
    if lvl == 4 and md in ('all', 'of'):
        run(lvl, 'of+skip')
