"""Print the per-stage timeline of CTA 0 from a KBNET_V2_TIMELINE dump (clock64 stamps).
usage: python tools/v2_timeline.py file [first_stage] [n_stages] [tma_chunks_per_stage=2]"""
import sys, collections
ev = collections.defaultdict(dict)
hdr = ''
for line in open(sys.argv[1]):
    if line.startswith('#'):
        hdr = line.strip(); continue
    r, i, t = line.split(); ev[int(r)][int(i)] = int(t)
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 24
cps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
print(hdr)
t0 = min(min(v.values()) for v in ev.values())
per = 2 + cps
print('stage | conv: a_empty  raw_full..  done | mma: a_full issued | cvt time  mma wait->issue   (clk since start)')
for s in range(first, first + n):
    c = [ev[0].get(s * per + k) for k in range(per)]
    m = [ev[2].get(s * 2 + k) for k in range(2)]
    if c[0] is None or m[0] is None: break
    f = lambda x: '%8d' % (x - t0) if x else '       -'
    print('%5d | %s | %s | %6d %6d' % (s, ' '.join(f(x) for x in c), ' '.join(f(x) for x in m), c[-1] - c[0], m[1] - m[0]))
print('tiles: mma acc_empty ok at', [v - t0 for k, v in sorted(ev[4].items())][:8])
print('epilogue (acc_full, done):', [(ev[3][2 * k] - t0, ev[3][2 * k + 1] - t0) for k in range(min(8, len(ev[3]) // 2))])
print('raw producer slot-free times:', [v - t0 for k, v in sorted(ev[1].items())][:24])
