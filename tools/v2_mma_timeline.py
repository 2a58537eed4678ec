"""Where the MMA issuer of CTA 0 spends a tile: from a KBNET_V2_TIMELINE dump (library built with KBNET_V2_STAMPS=1), per stage
the wait for the operand stage (A_FULL) and the issue time up to its commit; per tile the wait for the accumulator buffer.
usage: python tools/v2_mma_timeline.py file [tile=3]"""
import sys, collections
ev = collections.defaultdict(dict)
hdr = ''
for line in open(sys.argv[1]):
    if line.startswith('#'):
        hdr = line.strip(); continue
    r, i, t = line.split(); ev[int(r)][int(i)] = int(t)
print(hdr)
ns = int(hdr.split()[2])
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = ev[2]
n_tiles = len(m) // (2 * ns)
print('tiles stamped:', n_tiles)
for tl in range(max(0, tile - 1), min(n_tiles, tile + 2)):
    base = tl * ns * 2
    prev = ev[4].get(tl)
    t_start = prev
    line = []
    tot_w = tot_i = 0
    last = prev
    for s in range(ns):
        a, c = m[base + 2 * s], m[base + 2 * s + 1]
        line.append('%d:w%d/i%d' % (s, a - last, c - a))
        tot_w += a - last; tot_i += c - a
        last = c
    tprev_end = m[base - 1] if base else prev
    print('tile %d: acc wait %d | %s | wait %d issue %d total %d' % (tl, prev - tprev_end, ' '.join(line), tot_w, tot_i, last - prev))
    if 2 * tl + 1 in ev[3]:
        print('   epilogue: start +%d (after last commit of its tile), duration %d' % (ev[3][2 * tl] - last, ev[3][2 * tl + 1] - ev[3][2 * tl]))
if n_tiles > 2:
    t_first = ev[4][1]; t_last = ev[4][n_tiles - 1]
    print('steady-state clocks per tile: %.0f' % ((t_last - t_first) / (n_tiles - 2)))
