"""Top stalled SASS instructions of an `ncu --page source --csv` export, with their dominant stall reasons.
usage: python tools/ncu_top_stalls.py file.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
tot = 0
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    samp = int(r[ci['# Samples']] or 0)
    tot += samp
    data.append((samp, k, r))
print('total samples', tot)
for samp, k, r in sorted(data, key=lambda t: -t[0])[:n]:
    st = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print('%5d %5.1f%%  line %5d  exec %9s  %-70s %s' % (samp, 100.0 * samp / tot, k, r[ci['Instructions Executed']], r[ci['Source']].strip()[:70],
                                                    ' '.join('%s=%d' % (c, v) for v, c in st if v)))
