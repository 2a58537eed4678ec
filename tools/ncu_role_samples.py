"""Stall samples of an `ncu --page source --csv` export summed over SASS line ranges (roles of a warp-specialised kernel).
usage: python tools/ncu_role_samples.py file.csv name:lo-hi [name:lo-hi ...]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) >= len(hdr)]
tot = sum(int(r[ci['# Samples']] or 0) for r in body)
for spec in sys.argv[2:]:
    name, rng = spec.split(':'); lo, hi = map(int, rng.split('-'))
    samp = inst = 0; st = collections.Counter(); wait = 0
    for k, r in enumerate(body):
        if lo <= k <= hi:
            s = int(r[ci['# Samples']] or 0); samp += s; inst += int(r[ci['Instructions Executed']] or 0)
            for c in stall_cols: st[c[6:]] += int(r[ci[c]] or 0)
            src = r[ci['Source']]
            if 'TRYWAIT' in src or ('BRA' in src and int(r[ci['stall_long_sb']] or 0) > 0.8 * max(s, 1) and s > 20): wait += s
    print('%-10s lines %5d-%5d: samples %6d (%5.1f%%)  in mbarrier waits %6d (%4.1f%% of role)  warp-instr %9d  top stalls: %s'
          % (name, lo, hi, samp, 100.0 * samp / tot, wait, 100.0 * wait / max(samp, 1), inst, ' '.join('%s=%d' % kv for kv in st.most_common(4))))
