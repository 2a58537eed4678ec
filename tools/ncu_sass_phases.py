"""Aggregate an `ncu --page source --csv` export (SASS view) by phase (split at BAR.SYNC) and opcode.
usage: python tools/ncu_sass_phases.py gpurun_out/prof_source.csv [npx]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
npx = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
phase = 0
per_phase = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
tot_inst = tot_samp = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    sass = r[ci['Source']].strip()
    op = sass.split()[0] if not sass.startswith('@') else sass.split()[1]
    op = op.rstrip(';')
    inst = int(r[ci['Instructions Executed']] or 0)
    samp = int(r[ci['# Samples']] or 0)
    p = per_phase[phase]
    p[0] += inst; p[1] += samp
    base = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'BAR', 'HMMA', 'LDSM')) and '.' in op else '')
    p[2][base] += inst; p[3][base] += samp
    tot_inst += inst; tot_samp += samp
    if op.startswith('BAR'):
        phase += 1
print('total warp-instructions %d, samples %d%s' % (tot_inst, tot_samp, (', %.2f warp-instr/px' % (tot_inst / npx)) if npx else ''))
for ph in sorted(per_phase):
    inst, samp, ops, sops = per_phase[ph]
    print('\nphase %d: instr %d (%.1f%%)%s  stall-samples %d (%.1f%%)' % (ph, inst, 100.0 * inst / tot_inst, (' = %.2f/px' % (inst / npx)) if npx else '', samp, 100.0 * samp / max(tot_samp, 1)))
    print('   by instr  : ' + ', '.join('%s %.1f%%' % (o, 100.0 * c / max(inst, 1)) for o, c in ops.most_common(14)))
    print('   by samples: ' + ', '.join('%s %.1f%%' % (o, 100.0 * c / max(samp, 1)) for o, c in sops.most_common(10)))
