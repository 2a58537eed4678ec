"""SASS evidence for the Blackwell-native claims: per kernel of libkbnet_b200.so, the count of tcgen05 MMAs (UTC*MMA), TMEM
loads (LDTM), TMA tensor / bulk copies (UTMALDG / UBLKCP), legacy tensor-path MMAs (HMMA), ldmatrix (LDSM), packed fp32 FMAs
(FFMA2), 3-input min/max (FMNMX3), plus code size and registers.  usage: python tools/sass_evidence.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'calibrated-backprojection-network_b200', 'libkbnet_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True).stdout
res = subprocess.run(['cuobjdump', '-res-usage', LIB], stdout=subprocess.PIPE, text=True).stdout
regs = {}
cur = None
for l in res.splitlines():
    m = re.search(r'Function (\S+):', l)
    if m:
        cur = m.group(1)
    m = re.search(r'REG:(\d+).*?SHARED:(\d+)', l)
    if m and cur:
        regs[cur] = (int(m.group(1)), int(m.group(2)))
WANT = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'LDSM', 'FFMA2', 'FMNMX3', 'SYNCS', 'BAR']
funcs = collections.OrderedDict()
cur = None
for l in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', l)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
    if m and cur:
        funcs[cur][m.group(1)] += 1
        funcs[cur]['_n'] += 1
print('# %s  (cuobjdump -sass; sm_100a)' % os.path.relpath(LIB, ROOT))
print('%-86s %6s %5s %s' % ('kernel (demangled prefix)', 'instr', 'regs', '  '.join(WANT)))
tot = collections.Counter()
for f, c in funcs.items():
    name = subprocess.run(['c++filt', f], stdout=subprocess.PIPE, text=True).stdout.strip()
    name = re.sub(r'kbnet::\(anonymous namespace\)::', '', name)
    name = re.sub(r'\(.*', '', name)[:86]
    print('%-86s %6d %5s %s' % (name, c['_n'], regs.get(f, ('?',))[0], '  '.join('%*d' % (len(w), c[w]) for w in WANT)))
    tot.update(c)
print('%-86s %6d %5s %s' % ('TOTAL', tot['_n'], '', '  '.join('%*d' % (len(w), tot[w]) for w in WANT)))
