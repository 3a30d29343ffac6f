#!/usr/bin/env python
"""Static SASS instruction counts per kernel of libgagan_b200.so (no GPU needed):  python tools/sass_summary.py > profiles/r2_sass_summary.txt
The proof that the hot kernels are tcgen05 / TMA code: UTCHMMA (tcgen05.mma kind::tf32), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA box
loads), UBLKCP (bulk copies), SYNCS (mbarrier), UTCBAR (tcgen05.commit), UTCATOMSWS (tcgen05.alloc / dealloc)."""
import os, re, subprocess, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'ga-gan_b200', 'libgagan_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', so], stdout=subprocess.PIPE, text=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), stdout=subprocess.PIPE, text=True).stdout.splitlines()
ops = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'SYNCS', 'UTCBAR', 'UTCATOMSWS', 'FFMA', 'LDS', 'STS', 'LDG', 'STG']
rows = []
for name, block in zip(names, re.split(r'Function : \S+', sass)[1:]):
    insts = re.findall(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', block)
    cnt = collections.Counter(i.split('.')[0] for i in insts)
    short = re.sub(r'\(anonymous namespace\)::|void |\(.*', '', name)
    rows.append((short, len(insts), [cnt.get(o, 0) for o in ops]))
print('SASS summary of ga-gan_b200/libgagan_b200.so (cuobjdump -sass, sm_100a): static instruction counts per kernel function.')
print('UTCHMMA = tcgen05.mma kind::tf32, LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory), UTMALDG = cp.async.bulk.tensor (TMA box load),')
print('UBLKCP = cp.async.bulk, SYNCS = mbarrier operations, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc.')
print('Regenerate: python tools/sass_summary.py\n')
print(f"{'kernel':46s} {'total':>6s} " + ' '.join(f'{o:>8s}' for o in ops))
tc = [r for r in rows if r[2][0] > 0]
rest = [r for r in rows if r[2][0] == 0]
for short, total, c in sorted(tc, key=lambda r: -r[1]) + sorted(rest, key=lambda r: -r[1]):
    print(f'{short[:46]:46s} {total:6d} ' + ' '.join(f'{v:8d}' for v in c))
