"""SASS opcode census of libfokl_b200.so per kernel (cuobjdump -sass): the instructions that prove the design --
DMMA (FP64 tensor pipe), UTMALDG / UBLKCP (TMA tensor / bulk copies), SYNCS (mbarrier), UCGABAR (cluster barrier),
plus the memory / barrier opcodes.   usage: python tools/sass_census.py [lib.so] > profiles/rNN_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'fokl-gpy_b200', 'csrc', 'libfokl_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
WATCH = ['DMMA', 'UTMALDG', 'UBLKCP', 'UTMAPF', 'SYNCS', 'UCGABAR', 'BAR', 'LDG', 'STG', 'LDS', 'STS', 'LDSM', 'ATOM', 'RED',
         'SHFL', 'DFMA', 'DMUL', 'DADD', 'MUFU', 'MEMBAR', 'ERRBAR', 'UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'HMMA']
kern, counts, total = None, {}, collections.Counter()
for line in txt.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r'\(anonymous namespace\)::', '', kern)
        kern = kern.split('(')[0][-70:]
        counts[kern] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and kern:
        op = m.group(1)
        counts[kern]['_all'] += 1
        base = op.split('.')[0]
        if base.startswith('UCGABAR'):
            base = 'UCGABAR'
        if base in WATCH:
            counts[kern][base] += 1
            total[base] += 1
print('SASS census of', os.path.relpath(so, ROOT), '(sm_100a cubins)')
print('total over all kernels:', ', '.join('%s x%d' % (k, v) for k, v in total.most_common()))
print('no tcgen05 opcodes (UTC*MMA / LDTM / STTM) are expected: FP64 has no tcgen05 kind (SURVEY 8d)')
print()
cols = ['DMMA', 'UTMALDG', 'UBLKCP', 'SYNCS', 'UCGABAR', 'BAR', 'LDG', 'STG', 'LDS', 'STS', 'SHFL', 'DFMA', 'ATOM', 'MEMBAR']
print('%-72s %7s ' % ('kernel', 'instrs') + ' '.join('%7s' % c for c in cols))
for k in sorted(counts, key=lambda k: -counts[k]['_all']):
    c = counts[k]
    print('%-72s %7d ' % (k, c['_all']) + ' '.join('%7d' % c[x] for x in cols))
