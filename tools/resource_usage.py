"""Registers / stack / shared memory / spills of every kernel in libfokl_b200.so (cuobjdump --dump-resource-usage): the
`-Xptxas -v` facts, read from the built binary.   usage: python tools/resource_usage.py [lib.so] > profiles/rNN_resource_usage.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.split('\n')
    return [short(s) for s in out]


def short(sig):
    """'void ns::kern<1, true>(double*, int)' -> 'ns::kern<1, true>'; anonymous-namespace prefixes dropped."""
    sig = sig.replace('(anonymous namespace)::', '')
    sig = re.sub(r'^void ', '', sig)
    depth = 0
    for i, ch in enumerate(sig):
        if ch == '<':
            depth += 1
        elif ch == '>':
            depth -= 1
        elif ch == '(' and depth == 0:
            return sig[:i]
    return sig


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'fokl-gpy_b200', 'csrc', 'libfokl_b200.so')
    txt = subprocess.run(['cuobjdump', '--dump-resource-usage', lib], capture_output=True, text=True, check=True).stdout
    rows, unit = [], ''
    lines = txt.split('\n')
    for i, ln in enumerate(lines):
        m = re.match(r'identifier = (\S+)$', ln.strip())
        if m:
            unit = os.path.basename(m.group(1))
        m = re.match(r'\s*Function (\S+):', ln)
        if m and i + 1 < len(lines):
            f = dict(re.findall(r'([\w\[\]]+):(\d+)', lines[i + 1]))
            rows.append((unit, m.group(1), f))
    names = demangle([r[1] for r in rows])
    print('# cuobjdump --dump-resource-usage of %s (sm_100a): per kernel registers per thread, stack bytes per thread\n'
          '# (non-zero = local arrays or spills), static shared memory, constant bank 0 (parameters); dynamic shared\n'
          '# memory is set at launch (csrc/*.cu).' % os.path.relpath(lib, ROOT))
    print('%-14s %-72s %5s %6s %8s %6s' % ('unit', 'kernel', 'REG', 'STACK', 'SHARED', 'CONST0'))
    for (unit, _, f), name in sorted(zip(rows, names), key=lambda t: (t[0][0], t[1])):
        print('%-14s %-72s %5s %6s %8s %6s' % (unit, name[:72], f.get('REG', '?'), f.get('STACK', '?'), f.get('SHARED', '?'),
                                             f.get('CONSTANT[0]', '?')))


if __name__ == '__main__':
    main()
