"""Aggregate the per-instruction warp-stall samples of an `ncu --page source --csv` export: totals per stall reason
and per SASS opcode, plus the top instructions."""
import csv
import re
import sys
import collections


def main(path, top=18):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    src_i, samp_i = hdr.index('Source'), hdr.index('# Samples')
    by_reason = collections.Counter()
    by_op = collections.Counter()
    inst = []
    total = 0
    for r in rows[hdr_i + 1:]:
        if len(r) <= samp_i:
            continue
        try:
            n = int(r[samp_i])
        except ValueError:
            continue
        total += n
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[src_i])
        op = m.group(2) if m else '?'
        by_op[op.split('.')[0]] += n
        best = ('', 0)
        for i in stall_cols:
            try:
                v = int(r[i])
            except ValueError:
                continue
            by_reason[hdr[i]] += v
            if v > best[1]:
                best = (hdr[i], v)
        inst.append((n, r[src_i].strip()[:70], best[0]))
    print('samples', total)
    print('by reason:', ', '.join('%s %.1f%%' % (k.replace('stall_', ''), 100.0 * v / max(total, 1)) for k, v in by_reason.most_common(9)))
    print('by opcode:', ', '.join('%s %.1f%%' % (k, 100.0 * v / max(total, 1)) for k, v in by_op.most_common(10)))
    for n, s, b in sorted(inst, reverse=True)[:top]:
        print('%6d %5.1f%%  %-70s %s' % (n, 100.0 * n / max(total, 1), s, b))


if __name__ == '__main__':
    main(sys.argv[1])
