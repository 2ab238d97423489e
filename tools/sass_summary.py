#!/usr/bin/env python3
"""Summarise an `ncu --page source --print-source sass --csv` dump: opcode mix, stall reasons, hottest instructions."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ia, ie, it, iss = h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
ops, samples, tot, tthr, data = collections.Counter(), collections.Counter(), 0, 0, []
for r in rows[2:]:
    try:
        n = int(r[ie])
    except Exception:
        continue
    tok = r[ia].split()
    op = (tok[1] if tok[0].startswith('@') else tok[0]).split('.')[0]
    ops[op] += n; tot += n; tthr += int(r[it]); samples[op] += int(r[iss] or 0)
    data.append((int(r[iss] or 0), n, int(r[it]), r[ia]))
print('kernel:', rows[0][1][:90])
print('warp instructions', tot, 'SASS lines', len(data), 'avg active threads %.1f' % (tthr / max(tot, 1)))
for k, v in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
    print('  %-10s %14d %5.1f%%  samples %d' % (k, v, 100 * v / tot, samples[k]))
st = [c for c in h if c.startswith('stall_') and 'Not' not in c]
agg = collections.Counter()
for r in rows[2:]:
    for c in st:
        try:
            agg[c] += int(r[h.index(c)])
        except Exception:
            pass
s = sum(agg.values())
print('stalls:', ', '.join('%s %.1f%%' % (k[6:], 100 * v / s) for k, v in agg.most_common(8)))
