#!/usr/bin/env python3
"""Per-address-range histogram of an `ncu --page source --print-source sass --csv` dump: warp instructions per warp,
average active threads and stall samples for every SEG consecutive SASS lines.  usage: sass_segments.py file.csv n_warps [seg]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
W = float(sys.argv[2]); seg = int(sys.argv[3]) if len(sys.argv) > 3 else 400
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; blocks.append(cur)
    elif r and r[0] == 'Address':
        cur['h'] = r
    elif cur is not None and r:
        cur['rows'].append(r)
for b in blocks:
    h = b['h']; ia, ie, it, iss = h.index('Source'), h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
    data = [(r[ia], int(r[ie]), int(r[it]), int(r[iss] or 0)) for r in b['rows'] if r[ie].isdigit()]
    tot = sum(x[1] for x in data)
    print(b['name'][:80], 'SASS lines', len(data), 'instr/warp %.0f' % (tot / W), 'thr %.1f' % (sum(x[2] for x in data) / max(tot, 1)))
    for s in range(0, len(data), seg):
        d = data[s:s + seg]
        n = sum(x[1] for x in d); t = sum(x[2] for x in d); sm = sum(x[3] for x in d)
        print('%6d  instr/warp %8.0f  thr %4.1f  samples %7d  maxexec/warp %.1f' % (s, n / W, t / max(n, 1), sm, max(x[1] for x in d) / W))
