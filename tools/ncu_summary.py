#!/usr/bin/env python3
"""Turns `ncu --page raw --csv` dumps into the text summary kept under profiles/ and refreshes profiles/traffic.json.
    python tools/ncu_summary.py <out.txt> <header line> raw1.csv [raw2.csv ...]"""
import csv, json, os, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
out, header, files = sys.argv[1], sys.argv[2], sys.argv[3:]
lines, traffic = [header], {}
for f in files:
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    lines.append('== %s   units: %s' % (f, {k: units[hdr.index(k)] for k in KEYS if k in hdr}))
    seen = set()
    for r in rows[2:]:
        name = r[ki].replace('void ', '').split('(')[0]
        if name in seen:
            continue
        seen.add(name)
        lines.append('---')
        lines.append('  Kernel Name = ' + r[ki][:110])
        for k in KEYS:
            if k in hdr:
                lines.append('  %s = %s' % (k, r[hdr.index(k)]))
        b = 0.0
        for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            b += float(r[hdr.index(k)]) * SCALE.get(units[hdr.index(k)], 1.0)
        traffic[name] = b
open(out, 'w').write('\n'.join(lines) + '\n')
tj = os.path.join(os.path.dirname(out), 'traffic.json')
pick = lambda s: next((v for k, v in traffic.items() if k.startswith(s)), None)
# merge into the existing record: a capture of one kernel must not drop the entries of the others
try:
    old = json.load(open(tj))
except Exception:
    old = {}
rec = {'interior': pick('interior_kernel<Dual'), 'primary_edges': pick('primary_edge_kernel'), 'secondary_edges': pick('secondary_edge_kernel')}
for k in rec:
    if rec[k] is None:
        rec[k] = old.get(k)
allk = dict(old.get('all', {}))
allk.update({'%s (%s)' % (k, os.path.basename(out).split('_')[0]): v for k, v in traffic.items()})
rec['source'] = '%s for the kernels it holds, earlier captures otherwise (see the tags in "all"); dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full' % out
rec['all'] = allk
json.dump(rec, open(tj, 'w'), indent=1)
print(open(out).read()[:600])
