#!/usr/bin/env python3
"""Attributes an `ncu --page source --print-source sass --csv` dump to source lines: nvdisasm -g of the same kernel in
the object file gives file:line per instruction offset.   usage: sass_by_line.py sass.csv object.o <kernel substring> [n_warps] [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
prof, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
W = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout.splitlines()
line_of, cur, inside = {}, ('?', 0), False
for l in dis:
    if l.startswith('//-') and '.text.' in l:
        inside = kname in l
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(prof)))
h = rows[1]
ie, it, iss = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[2:]:
    if not r or not r[0].startswith('0x') or not r[ie].isdigit():
        if r and r[0] == 'Kernel Name' and base is not None:
            break
        continue
    a = int(r[0], 16)
    base = a if base is None else base
    k = line_of.get(a - base, ('?', 0))
    agg[k][0] += int(r[ie]); agg[k][1] += int(r[it]); agg[k][2] += int(r[iss] or 0)
tot = sum(v[0] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print('total warp instr/warp %.0f' % (tot / W))
src = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    if f not in src:
        try:
            src[f] = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'psdr_jit_b200', 'csrc', f)).read().splitlines()
        except Exception:
            src[f] = []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ''
    print('%5.1f%% samples %5.1f%% instr  thr %4.1f  %s:%d  %s' % (100 * v[2] / ts, 100 * v[0] / tot, v[1] / max(v[0], 1), f, ln, text))
