#!/usr/bin/env python
"""Hot spots of one kernel in an `ncu --page source --csv` export: python tools/ncu_hot.py <src.csv> <kernel substring> [top n]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdrs = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r]
want = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
for hi, h in enumerate(hdrs):
    name = rows[h - 1][1] if h > 0 and len(rows[h - 1]) > 1 else ''
    if want not in name:
        continue
    hdr = rows[h]; end = hdrs[hi + 1] - 1 if hi + 1 < len(hdrs) else len(rows)
    body = rows[h + 1:end]
    i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    num = lambda r, i: int(r[i]) if i < len(r) and r[i].isdigit() else 0
    tot = sum(num(r, i_s) for r in body); ex = sum(num(r, i_ex) for r in body)
    print(name[:80], '| samples', tot, '| warp instr', ex, '| static', len(body))
    stalls = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
    agg = {c: sum(num(r, hdr.index(c)) for r in body) for c in stalls}
    print('  ' + '  '.join(f'{c[6:]}={v}' for c, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    ops = collections.Counter()
    for r in body:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[i_src]); ops[m.group(2).split('.')[0] if m else '?'] += num(r, i_ex)
    print('  ops: ' + '  '.join(f'{o}={n}' for o, n in ops.most_common(14)))
    idx = sorted(range(len(body)), key=lambda k: -num(body[k], i_s))[:top]
    il = hdr.index('stall_long_sb')
    for k in sorted(idx):
        r = body[k]; print(f'  {k:5d} samp {num(r, i_s):6d} long {num(r, il):6d} exec {num(r, i_ex):8d}  {r[i_src][:90]}')
    break
