#!/usr/bin/env python
"""Condenses one GPU visit (gpurun_out/*_<tag>.*) into tracked files under profiles/:
    python profiles/summarise.py <tag>
writes profiles/<tag>_bench_c3.json, <tag>_timeline_c3.txt, <tag>_launches_c3_summary.txt, <tag>_ncu_keys.txt, <tag>_ncu_hot.txt"""
import collections, csv, os, shutil, subprocess, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = lambda n: os.path.join(root, 'gpurun_out', n)
o = lambda n: os.path.join(root, 'profiles', n)
for src, dst in ((f'bench_c3_{tag}.json', f'{tag}_bench_c3.json'), (f'timeline_c3_{tag}.txt', f'{tag}_timeline_c3.txt')):
    if os.path.exists(g(src)):
        shutil.copy(g(src), o(dst))
lc = g(f'launches_c3_{tag}.csv')
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr or len(r) <= vi or r[ki] == 'Kernel Name':
            continue
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        tot[r[ki]][0] += 1; tot[r[ki]][1] += v
    allt = sum(v[1] for v in tot.values())
    with open(o(f'{tag}_launches_c3_summary.txt'), 'w') as f:
        f.write(f'# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline\n')
        f.write('# per-launch times are cold-cache and serialised: compare SHARES (ns)\n')
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{k[:70]:70s} n={n:4d} avg_us={t / n / 1e3:9.1f} share={100 * t / allt:5.1f}%\n')
raw, src = g(f'prof_{tag}_raw.csv'), g(f'prof_{tag}_src.csv')
if os.path.exists(raw):
    open(o(f'{tag}_ncu_keys.txt'), 'w').write(subprocess.run([sys.executable, os.path.join(root, 'tools', 'ncu_keys.py'), raw], capture_output=True, text=True).stdout)
if os.path.exists(src):
    out = ''
    for k in ('states_specialised', 'control_rows', 'sample_rollouts', 'rollout_weights', 'weighted_update'):
        out += subprocess.run([sys.executable, os.path.join(root, 'tools', 'ncu_hot.py'), src, k, '24'], capture_output=True, text=True).stdout + '\n'
    open(o(f'{tag}_ncu_hot.txt'), 'w').write(out)
print(sorted(n for n in os.listdir(os.path.join(root, 'profiles')) if n.startswith(tag)))
