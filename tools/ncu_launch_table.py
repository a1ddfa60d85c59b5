#!/usr/bin/env python3
"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --steps 1 --warmup 1` into
(a) one CSV row per launch of the middle (timed) step and (b) a per-kernel share table in markdown.
usage: ncu_launch_table.py launches.csv out_prefix [n_steps_in_capture]"""
import collections
import csv
import io
import re
import sys

txt = open(sys.argv[1]).read()
rows = [x for x in csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])) if x['Metric Name'] == 'gpu__time_duration.sum']
# A batch call starts with VtxProblemFn over the whole batch (the launch with the largest grid of that kernel): the capture holds the
# warm-up step, the timed step and the e2e step, followed by bench.py's single-path latency calls (tiny launches, not part of a step).
def grid0(x):
    return int(x['Grid Size'].strip('()').split(',')[0].strip())


vtx = [i for i, x in enumerate(rows) if 'VtxProblemFn' in x['Kernel Name']]
big = max(grid0(rows[i]) for i in vtx)
starts = [i for i in vtx if grid0(rows[i]) == big]
nsteps = len(starts)
n = starts[1] - starts[0]
step = rows[starts[1]:starts[1] + n]


def short(name):
    name = re.sub(r'\(.*', '', name).replace('void <unnamed>::', '').replace('tg::', '')
    return re.sub(r'cub::(\w+)<.*', r'cub::\1', name)


with open(sys.argv[2] + '_step.csv', 'w') as f:
    f.write('launch,kernel,grid,block,duration_us\n')
    for i, x in enumerate(step):
        grid = x['Grid Size'].strip('()').split(',')[0].strip()
        block = x['Block Size'].strip('()').split(',')[0].strip()
        f.write('%d,"%s",%s,%s,%.2f\n' % (i, short(x['Kernel Name']), grid, block, float(x['Metric Value'].replace(',', '')) / 1e3))
agg = collections.defaultdict(lambda: [0, 0.0])
for x in step:
    k = short(x['Kernel Name'])
    agg[k][0] += 1
    agg[k][1] += float(x['Metric Value'].replace(',', '')) / 1e6
tot = sum(v[1] for v in agg.values())
with open(sys.argv[2] + '_summary.md', 'w') as f:
    f.write(f"ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`) of `python bench.py --steps 1 --warmup 1 --no-cpu --no-profile`: "
            f"{len(rows)} launches in the capture = {nsteps} steps (warm-up, timed, e2e) of {n} launches each, then the single-path latency calls; the table is the timed step. "
            f"Durations under ncu are serialised and cold-cache: compare SHARES with bench.py's `kernel_profile`, not absolutes.\n\n")
    f.write(f"| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.0005:
            continue
        f.write(f"| `{k}` | {v[0]} | {v[1]:.2f} | {v[1] / tot:.3f} |\n")
    f.write(f"| total | {len(step)} | {tot:.2f} | 1 |\n")
print(open(sys.argv[2] + '_summary.md').read())
