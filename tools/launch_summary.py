"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: python tools/launch_summary.py gpurun_out/launches.csv profiles/out.md ["title"]"""
import collections, csv, sys
src, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else src
lines = [l for l in open(src) if not l.startswith('==')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)
    a = agg.setdefault(row['Kernel Name'][:90], [0, 0.0, []]); a[0] += 1; a[1] += v; a[2].append(v)
tot = sum(a[1] for a in agg.values())
res = [f"# {title}", "", f"source: `{src}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)", "",
       "| kernel | launches | total us | share | avg us | min us | max us |", "|---|---|---|---|---|---|---|"]
for k, (n, t, l) in sorted(agg.items(), key=lambda x: -x[1][1]):
    res.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.1f} | {min(l):.1f} | {max(l):.1f} |")
open(out, 'w').write("\n".join(res) + "\n")
print("\n".join(res))
