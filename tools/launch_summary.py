"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step = the launches after one adamw_kernel
launch up to and including the next (round 1 lists: between two multi_cast launches)."""
import csv, re, collections, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = [(r['Kernel Name'], float(r['Metric Value'].replace(',', '')) / 1e3) for r in csv.DictReader(lines)]
idx = [i for i, (n, _) in enumerate(rows) if 'adamw_kernel' in n]
if len(idx) < 2:
    idx = [i - 1 for i, (n, _) in enumerate(rows) if 'multi_cast' in n]
if len(idx) < 2: sys.exit(f"need two step markers, found {idx}")
step = rows[idx[-2] + 1:idx[-1] + 1]
tot = sum(t for _, t in step)
print(f"launches {len(step)}  sum of kernel durations {tot:.1f} us")
acc = collections.defaultdict(lambda: [0, 0.0])
for n, t in step:
    n = re.sub(r'\(.*', '', n)[:100]
    acc[n][0] += 1; acc[n][1] += t
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for n, v in sorted(acc.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"| `{n}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |")
