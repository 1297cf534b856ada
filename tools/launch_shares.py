"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X`):
    python tools/launch_shares.py X.csv [skip_setup]
Prints launches, total and mean duration and the share of the listed time per kernel (the per-launch times of such a
pass are serialised and cold-cache: the SHARES are what is compared with the CUDA-event split bench.py prints)."""
import collections
import csv
import sys

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
t, n = collections.defaultdict(float), collections.Counter()
for r in rows:
    k = r['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
    t[k] += float(r['Metric Value'])
    n[k] += 1
tot = sum(t.values())
print("%d launches, %.3f ms listed" % (len(rows), tot / 1e6))
for k, v in sorted(t.items(), key=lambda kv: -kv[1]):
    print("%-44s %6d launches %12.1f us %6.2f %%  mean %9.1f us" % (k, n[k], v / 1e3, 100 * v / tot, v / n[k] / 1e3))
