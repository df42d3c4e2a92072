"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of ONE resident bench step."""
import collections, csv, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
def ms(x):
    v = float(x['Metric Value'].replace(',', '')); u = x['Metric Unit']
    return v / 1e6 if u == 'ns' else v / 1e3 if u == 'us' else v
idx = [i for i, x in enumerate(rows) if 'mc_seed_kernel' in x['Kernel Name']]
start = idx[1] if len(idx) > 1 else idx[0]          # warm-up 1 + the timed resident step
if "--skip-for" in sys.argv:      # how many launches matching the regex precede the timed pass (for ncu --launch-skip)
    import re
    rx = re.compile(sys.argv[sys.argv.index("--skip-for") + 1])
    print(sum(1 for x in rows[:start] if rx.search(x['Kernel Name'])))
    sys.exit(0)
agg = collections.OrderedDict(); tot = 0
for x in rows[start:]:
    if x['Kernel Name'].startswith(('mc_seedcap', 'mc_profsum')): break
    k = x['Kernel Name'].split('(')[0].replace('void ', '')[:40]; a = agg.setdefault(k, [0, 0.0, []]); a[0] += 1; a[1] += ms(x); a[2].append(round(ms(x), 3)); tot += ms(x)
print("kernel                                     launches   total ms   share   per launch")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-42s %5d %10.3f %6.1f%%   %s" % (k, a[0], a[1], 100 * a[1] / tot, a[2] if a[0] <= 6 else ''))
print("sum of kernel time in the step: %.3f ms" % tot)
