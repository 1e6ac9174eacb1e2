"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel name: launches, total ms, share.
  python scripts/launch_summary.py gpurun_out/launches.csv [skip_first_n]"""
import collections, csv, io, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
while lines and not lines[0].startswith('"ID"'):
    lines.pop(0)
rows = [r for r in csv.DictReader(io.StringIO(''.join(lines))) if r['Metric Name'] == 'gpu__time_duration.sum']
rows = rows[int(sys.argv[2]) if len(sys.argv) > 2 else 0:]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r['Kernel Name'].split('(')[0][:64], [0, 0.0])
    a[0] += 1
    a[1] += float(r['Metric Value'].replace(',', '')) / 1e6
tot = sum(v[1] for v in agg.values())
print('%d launches, %.3f ms' % (len(rows), tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print('%-66s %5d %9.3f ms %5.1f%%  avg %.4f' % (k, c, t, 100 * t / tot, t / c))
