"""Prints the launches of an ncu launch list in order (name, ms), for the LAST `n` launches.
  python scripts/launch_sequence.py gpurun_out/launches.csv 80"""
import csv, io, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
while lines and not lines[0].startswith('"ID"'):
    lines.pop(0)
rows = [r for r in csv.DictReader(io.StringIO(''.join(lines))) if r['Metric Name'] == 'gpu__time_duration.sum']
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for r in rows[-n:]:
    print('%-60s %9.4f ms  grid %s' % (r['Kernel Name'].split('(')[0][:60], float(r['Metric Value'].replace(',', '')) / 1e6, r.get('Grid Size', '')))
