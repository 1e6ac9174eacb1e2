"""Turns gpurun_out/{launches_*.csv, *.ncu-rep} into the small tracked summaries under profiles/.

  python scripts/summarize_profiles.py launches gpurun_out/launches_r01.csv profiles/r01_launches.md
  python scripts/summarize_profiles.py ncu gpurun_out/gemm_r01a.ncu-rep profiles/r01_gemm_ncu.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum']


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith('==')]
    while lines and not lines[0].startswith('"ID"'):
        lines.pop(0)
    rows = [r for r in csv.DictReader(io.StringIO(''.join(lines))) if r['Metric Name'] == 'gpu__time_duration.sum']
    names = [r['Kernel Name'] for r in rows]
    marks = [i for i, n in enumerate(names) if 'k_transform_params' in n]
    # a bench step = [k_transform_params .. next k_transform_params); take the last complete device-arm step
    # (scripts/one_step.py runs exactly two steps: take the first, complete one)
    s, e = (marks[-3], marks[-2]) if len(marks) >= 3 else (marks[0], marks[1])
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[s:e]:
        t = float(r['Metric Value'].replace(',', '')) / 1e6
        a = agg.setdefault(r['Kernel Name'].split('(')[0][:70], [0, 0.0])
        a[0] += 1; a[1] += t; tot += t  # noqa: E702
    with open(dst, 'w') as f:
        f.write('# Launch list of ONE bench step (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n')
        f.write('Source: `%s` (launches %d..%d).  Times are cold-cache and serialised: compare shares.\n\n' % (src, s, e))
        f.write('| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n')
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.3f | %.1f%% |\n' % (n, c, t, 100 * t / tot))
        f.write('| **sum** | %d | %.3f | 100%% |\n' % (e - s, tot))
    print(open(dst).read())


def ncu(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, 'w') as f:
        f.write('# ncu --set full summary (`%s`)\n\n' % src)
        f.write('| metric | unit | ' + ' | '.join('`%s` grid %s' % (r[hdr.index('Kernel Name')][:40], r[hdr.index('launch__grid_size')]) for r in data) + ' |\n')
        f.write('|---|---|' + '---:|' * len(data) + '\n')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write('| %s | %s | ' % (k, units[i]) + ' | '.join(r[i] for r in data) + ' |\n')
        for i, h in enumerate(hdr):
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                try:
                    if max(float(r[i]) for r in data) >= 0.4:
                        f.write('| %s | ratio | ' % h.replace('smsp__average_warps_issue_stalled_', 'stall: ').replace('_per_issue_active.ratio', '')
                                + ' | '.join('%.2f' % float(r[i]) for r in data) + ' |\n')
                except ValueError:
                    pass
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'ncu': ncu}[sys.argv[1]](sys.argv[2], sys.argv[3])
