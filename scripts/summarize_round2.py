"""gpurun_out/ (scripts/profile_round2.sh) -> tracked summaries under profiles/."""
import collections, csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def launch_rows(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    while lines and not lines[0].startswith('"ID"'):
        lines.pop(0)
    return [r for r in csv.DictReader(io.StringIO(''.join(lines))) if r['Metric Name'] == 'gpu__time_duration.sum']


def launches(mode):
    rows = launch_rows(os.path.join(G, 'launches_%s_step.csv' % mode))
    names = [r['Kernel Name'] for r in rows]
    # scripts/one_step_mode.py: the measured step follows the LAST k_transform_params (one warm-up step precedes it)
    start = max(i for i, n in enumerate(names) if 'k_transform_params' in n)
    rows = rows[start:]
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        t = float(r['Metric Value'].replace(',', '')) / 1e6
        a = agg.setdefault(r['Kernel Name'].split('(')[0][:72], [0, 0.0])
        a[0] += 1; a[1] += t; tot += t  # noqa: E702
    with open(os.path.join(P, 'r02_launches_%s.md' % mode), 'w') as f:
        f.write('# Launch list of ONE training step at cfg4, compute mode `%s` (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n' % mode)
        f.write('`python scripts/one_step_mode.py %s` under ncu; %d launches of the measured step, %.3f ms serialised.  Times are cold-cache and\n'
                'serialised (no stream overlap): compare shares, not absolutes.\n\n' % (mode, len(rows), tot))
        f.write('| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|\n')
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.3f | %.1f%% | %.4f |\n' % (n, c, t, 100 * t / tot, t / c))


KEYS = [('gpu__time_duration.sum', 'time'), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'FP64 pipe active %'),
        ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
        ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 %'),
        ('dram__bytes_read.sum', 'DRAM read'), ('dram__bytes_write.sum', 'DRAM write'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('smsp__inst_executed.sum', 'warp instructions'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe_throttle'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait')]


def ncu(rep, dst, title):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    with open(dst, 'w') as f:
        f.write('# %s\n\n`ncu --set full --clock-control none --import-source on` of `python scripts/one_step_mode.py <mode>` (cfg4: 65536 rows,\n'
                'M = 1024, D = 8; 16384-row chunks), one capture per distinct kernel / shape.  Times under ncu replay are not bench values.\n\n' % title)
        for r in rows[2:]:
            name = r[idx['Kernel Name']].split('(')[0]
            key = (name, r[idx['launch__grid_size']], r[idx['gpu__time_duration.sum']][:3])
            if key in seen:
                continue
            seen.add(key)
            f.write('## `%s`  grid %s\n\n| metric | value |\n|---|---:|\n' % (name, r[idx['launch__grid_size']]))
            for k, label in KEYS:
                if k in idx:
                    f.write('| %s | %s %s |\n' % (label, r[idx[k]], units[idx[k]]))
            f.write('\n')


def traffic(reps):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel of each mode, tied to the hash of the profiled
    library (bench.py reports it only when the hash matches the library it runs)."""
    import hashlib, json

    def mean_bytes(rep, kernel):
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}

        def to_bytes(v, u):
            return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        vals = [to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']]) +
                to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
                for r in rows[2:] if kernel in r[idx['Kernel Name']]]
        return sum(vals) / len(vals), len(vals)
    sys.path.insert(0, ROOT)
    import bench
    h = bench.lib_hash()          # hash of the library sources (the binary is not byte-reproducible)
    out = {}
    if 'i8crt' in reps:
        t, n = mean_bytes(reps['i8crt'], 'gemm_i8_mod_kernel')
        out['cfg4:i8crt'] = {'lib_hash': h, 'traffic': t, 'launches_captured': n,
                             'source': 'profiles/r02_i8crt_ncu.md (ncu --set full of scripts/one_step_mode.py i8crt): mean DRAM read + write '
                                       'bytes per gemm_i8_mod_kernel launch; algorithmic operand + result bytes per launch (one 16384-row '
                                       'chunk, 15 planes: A 252 MB + W 31 MB + result 503 MB) = 0.79 GB'}
    if 'f64' in reps:
        t, n = mean_bytes(reps['f64'], 'gemm_f64_kernel')
        out['cfg4:f64'] = {'lib_hash': h, 'traffic': t, 'launches_captured': n,
                           'source': 'profiles/r02_f64_ncu.md (ncu --set full of scripts/one_step_mode.py f64): mean DRAM read + write bytes '
                                     'per gemm_f64_kernel launch (32768-row chunk: operand 268 MB + triangular weight 4 MB + result 268 MB '
                                     '= 0.54 GB algorithmic)'}
    json.dump(out, open(os.path.join(P, 'r02_roofline_traffic.json'), 'w'), indent=1)


if __name__ == '__main__':
    for m in ('i8crt', 'f64'):
        if os.path.exists(os.path.join(G, 'launches_%s_step.csv' % m)):
            launches(m)
    rep = os.path.join(G, 'r02_i8crt.ncu-rep')
    if os.path.exists(rep):
        ncu(rep, os.path.join(P, 'r02_i8crt_ncu.md'), 'ncu summaries of the kernels of compute mode i8crt (round 2)')
    rep64 = os.path.join(G, 'r02_f64.ncu-rep')
    if os.path.exists(rep64):
        ncu(rep64, os.path.join(P, 'r02_f64_ncu.md'), 'ncu summaries of gemm_f64_kernel, compute mode f64 (round 2)')
    traffic({k: v for k, v in (('i8crt', rep), ('f64', rep64)) if os.path.exists(v)})
    with open(os.path.join(P, 'r02_sanitizer.txt'), 'w') as f:
        for n in ('sanitizer_memcheck.log', 'sanitizer_racecheck.log'):
            p = os.path.join(G, n)
            if os.path.exists(p):
                f.write('==== %s (compute-sanitizer under gpurun, scripts/profile_round2.sh) ====\n' % n)
                f.write(''.join(open(p).readlines()[-8:]) + '\n')
