"""Per-kernel counts of the SASS mnemonics that identify the Blackwell-native paths (B200_PROFILING.md): UTC*MMA (tcgen05.mma),
LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA loads), DMMA (FP64 mma.sync), LDGSTS (cp.async), IDP4A, plus totals.
  python scripts/sass_counts.py tgp/pytorch_b200/libtgp_b200.so > profiles/r02_sass_counts.txt"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else 'tgp/pytorch_b200/libtgp_b200.so'
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
KEYS = ['UTCIMMA', 'UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'DMMA', 'HMMA', 'LDGSTS', 'IDP', 'DFMA', 'SYNCS', 'UTCBAR']
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        counts.setdefault(cur, collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        counts[cur]['total'] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print('# SASS mnemonic counts per kernel of %s (cuobjdump -sass, sm_100a)' % so)
print('%-62s %7s  %s' % ('kernel', 'instrs', 'Blackwell / tensor-path mnemonics'))
for k, c in counts.items():
    tags = ' '.join('%s=%d' % (n, c[n]) for n in KEYS if c[n])
    print('%-62s %7d  %s' % (k[:62], c['total'], tags))
