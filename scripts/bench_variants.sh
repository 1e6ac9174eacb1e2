#!/bin/bash
# diagnostic: device-arm step time with / without the GEMM event instrumentation and the nvidia-smi sampler
for f in "--no-gemm-timing --no-clocks" "--no-clocks" "--no-gemm-timing" ""; do
  python bench.py --steps 10 --warmup 3 --no-cpu $f 2>/tmp/b.err > /tmp/b.json || tail -3 /tmp/b.err
  python - "$f" <<'PY'
import json, sys
d = json.loads(open('/tmp/b.json').read())
print(repr(sys.argv[1]), 'ms/step', round(d['ms_per_step'], 2), 'e2e ms/step', round(65536 / d['e2e']['value'] * 1e3, 2), 'other', round(d['other_mode']['ms_per_step'], 2))
PY
done
