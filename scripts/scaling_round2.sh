#!/bin/bash
# 8-GPU box: weak and strong scaling of the headline mode, plus the multi-rank parity run (scripts/dist_parity.py).
mkdir -p gpurun_out
run() { # n scaling tag extra
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $1 --steps 10 --warmup 3 --no-other-mode --scaling $2 $4 > gpurun_out/scale_$3.json 2> gpurun_out/scale_$3.err
  python -c "
import json; d=json.loads(open('gpurun_out/scale_$3.json').read().strip().splitlines()[-1]); print('$3', d['n_gpus'], d['scaling'], d['compute'], 'ms/step %.3f' % d['ms_per_step'], 'rows/s %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'dist_parity', d.get('dist_parity', {}).get('max_rel_err'))" || tail -5 gpurun_out/scale_$3.err
}
run 8 weak weak8
run 8 strong strong8
run 4 strong strong4
run 2 strong strong2
[ -n "$TGP_SCALE_F64" ] && run 8 strong strong8_f64 "--compute f64"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 scripts/dist_parity.py 2>&1 | grep "DIST_PARITY\|world 8 sync 0" | tail -9
