#!/bin/bash
# accuracy / speed of the tensor-core mode vs the TMEM accumulation chunk length (diagnostic)
python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('f64      ms/step %.2f loss %.6f' % (d['ms_per_step'], d['final_loss']))"
for k in 32 64 128; do
  if [ $k = 128 ]; then unset TGP_B200_LIB; else export TGP_B200_LIB=$PWD/build/libtgp_k$k.so; fi
  python bench.py --compute tf32x3 --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('kchunk $k ms/step %.2f loss %.6f' % (d['ms_per_step'], d['final_loss']))"
  python -m pytest tests/test_gpu_tensorcore.py -q -s -k "gemm_matches" 2>&1 | grep -E "AssertionError: |passed|failed" | head -8
done
