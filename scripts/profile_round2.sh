#!/bin/bash
# Round-2 evidence, run under gpurun (one GPU): launch lists of one training step per compute mode, ncu --set full captures of
# the dominant kernels, compute-sanitizer runs on the small fixtures.  Outputs land in gpurun_out/; scripts/summarize_round2.py
# turns them into the tracked summaries under profiles/.
mkdir -p gpurun_out
PART=${1:-all}   # gpurun merges at most 64 MiB per call: run part 1 and part 2 in separate calls
if [ "$PART" = 1 ] || [ "$PART" = all ]; then
for m in i8crt f64; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${m}_step.csv python scripts/one_step_mode.py $m > gpurun_out/one_step_$m.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"gemm_i8_mod|k_crt_combine|k_rbf_residues|k_abbar_stats|k_abbar_residues|k_wt_residues|k_kernel_grads|k_row_quad" -c 20 -o gpurun_out/r02_i8crt -f python scripts/one_step_mode.py i8crt > gpurun_out/ncu_i8crt.log 2>&1
fi
if [ "$PART" = 2 ] || [ "$PART" = all ]; then
ncu --set full --clock-control none --import-source on -k regex:"gemm_f64_kernel" -s 46 -c 4 -o gpurun_out/r02_f64 -f python scripts/one_step_mode.py f64 > gpurun_out/ncu_f64.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_crt.py tests/test_gpu_flow_mlp.py tests/test_gpu_session.py tests/test_gpu_multiclass.py tests/test_gpu_kmeans.py tests/test_gpu_eval_bundle.py -q -x -k "not 1024 and not 2048 and not 4096 and not steptanh102" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "synth_reg_d8_m64_p1 or boston_svgp_p1 or test_prepare_cholesky_inverse_kl and 64" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -n 3 gpurun_out/sanitizer_memcheck.log; tail -n 3 gpurun_out/sanitizer_racecheck.log
fi
ls -la gpurun_out/*.ncu-rep
