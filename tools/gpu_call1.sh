#!/bin/bash
# round 2, call 1: state check + BPR ablations + sanitizer + spmm profile
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2c1_pytest.log; cat gpurun_out/r2c1_pytest.log
B200REC_LIB=$PWD/gpurun_build/libb200rec_abl.so timeout 600 python tools/probe_bpr_ablate.py > gpurun_out/r2c1_ablate.log 2>&1; tail -50 gpurun_out/r2c1_ablate.log
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/r2c1_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -q -x -k "test_fused_step_with_collisions_is_close or test_tiny_sgd_exact_trajectory or (test_tc_equals_exact and 300)" 2>&1 | tail -3
tail -5 gpurun_out/r2c1_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --log-file gpurun_out/r2c1_racecheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -q -x -k "test_fused_step_with_collisions_is_close or test_tiny_sgd_exact_trajectory or (test_tc_equals_exact and 300)" 2>&1 | tail -3
tail -5 gpurun_out/r2c1_racecheck.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm -s 6 -c 2 -o gpurun_out/r2c1_spmm python tools/bench_lightgcn.py > gpurun_out/r2c1_ncu_spmm.log 2>&1; tail -2 gpurun_out/r2c1_ncu_spmm.log
