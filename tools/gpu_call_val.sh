#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2v_pytest.log; cat gpurun_out/r2v_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; cut -c1-1500 gpurun_out/r2v_bench.json; tail -5 gpurun_out/r2v_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bpr_step_group -s 8 -c 1 -o gpurun_out/r2v_bpr_group python bench.py --no-cpu --no-legs --steps 5 > gpurun_out/r2v_ncu_bpr.log 2>&1; tail -2 gpurun_out/r2v_ncu_bpr.log
