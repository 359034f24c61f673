#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 100 $TR tools/probe_p2p_modes.py 2>&1 | grep "^rank 0" | grep "variant 16" > gpurun_out/r2n8c_p2p_modes_symm.log; cat gpurun_out/r2n8c_p2p_modes_symm.log
timeout 200 $TR bench.py --gpus 8 --layout p2p --steps 10 --warmup 3 --no-legs > gpurun_out/r2n8c_bench_p2p_symm.json 2> gpurun_out/r2n8c_bench_p2p_symm.err; cut -c1-330 gpurun_out/r2n8c_bench_p2p_symm.json; grep -v "^\*\|OMP\|^$\|NCCL version" gpurun_out/r2n8c_bench_p2p_symm.err | tail -4
