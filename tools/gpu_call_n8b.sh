#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2n8b_bench_p2p.json 2> gpurun_out/r2n8b_bench_p2p.err; cut -c1-400 gpurun_out/r2n8b_bench_p2p.json; grep -v "^\*\|OMP\|^$\|NCCL version" gpurun_out/r2n8b_bench_p2p.err | tail -5
timeout 200 $TR tools/probe_p2p_modes.py 2>&1 | grep "^rank" | grep "variant 16\|items" | sort > gpurun_out/r2n8b_p2p_modes.log; grep "rank 0" gpurun_out/r2n8b_p2p_modes.log
timeout 200 $TR tests/p2p_worker.py 2>&1 | grep -c P2P_WORKER_OK
