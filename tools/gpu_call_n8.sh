#!/bin/bash
# round 2: the 8-GPU measurements (cfg3 = 10M users x 1M items, d=128) in one call
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/p2p_worker.py > gpurun_out/r2n8_worker.log 2>&1; grep -c P2P_WORKER_OK gpurun_out/r2n8_worker.log; grep -E "Error|assert|Mismatch|rank" gpurun_out/r2n8_worker.log | head -20
timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2n8_bench_p2p.json 2> gpurun_out/r2n8_bench_p2p.err; cut -c1-300 gpurun_out/r2n8_bench_p2p.json
timeout 200 $TR tools/probe_p2p.py 2>&1 | grep "^rank [07]"
timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 3 --layout user_sharded --no-secondary > gpurun_out/r2n8_bench_user_sharded.json 2> gpurun_out/r2n8_bench_user_sharded.err; cut -c1-300 gpurun_out/r2n8_bench_user_sharded.json
timeout 500 $TR bench.py --gpus 8 --steps 5 --warmup 3 --layout item_sharded > gpurun_out/r2n8_bench_item_sharded_nccl.json 2> gpurun_out/r2n8_bench_item_sharded_nccl.err; cut -c1-300 gpurun_out/r2n8_bench_item_sharded_nccl.json
tail -3 gpurun_out/r2n8_bench_*.err | grep -v "^\*\|OMP\|^$" | tail -12
