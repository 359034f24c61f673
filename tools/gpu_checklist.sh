#!/usr/bin/env bash
# The GPU-side checks of this repo (each block is what went inside one `gpurun -- '<command>'` call; outputs under
# gpurun_out/, summaries copied to profiles/ by hand).  Usage on a B200 box, from the repo root:
#   bash tools/gpu_checklist.sh [quick|full]            (1 GPU)
# Multi-GPU (N = 2, 4, 8), see the bottom.
set -u
mode=${1:-quick}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5                        # parity suite
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
[ "$mode" = quick ] && exit 0
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json
# per-kernel evidence (ncu replays every launch ~40x: never quote a number printed under it)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bpr_step_group -s 8 -c 1 -o gpurun_out/bpr_step \
    python bench.py --no-cpu --no-legs --steps 5 > gpurun_out/ncu_bpr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_candidate -c 1 -o gpurun_out/tc_candidate \
    python bench.py --no-cpu --no-legs --steps 5 > gpurun_out/ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm -s 6 -c 2 -o gpurun_out/spmm \
    python tools/bench_lightgcn.py > gpurun_out/ncu_spmm.log 2>&1
python tools/ncu_hot.py gpurun_out/bpr_step.ncu-rep 0.03 > gpurun_out/bpr_step_ncu.md      # headline metrics + stall hot spots
# profiling-only build (ablation switches of the fused step, B200REC_TC_* diagnostics of the scoring path)
bash tools/build_ablate.sh
B200REC_LIB=$PWD/gpurun_build/libb200rec_abl.so timeout 600 python tools/probe_bpr_ablate.py | tail -40
timeout 300 python tools/probe_bpr_variants.py                                     # kernel-structure sweep (B200REC_STEP_VARIANT)
./gpurun_build/gather_bw 1000000 1000000                                           # random-row gather roofline (nvcc tools/gather_bw.cu)
timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -k "collisions or tiny_sgd" | tail -3
# ---- multi-GPU:  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511"
#   $TR bench.py --gpus N --steps 20 --warmup 3                  default layout p2p (item + user tables sharded, P2P-fused step)
#   $TR bench.py --gpus N --layout user_sharded | item_sharded   the NCCL-collective layouts at the same (cfg3) shape
#   $TR tests/p2p_worker.py ; $TR tests/dist_worker.py           real CUDA-IPC / VMM mappings and NCCL, any N
#   $TR tools/probe_p2p.py ; $TR tools/probe_p2p_modes.py        phase timing / chunk order / peer read+write anatomy
#   ./gpurun_build/peer_duplex ; ./gpurun_build/peer_multi ; ./gpurun_build/peer_under_load     NVLink microbenchmarks
