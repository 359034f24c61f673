#!/usr/bin/env bash
# The GPU-side checks of this repo in the order they were run during round 1 (each line is what went inside one
# `gpurun -- '<command>'` call; outputs under gpurun_out/, summaries copied to profiles/ by hand).
# Usage on a B200 box, from the repo root:   bash tools/gpu_checklist.sh [quick|full]
set -u
mode=${1:-quick}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5                      # parity suite (XPASS = experimental path validated)
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
[ "$mode" = quick ] && exit 0
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json
# experimental paths (off by default): first device validation
timeout 300 python tests/pointwise_worker.py
timeout 300 python tests/rerank_staged_worker.py
B200REC_RERANK=2 B200REC_TC_TIME=1 REPS=3 timeout 300 python tools/probe_tc_call.py 2>&1 | tail -6   # staged re-rank timing
# per-kernel evidence
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_candidate -c 1 -o gpurun_out/tc_candidate \
    python bench.py --no-cpu --steps 5 > gpurun_out/ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bpr_step -s 5 -c 1 -o gpurun_out/bpr_step \
    python bench.py --no-cpu --steps 5 > gpurun_out/ncu_bpr.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --no-cpu --steps 20 --warmup 3 > gpurun_out/launches.log 2>&1
python profiles/summarize_ncu.py gpurun_out/tc_candidate.ncu-rep gpurun_out/tc_candidate_ncu.md "tcgen05 candidate kernel"
python profiles/summarize_ncu.py gpurun_out/bpr_step.ncu-rep gpurun_out/bpr_step_ncu.md "fused BPR step"
# secondary configs
timeout 600 python tools/bench_lightgcn.py > gpurun_out/lightgcn_cfg4.json      # cfg4
timeout 600 python tools/probe_cfg5.py 2>&1 | grep "^d="                         # cfg5 slice
timeout 300 python tools/probe_tc.py 2>&1 | grep -v "^\[b200"                    # random / lognormal tables, K=10/100
# multi-GPU (N = 2, 4, 8):  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
#     --master-port 29511 bench.py --gpus N --steps 20 --warmup 3 [--exchange diff|buffer]
