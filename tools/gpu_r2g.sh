#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2g_pytest.log; tail -12 gpurun_out/r2g_pytest.log
timeout 600 python tools/ab_bench.py --trials 128 --rounds 2 kernel=3 kernel=5 > gpurun_out/r2g_ab.log 2>&1; cat gpurun_out/r2g_ab.log
timeout 300 python tools/train_timing.py > gpurun_out/r2g_train_timing.log 2>&1; cat gpurun_out/r2g_train_timing.log
GNODE_BWD_GRAPH=0 timeout 300 python tools/train_timing.py > gpurun_out/r2g_train_timing_nograph.log 2>&1; cat gpurun_out/r2g_train_timing_nograph.log
timeout 300 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2g_train_n1.json 2>gpurun_out/r2g_train.err; cat gpurun_out/r2g_train_n1.json
