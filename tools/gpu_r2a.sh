#!/bin/bash
# round 2, call A: GPU tests with the new goldens + stream kernels, A/B of the step kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -150 > gpurun_out/r2a_pytest.log
timeout 600 python tools/ab_bench.py --trials 64 --rounds 3 kernel=3 kernel=5 kernel=6 > gpurun_out/r2a_ab.log 2>&1
tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_ab.log
