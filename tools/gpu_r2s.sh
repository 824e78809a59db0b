#!/bin/bash
# round 2, call S: default numerics = N = 160 stacked operand, 4 terms, truncation split: whole GPU suite, A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2s_pytest.log; cat gpurun_out/r2s_pytest.log
timeout 600 python tools/ab_bench.py --trials 128 --rounds 3 kernel=10 kernel=5 kernel=7 > gpurun_out/r2s_ab.log 2>&1; cat gpurun_out/r2s_ab.log
