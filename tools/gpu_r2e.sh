#!/bin/bash
# round 2, call E: A/B deferred store wait (5) vs barrier (6) vs round-1 kernel (3); full GPU tests after the prune
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python tools/ab_bench.py --trials 128 --rounds 3 kernel=6 kernel=5 kernel=3 > gpurun_out/r2e_ab.log 2>&1; cat gpurun_out/r2e_ab.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2e_pytest.log; tail -8 gpurun_out/r2e_pytest.log
