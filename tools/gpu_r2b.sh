#!/bin/bash
# round 2, call B: low-bits probe, full GPU tests (new goldens, stream kernels, trials / selection / L1), A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
./tools/umma_lowbits_probe > gpurun_out/r2b_lowbits.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | tail -400 > gpurun_out/r2b_pytest.log
timeout 600 python tools/ab_bench.py --trials 64 --rounds 2 kernel=3 kernel=5 kernel=6 > gpurun_out/r2b_ab.log 2>&1
cat gpurun_out/r2b_lowbits.log | tail -20; grep -E "passed|failed|FAILED|Error" gpurun_out/r2b_pytest.log | tail -40; cat gpurun_out/r2b_ab.log
