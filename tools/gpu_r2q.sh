#!/bin/bash
# round 2, call Q: default numerics = N = 160 stacked operand, 3xTF32, truncation split: whole GPU suite, error table, A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2q_pytest.log; cat gpurun_out/r2q_pytest.log
timeout 600 python tools/kernel_error_table.py 0 3 5 10 > gpurun_out/r2q_kernel_error_table.log 2>&1; cat gpurun_out/r2q_kernel_error_table.log
timeout 600 python tools/ab_bench.py --trials 128 --rounds 3 kernel=10 kernel=5 kernel=8 kernel=9 > gpurun_out/r2q_ab.log 2>&1; cat gpurun_out/r2q_ab.log
