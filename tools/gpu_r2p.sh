#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python tools/kernel_error_table.py 0 10 5 11 12 13 > gpurun_out/r2p_kernel_error_table.log 2>&1; cat gpurun_out/r2p_kernel_error_table.log
GNODE_STEP_KERNEL=13 timeout 600 python -m pytest tests/test_variants_gpu.py -m gpu -q -x -s -k "teacher" 2>&1 | grep -E "rhs err|passed|failed" > gpurun_out/r2p_rhs_k13.log; cat gpurun_out/r2p_rhs_k13.log
