#!/bin/bash
# round 2, call Z: I_{k+1} stored by TMA from the operand tile: parity + A/B (separate processes, env switch read once)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py tests/test_edge_cases_gpu.py tests/test_trials_gpu.py tests/test_backward_gpu.py -m gpu -q -x -k "not ba2m and not maxtime80" 2>&1 | tail -5 > gpurun_out/r2z_pytest.log; cat gpurun_out/r2z_pytest.log
for i in 1 2; do
GNODE_NO_TMA_ISTORE=1 timeout 300 python tools/ab_bench.py --trials 128 --rounds 3 kernel=5 > gpurun_out/r2z_ab_stg_$i.log 2>&1; cat gpurun_out/r2z_ab_stg_$i.log
timeout 300 python tools/ab_bench.py --trials 128 --rounds 3 kernel=5 > gpurun_out/r2z_ab_tma_$i.log 2>&1; cat gpurun_out/r2z_ab_tma_$i.log
done
