#!/bin/bash
# round 2, call T: GEMM1 hi pass issued as soon as the S tile has landed: parity + A/B against the round-2h operands (kernel 10)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x -k "not ba2m and not maxtime80" 2>&1 | tail -5 > gpurun_out/r2t_pytest.log; cat gpurun_out/r2t_pytest.log
timeout 600 python tools/ab_bench.py --trials 128 --rounds 3 kernel=10 kernel=5 kernel=7 > gpurun_out/r2t_ab.log 2>&1; cat gpurun_out/r2t_ab.log
