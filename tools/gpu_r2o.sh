#!/bin/bash
# round 2, call O: N = 160 stacked-operand GEMM (parity + A/B against the N = 80 operands), 3xTF32 and truncation-split variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x -k "not ba2m and not maxtime80" 2>&1 | tail -5 > gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_pytest.log
timeout 600 python tools/ab_bench.py --trials 128 --rounds 3 variant=0 kernel=5 kernel=10 kernel=11 kernel=12 kernel=13 > gpurun_out/r2o_ab_gemm_variants.log 2>&1; cat gpurun_out/r2o_ab_gemm_variants.log
