#!/bin/bash
# round 2, call D: ablation of the two overlaps of the stream kernel, parity of the new default, real-graph sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python tools/ab_bench.py --trials 128 --rounds 3 kernel=6 kernel=7 kernel=8 kernel=5 > gpurun_out/r2d_ab.log 2>&1; cat gpurun_out/r2d_ab.log
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_parity_gpu.py tests/test_edge_cases_gpu.py tests/test_trials_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2d_pytest.log; tail -5 gpurun_out/r2d_pytest.log
timeout 900 python tools/config_sweep.py > gpurun_out/r2d_config_sweep.log 2>&1; cat gpurun_out/r2d_config_sweep.log
