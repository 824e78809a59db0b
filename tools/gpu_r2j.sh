#!/bin/bash
# round 2, call J: A/B of the stream kernel without the gather -> update barrier (kernel 7), BA-2M / maxTime-80 parity,
# reverse sweep with the prefetching tile kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python tools/ab_bench.py --trials 128 --rounds 3 kernel=5 kernel=7 > gpurun_out/r2j_ab.log 2>&1; cat gpurun_out/r2j_ab.log
timeout 900 python -m pytest tests/test_edge_cases_gpu.py -m gpu -q -s -k "ba2m or maxtime80" 2>&1 | grep -E "BA-2M|maxTime 80|passed|failed|Error" > gpurun_out/r2j_pytest_new.log; cat gpurun_out/r2j_pytest_new.log
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_trials_gpu.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2j_pytest.log; cat gpurun_out/r2j_pytest.log
timeout 300 python tools/train_timing.py > gpurun_out/r2j_train_timing.log 2>&1; cat gpurun_out/r2j_train_timing.log
