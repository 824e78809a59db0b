#!/bin/bash
# round 2, call 3A: bwd_gz_kernel with the row metadata resolved one iteration ahead: parity + timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_trials_gpu.py tests/test_dp_gpu.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r3a_pytest.log; cat gpurun_out/r3a_pytest.log
timeout 300 python tools/train_timing.py > gpurun_out/r3a_train_timing.log 2>&1; cat gpurun_out/r3a_train_timing.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r3a_launches_training_step.csv \
    python tools/bwd_once.py 8 > gpurun_out/r3a_ncu_launch.log 2>&1; echo "launch-list rc=$?"
