#!/bin/bash
# round 2, call I: reverse sweep with the forward's auxiliary storage (I'_k, A I'_k): parity, A/B timing, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1200 python -m pytest tests/test_backward_gpu.py tests/test_trials_gpu.py tests/test_dp_gpu.py tests/test_scripts_gpu.py tests/test_sharding_gpu.py tests/test_variants_gpu.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2i_pytest.log; tail -8 gpurun_out/r2i_pytest.log
timeout 900 python -m pytest tests/test_edge_cases_gpu.py -m gpu -q -s -k "ba2m or maxtime80" 2>&1 | tail -25 > gpurun_out/r2i_pytest_new.log; tail -8 gpurun_out/r2i_pytest_new.log
timeout 300 python tools/train_timing.py > gpurun_out/r2i_train_timing_aux.log 2>&1; cat gpurun_out/r2i_train_timing_aux.log
GNODE_AUX_STORAGE=0 timeout 300 python tools/train_timing.py > gpurun_out/r2i_train_timing_noaux.log 2>&1; cat gpurun_out/r2i_train_timing_noaux.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r2i_launches_training_step.csv \
    python tools/bwd_once.py 8 > gpurun_out/r2i_ncu_launch.log 2>&1; echo "launch-list rc=$?"
timeout 300 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2i_train_n1.json 2>gpurun_out/r2i_train.err; cat gpurun_out/r2i_train_n1.json
