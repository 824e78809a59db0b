#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python tools/diag_bench_graph.py > gpurun_out/r2f_diag.log 2>&1; cat gpurun_out/r2f_diag.log
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_trials_gpu.py tests/test_scripts_gpu.py tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -12
timeout 300 python tools/train_timing.py > gpurun_out/r2f_train_timing.log 2>&1; cat gpurun_out/r2f_train_timing.log
timeout 300 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2f_train_n1.json 2>gpurun_out/r2f_train.err; cat gpurun_out/r2f_train_n1.json
