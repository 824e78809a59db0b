#!/bin/bash
# round 2, call 3P: training timings and the latency-bound real-graph batches with the final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 60 python tools/train_timing.py > gpurun_out/r3p_train_timing.log 2>&1; cat gpurun_out/r3p_train_timing.log
timeout 60 python tools/config_sweep.py --small > gpurun_out/r3p_config_sweep_small.log 2>&1; cat gpurun_out/r3p_config_sweep_small.log
