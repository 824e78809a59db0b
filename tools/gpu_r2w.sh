#!/bin/bash
# round 2, call W: the other workloads with the final build: real-graph / size sweep, BA-2M stress graph, training step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python tools/config_sweep.py > gpurun_out/r2w_config_sweep.log 2>&1; cat gpurun_out/r2w_config_sweep.log
timeout 900 python bench.py --workload ba2m --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2w_bench_ba2m_n1.json 2> gpurun_out/r2w_bench_ba2m.err; echo "ba2m rc=$?"; cut -c1-400 gpurun_out/r2w_bench_ba2m_n1.json
timeout 300 python tools/train_timing.py > gpurun_out/r2w_train_timing.log 2>&1; cat gpurun_out/r2w_train_timing.log
timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2w_train_n1.json 2> gpurun_out/r2w_train_n1.err; echo "train rc=$?"; cut -c1-300 gpurun_out/r2w_train_n1.json
