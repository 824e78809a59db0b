#!/bin/bash
# round 2, call Y (8 GPUs): rollout bench (configs[3], 4096 trials sharded over the ranks) and the data-parallel training
# step (configs[2], 64 instances per graph) under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2y_bench_n8.json 2> gpurun_out/r2y_bench_n8.err; echo "bench rc=$?"; cat gpurun_out/r2y_bench_n8.json
timeout 300 $TR --master-port 29522 bench.py --gpus 8 --mode train --train-per-graph 64 --steps 10 --warmup 3 > gpurun_out/r2y_train64_n8.json 2> gpurun_out/r2y_train64_n8.err; echo "train64 rc=$?"; cat gpurun_out/r2y_train64_n8.json
tail -n 3 gpurun_out/r2y_bench_n8.err; tail -n 3 gpurun_out/r2y_train64_n8.err
