#!/bin/bash
# round 2, call K (2 GPUs): NCCL data-parallel training check, rollout bench and training bench under torchrun, MN-major UMMA probe
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 ./tools/umma_mn_probe > gpurun_out/r2k_umma_mn_probe.log 2>&1; cat gpurun_out/r2k_umma_mn_probe.log
timeout 600 $TR --master-port 29511 tools/dp_check.py 2>&1 | grep -E "DP_CHECK|Error|error" > gpurun_out/r2k_dp_check_nccl.log; cat gpurun_out/r2k_dp_check_nccl.log
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err; echo "bench rc=$?"; cat gpurun_out/r2k_bench_n2.json
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --mode train --steps 10 --warmup 3 > gpurun_out/r2k_train_n2.json 2> gpurun_out/r2k_train_n2.err; echo "train rc=$?"; cat gpurun_out/r2k_train_n2.json
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --mode train --train-per-graph 64 --steps 10 --warmup 3 > gpurun_out/r2k_train64_n2.json 2> gpurun_out/r2k_train64_n2.err; echo "train64 rc=$?"; cat gpurun_out/r2k_train64_n2.json
timeout 300 python bench.py --mode train --train-per-graph 64 --steps 10 --warmup 3 > gpurun_out/r2k_train64_n1.json 2> gpurun_out/r2k_train64_n1.err; echo "train64 n1 rc=$?"; cat gpurun_out/r2k_train64_n1.json
tail -3 gpurun_out/r2k_bench_n2.err gpurun_out/r2k_train_n2.err
