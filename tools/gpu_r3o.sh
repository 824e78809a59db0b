#!/bin/bash
# round 2, call 3O: 2-GPU sanity of the final build (torchrun, NCCL process group, trials sharded over the ranks)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3o_bench_n2.json 2> gpurun_out/r3o_bench_n2.err; echo "rc=$?"
cut -c1-260 gpurun_out/r3o_bench_n2.json; tail -3 gpurun_out/r3o_bench_n2.err
