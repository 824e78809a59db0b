#!/bin/bash
# round 2, call U: bench line, launch list and ncu capture of the step kernel with the stacked N = 160 operand + truncation split
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; cat gpurun_out/r2u_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2u_launches.csv \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r2u_ncu_launch.log 2>&1; echo "launch-list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_stream_kernel -s 60 -c 1 -f -o gpurun_out/prof_r2u \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r2u_ncu_full.log 2>&1; echo "ncu-full rc=$?"
ls -la gpurun_out | tail -5
