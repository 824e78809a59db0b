#!/bin/bash
# round 2, call 3N: final build -- full GPU suite, bench line, launch list and ncu capture of the step kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r3n_pytest.log; cat gpurun_out/r3n_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r3n_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3n_launches.csv \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r3n_ncu_launch.log 2>&1; echo "launch-list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_stream_kernel -s 58 -c 2 -f -o gpurun_out/prof_r3n \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r3n_ncu_full.log 2>&1; echo "ncu-full rc=$?"
