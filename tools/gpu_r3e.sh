#!/bin/bash
# round 2, call 3E: full GPU suite, bench line, launch list and ncu capture with the descriptor-fed encoder fused into step 0
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r3e_pytest.log; cat gpurun_out/r3e_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc=$?"; cat gpurun_out/r3e_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3e_launches.csv \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r3e_ncu_launch.log 2>&1; echo "launch-list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_stream_kernel -s 58 -c 3 -f -o gpurun_out/prof_r3e \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r3e_ncu_full.log 2>&1; echo "ncu-full rc=$?"
ls -la gpurun_out | tail -5
