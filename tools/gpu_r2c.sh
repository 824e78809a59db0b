#!/bin/bash
# round 2, call C: full GPU tests (trials / selection / L1 / MC / DP), the new 4096-trial bench, launch list, ncu capture
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -300 > gpurun_out/r2c_pytest.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2c_pytest.log | tail -30
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; cat gpurun_out/r2c_bench.json
timeout 300 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/r2c_train_n1.json 2> gpurun_out/r2c_train.err; echo "train rc=$?"; cat gpurun_out/r2c_train_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c_launches.csv \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r2c_ncu_launch.log 2>&1; echo "launch-list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_stream_kernel -s 60 -c 1 -f -o gpurun_out/prof_r2c \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r2c_ncu_full.log 2>&1; echo "ncu-full rc=$?"
ls -la gpurun_out | tail -8
