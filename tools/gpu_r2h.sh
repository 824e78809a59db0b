#!/bin/bash
# round 2, call H (third session): state of the restored tree: full GPU tests, bench, launch list, ncu capture of the stream kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2h_pytest.log; tail -6 gpurun_out/r2h_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; cat gpurun_out/r2h_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2h_launches.csv \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r2h_ncu_launch.log 2>&1; echo "launch-list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_stream_kernel -s 60 -c 1 -f -o gpurun_out/prof_r2h \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r2h_ncu_full.log 2>&1; echo "ncu-full rc=$?"
ls -la gpurun_out | tail -8
