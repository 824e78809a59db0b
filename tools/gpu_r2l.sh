#!/bin/bash
# round 2, call L: BA-2M stress graph (bench line + ncu DRAM bytes of one step launch), real-graph sweep with the final
# build, reverse-sweep tile-kernel test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_pytest.log
timeout 900 python bench.py --workload ba2m --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2l_bench_ba2m_n1.json 2> gpurun_out/r2l_bench_ba2m.err; echo "ba2m rc=$?"; cat gpurun_out/r2l_bench_ba2m_n1.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_stream_kernel -s 20 -c 1 -f -o gpurun_out/prof_r2l_ba2m \
    python bench.py --workload ba2m --trials 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2l_ncu_ba2m.log 2>&1; echo "ncu rc=$?"
timeout 900 python tools/config_sweep.py > gpurun_out/r2l_config_sweep.log 2>&1; cat gpurun_out/r2l_config_sweep.log
