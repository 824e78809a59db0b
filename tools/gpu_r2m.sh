#!/bin/bash
# round 2, call M: lean decoder-backward kernel (parity + timing), ncu capture of the reverse sweep's tile kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 600 python -m pytest tests/test_backward_gpu.py tests/test_trials_gpu.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2m_pytest.log; cat gpurun_out/r2m_pytest.log
timeout 300 python tools/train_timing.py > gpurun_out/r2m_train_timing.log 2>&1; cat gpurun_out/r2m_train_timing.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r2m_launches_training_step.csv \
    python tools/bwd_once.py 8 > gpurun_out/r2m_ncu_launch.log 2>&1; echo "launch-list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_vjp2_kernel -s 20 -c 1 -f -o gpurun_out/prof_r2m_vjp2 \
    python tools/bwd_once.py 8 > gpurun_out/r2m_ncu_vjp2.log 2>&1; echo "ncu vjp2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_dec_kernel -s 20 -c 1 -f -o gpurun_out/prof_r2m_dec \
    python tools/bwd_once.py 8 > gpurun_out/r2m_ncu_dec.log 2>&1; echo "ncu dec rc=$?"
