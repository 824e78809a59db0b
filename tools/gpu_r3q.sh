#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 80 python tools/train_fwd_probe.py 8 > gpurun_out/r3q_train_fwd_probe.log 2>&1; cat gpurun_out/r3q_train_fwd_probe.log
