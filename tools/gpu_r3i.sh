#!/bin/bash
# round 2, call 3I: timing-only ablation: the step kernel without the hid(R) recurrence (W3 I'_k dots, butterfly, hid_r
# read-modify-write) -- the upper bound of what carrying R through the conserved sum would buy
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
GNODE_B200_LIB=$PWD/tools/_ab/libgnode_abl.so timeout 600 python tools/ab_bench.py --trials 128 --rounds 3 --reps 2 kernel=5 kernel=13 kernel=8 > gpurun_out/r3i_ab_no_hidr.log 2>&1
cat gpurun_out/r3i_ab_no_hidr.log
