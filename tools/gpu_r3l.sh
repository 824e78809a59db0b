#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
GNODE_B200_LIB=$PWD/tools/_ab/libgnode_b200_r3j.so python tools/_ab/dump_modes.py gpurun_out/r3l_prev.pt 2>&1 | tail -3
python tools/_ab/dump_modes.py gpurun_out/r3l_new.pt 2>&1 | tail -3
python tools/_ab/cmp.py gpurun_out/r3l_prev.pt gpurun_out/r3l_new.pt | tee gpurun_out/r3l_cmp.log
