#!/bin/bash
# round 2, call X: training forward (trajectory + auxiliary storage) with and without the gather work items
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
for i in 1 2; do
echo default; timeout 300 python tools/train_timing.py 2>&1 | grep epinions
echo no_quads; GNODE_NO_QUADS=1 timeout 300 python tools/train_timing.py 2>&1 | grep epinions
echo no_pair_sort; GNODE_NO_PAIR_SORT=1 timeout 300 python tools/train_timing.py 2>&1 | grep epinions
done
