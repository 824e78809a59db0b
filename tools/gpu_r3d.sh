#!/bin/bash
# round 2, call 3D: hub-relay threshold 96 for latency-bound batches (at most one tile per pipeline): parity, then A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_edge_cases_gpu.py tests/test_variants_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r3d_pytest.log; cat gpurun_out/r3d_pytest.log
for i in 1 2; do
  echo "== threshold 512 (GNODE_HUB_DEG_512=1), pass $i"; GNODE_HUB_DEG_512=1 timeout 300 python tools/config_sweep.py --small 2>&1 | grep -v "^$"
  echo "== threshold 96 for latency-bound batches (default), pass $i"; timeout 300 python tools/config_sweep.py --small 2>&1 | grep -v "^$"
done > gpurun_out/r3d_ab_hub_threshold.log 2>&1
cat gpurun_out/r3d_ab_hub_threshold.log
