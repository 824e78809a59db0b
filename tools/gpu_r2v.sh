#!/bin/bash
# round 2, call V: four-row gather items (rows of at most 6 neighbours, two per half-warp): parity + A/B (separate processes:
# the items are built at batch creation)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x -k "not ba2m and not maxtime80" 2>&1 | tail -5 > gpurun_out/r2v_pytest.log; cat gpurun_out/r2v_pytest.log
for i in 1 2; do
GNODE_NO_QUADS=1 timeout 300 python tools/ab_bench.py --trials 128 --rounds 3 kernel=5 > gpurun_out/r2v_ab_noquads_$i.log 2>&1; cat gpurun_out/r2v_ab_noquads_$i.log
timeout 300 python tools/ab_bench.py --trials 128 --rounds 3 kernel=5 > gpurun_out/r2v_ab_quads_$i.log 2>&1; cat gpurun_out/r2v_ab_quads_$i.log
done
