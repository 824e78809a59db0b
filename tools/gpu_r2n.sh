#!/bin/bash
# round 2, call N: degree-sorted row pairs in the gather (parity + A/B), timing-only ablations of the GEMMs and the
# operand packing pass (library built with -DGNODE_ABLATIONS)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py tests/test_edge_cases_gpu.py -m gpu -q -x -k "not ba2m and not maxtime80" 2>&1 | tail -3 > gpurun_out/r2n_pytest.log; cat gpurun_out/r2n_pytest.log
timeout 300 python tools/ab_bench.py --trials 128 --rounds 3 kernel=5 kernel=7 kernel=8 kernel=9 > gpurun_out/r2n_ab_ablations.log 2>&1; cat gpurun_out/r2n_ab_ablations.log
for i in 1 2; do
GNODE_NO_PAIR_SORT=1 timeout 300 python tools/ab_bench.py --trials 128 --rounds 2 kernel=5 > gpurun_out/r2n_ab_nosort_$i.log 2>&1; cat gpurun_out/r2n_ab_nosort_$i.log
timeout 300 python tools/ab_bench.py --trials 128 --rounds 2 kernel=5 > gpurun_out/r2n_ab_sort_$i.log 2>&1; cat gpurun_out/r2n_ab_sort_$i.log
done
