#!/bin/bash
# round 2, call 3J: inference R through the conserved sum (hid(R_k) = W3 (S_0 + I_0 + R_0) - hid(S_k) - hid(I_k)): parity,
# error table on the goldens, A/B against the previous build (hid(R) advanced by linearity every step)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py tests/test_trials_gpu.py tests/test_edge_cases_gpu.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r3j_pytest.log; cat gpurun_out/r3j_pytest.log
(echo "== previous build"; GNODE_B200_LIB=$PWD/tools/_ab/libgnode_b200_r3h.so timeout 300 python tools/kernel_error_table.py 3 5 7; echo "== conserved-sum build"; timeout 300 python tools/kernel_error_table.py 3 5 7) > gpurun_out/r3j_kernel_error_table.log 2>&1; cat gpurun_out/r3j_kernel_error_table.log
for i in 1 2 3; do
  GNODE_B200_LIB=$PWD/tools/_ab/libgnode_b200_r3h.so timeout 300 python bench.py --trials 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3j_ab_prev_$i.json 2> gpurun_out/r3j_ab_prev_$i.err
  timeout 300 python bench.py --trials 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3j_ab_new_$i.json 2> gpurun_out/r3j_ab_new_$i.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3j_ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4e'%d['value'], '%.4f'%d['roofline']['frac'], 'e2e %.4e'%d['e2e']['value'], d['clocks']['sm_mhz'], d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
P
