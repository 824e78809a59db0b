#!/bin/bash
# round 2, call 3K: row sums, SIR update, accumulator sums and sigmoid on packed fp32 pairs (FADD2 / FMUL2 / FFMA2): bitwise
# parity (variants: pair form vs scalar form), then A/B against the previous build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_parity_gpu.py tests/test_trials_gpu.py tests/test_edge_cases_gpu.py tests/test_backward_gpu.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r3k_pytest.log; cat gpurun_out/r3k_pytest.log
for i in 1 2 3; do
  GNODE_B200_LIB=$PWD/tools/_ab/libgnode_b200_r3j.so timeout 300 python bench.py --trials 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3k_ab_prev_$i.json 2> gpurun_out/r3k_ab_prev_$i.err
  timeout 300 python bench.py --trials 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3k_ab_new_$i.json 2> gpurun_out/r3k_ab_new_$i.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3k_ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4e'%d['value'], '%.4f'%d['roofline']['frac'], 'e2e %.4e'%d['e2e']['value'], d['clocks']['sm_mhz'], d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
P
