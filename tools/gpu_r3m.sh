#!/bin/bash
# round 2, call 3M: packed pairs with the update's sums kept scalar (ptxas fuses mul.rn.f32x2 + add.rn.f32x2): bitwise check
# against the previous build in every mode / kernel, parity, A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
GNODE_B200_LIB=$PWD/tools/_ab/libgnode_b200_r3j.so python tools/_ab/dump_modes.py gpurun_out/r3m_prev.pt 2>&1 | tail -3
python tools/_ab/dump_modes.py gpurun_out/r3m_new.pt 2>&1 | tail -3
python tools/_ab/cmp.py gpurun_out/r3m_prev.pt gpurun_out/r3m_new.pt | tee gpurun_out/r3m_cmp.log
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_parity_gpu.py tests/test_trials_gpu.py tests/test_edge_cases_gpu.py tests/test_backward_gpu.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r3m_pytest.log; cat gpurun_out/r3m_pytest.log
for i in 1 2 3; do
  GNODE_B200_LIB=$PWD/tools/_ab/libgnode_b200_r3j.so timeout 300 python bench.py --trials 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3m_ab_prev_$i.json 2> gpurun_out/r3m_ab_prev_$i.err
  timeout 300 python bench.py --trials 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3m_ab_new_$i.json 2> gpurun_out/r3m_ab_new_$i.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3m_ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4e'%d['value'], '%.4f'%d['roofline']['frac'], 'e2e %.4e'%d['e2e']['value'], d['clocks']['sm_mhz'], d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
P
