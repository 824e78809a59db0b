#!/bin/bash
# round 2, call 3B: descriptor-fed encoder (two-row table + store stream): parity, then same-box A/B on the bench job
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_trials_gpu.py tests/test_parity_gpu.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r3b_pytest.log; cat gpurun_out/r3b_pytest.log
for i in 1 2; do
  GNODE_TRIALS_ENCODE=dense timeout 300 python bench.py --trials 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3b_ab_dense_$i.json 2> gpurun_out/r3b_ab_dense_$i.err
  timeout 300 python bench.py --trials 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3b_ab_table_$i.json 2> gpurun_out/r3b_ab_table_$i.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3b_ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4e'%d['value'], '%.4f'%d['roofline']['frac'], 'e2e %.4e'%d['e2e']['value'], d['clocks']['sm_mhz'], d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
P
