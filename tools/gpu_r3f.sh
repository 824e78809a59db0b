#!/bin/bash
# round 2, call 3F: fused step 0 with the neighbours' seed bits looked up together (staged colidx slice -> flags): parity, A/B, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_trials_gpu.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r3f_pytest.log; cat gpurun_out/r3f_pytest.log
for i in 1 2; do
  for m in fill fused; do
    GNODE_TRIALS_ENCODE=$m timeout 300 python bench.py --trials 1024 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r3f_ab_${m}_$i.json 2> gpurun_out/r3f_ab_${m}_$i.err
  done
done
python - <<'P'
import json,glob
for f in sorted(glob.glob('gpurun_out/r3f_ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4e'%d['value'], '%.4f'%d['roofline']['frac'], 'e2e %.4e'%d['e2e']['value'], d['clocks']['sm_mhz'], d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3f_launches.csv \
    python bench.py --steps 1 --warmup 1 --trials 256 --no-cpu-baseline > gpurun_out/r3f_ncu_launch.log 2>&1; echo "launch-list rc=$?"
grep "193>" gpurun_out/r3f_launches.csv | head -3
