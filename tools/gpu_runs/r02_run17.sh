#!/bin/bash
# cluster preparation kernel (k_prep_cluster) + stored-activation prefetch a whole group ahead: parity, A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r17_pytest_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r17_pytest_parity.log
for pc in 0 1; do
  for c in water10k water1k; do
    ANI_B200_PREP_CLUSTER=$pc timeout 300 python bench.py --config $c --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r17_pc${pc}_${c}.json 2> gpurun_out/r17_pc${pc}_${c}.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r17_pc*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r17_pytest_api.log 2>&1; echo "api rc=$?"; tail -3 gpurun_out/r17_pytest_api.log
