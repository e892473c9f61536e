#!/bin/bash
# A/B of the layer-1 backward member grouping (ANI_B200_L1B_GROUP) + parity of the default
mkdir -p gpurun_out
for g in 1 2 4; do
  for c in water10k water1k protein50k; do
    ANI_B200_L1B_GROUP=$g timeout 300 python bench.py --config $c --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r16_g${g}_${c}.json 2> gpurun_out/r16_g${g}_${c}.err
  done
done
for g in 1 2 4; do ANI_B200_L1B_GROUP=$g timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r16_pytest_g$g.log 2>&1; tail -1 gpurun_out/r16_pytest_g$g.log; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r16_g*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d.get('stages_us'))
    except Exception as e: print(f, 'ERR', e)
PY
