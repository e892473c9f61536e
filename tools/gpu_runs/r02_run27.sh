#!/bin/bash
# data-flow launch: tile maps built side by side, stored-activation loads during the main loop (ydep barrier); parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r27_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r27_pytest.log
for c in water10k water1k gdb256 protein50k; do
  timeout 300 python bench.py --config $c --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r27_${c}.json 2> gpurun_out/r27_${c}.err
done
ANI_B200_MLP_FUSED=1 timeout 300 python bench.py --config protein50k --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r27_fused_protein50k.json 2> gpurun_out/r27_fused_protein50k.err
ANI_B200_MLP_FUSED=1 timeout 300 python bench.py --config water1k --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r27_fused_water1k.json 2> gpurun_out/r27_fused_water1k.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r27_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
