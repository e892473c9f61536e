#!/bin/bash
# sixteen-warp register-direct epilogue of the data-flow launch (ANI_B200_EPI_WARPS=16): parity, A/B, timeline
mkdir -p gpurun_out
ANI_B200_EPI_WARPS=16 ANI_B200_MLP_FUSED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r30_pytest16.log 2>&1; echo "pytest (16 warps, data-flow forced) rc=$?"; tail -3 gpurun_out/r30_pytest16.log
for w in 8 16; do
  for c in water10k; do
    ANI_B200_EPI_WARPS=$w ANI_B200_MLP_FUSED=1 timeout 300 python bench.py --config $c --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r30_w${w}_${c}.json 2> gpurun_out/r30_w${w}_${c}.err
  done
done
ANI_B200_EPI_WARPS=16 ANI_B200_EPI16_PAIRS=1 ANI_B200_MLP_FUSED=1 timeout 300 python bench.py --config water10k --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r30_w16pairs_water10k.json 2> gpurun_out/r30_w16pairs_water10k.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r30_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
ANI_B200_EPI_WARPS=16 ANI_B200_MLP_FUSED=1 timeout 300 python bench.py --config protein50k --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r30_w16_protein50k.json 2> gpurun_out/r30_w16_protein50k.err
python -c "
import json; d=json.loads(open('gpurun_out/r30_w16_protein50k.json').read().strip().splitlines()[-1]); print('w16 protein50k', d['ms_per_step'], d['stage_ms']['mlp_forward_backward'])"
