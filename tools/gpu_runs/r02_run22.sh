#!/bin/bash
# 32-aligned angular block (ANI_B200_ALIGN_ANGULAR), row-tile windows of the data-flow launch (ANI_B200_MLP_CHUNKS),
# register-direct epilogue in the chained launches only; parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r22_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r22_pytest.log
for al in 0 1; do
  for ch in 1 2 3 4; do
    ANI_B200_ALIGN_ANGULAR=$al ANI_B200_MLP_CHUNKS=$ch timeout 300 python bench.py --config water10k --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r22_al${al}_ch${ch}_water10k.json 2> gpurun_out/r22_al${al}_ch${ch}_water10k.err
  done
done
for ch in 1 4 8 16; do
  ANI_B200_MLP_CHUNKS=$ch timeout 300 python bench.py --config protein50k --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r22_al1_ch${ch}_protein50k.json 2> gpurun_out/r22_al1_ch${ch}_protein50k.err
done
ANI_B200_ALIGN_ANGULAR=0 ANI_B200_MLP_CHUNKS=1 timeout 300 python bench.py --config protein50k --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r22_al0_ch1_protein50k.json 2> gpurun_out/r22_al0_ch1_protein50k.err
for c in water1k gdb256; do
  timeout 300 python bench.py --config $c --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r22_al1_auto_${c}.json 2> gpurun_out/r22_al1_auto_${c}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r22_*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
