#!/bin/bash
# cluster preparation at 10k atoms (dynamic bin_start copy, parallel live blocks), trash-bucket order by slot (gdb256),
# role timeline of the data-flow MLP launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r19_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r19_pytest.log
for pc in 0 1; do
  for c in water10k water1k gdb256; do
    ANI_B200_PREP_CLUSTER=$pc timeout 300 python bench.py --config $c --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r19_pc${pc}_${c}.json 2> gpurun_out/r19_pc${pc}_${c}.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r19_*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 python tools/prep_trace.py 333 3333 5000 > gpurun_out/r19_prep_trace.log 2>&1; cat gpurun_out/r19_prep_trace.log | tail -4
ANI_B200_MLP_FUSED=1 timeout 300 python tools/gemm_trace.py > gpurun_out/r19_fused_trace.log 2>&1; head -40 gpurun_out/r19_fused_trace.log
