#!/bin/bash
# compact layer-1 backward operand (ANI_B200_B1_COMPACT) A/B + member grouping with it; preparation phase timeline;
# role timeline of the chained GEMMs with the new prefetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r18_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r18_pytest.log
for bc in 0 1; do
  for c in water10k water1k protein50k gdb256; do
    ANI_B200_B1_COMPACT=$bc timeout 300 python bench.py --config $c --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r18_bc${bc}_${c}.json 2> gpurun_out/r18_bc${bc}_${c}.err
  done
done
for g in 1 4; do
  ANI_B200_L1B_GROUP=$g timeout 300 python bench.py --config water10k --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r18_g${g}_water10k.json 2> gpurun_out/r18_g${g}_water10k.err
done
ANI_B200_MLP_FUSED=0 timeout 300 python bench.py --config water10k --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r18_chained_water10k.json 2> gpurun_out/r18_chained_water10k.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r18_*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 python tools/prep_trace.py 333 3333 5000 > gpurun_out/r18_prep_trace.log 2>&1; cat gpurun_out/r18_prep_trace.log | tail -4
ANI_B200_MLP_FUSED=0 timeout 300 python tools/gemm_trace.py > gpurun_out/r18_gemm_trace.log 2>&1; grep "==" gpurun_out/r18_gemm_trace.log | grep "cta 0"
