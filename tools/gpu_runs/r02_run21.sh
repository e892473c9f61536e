#!/bin/bash
# register-direct GEMM epilogue (ANI_B200_EPI_DIRECT) A/B, parity, cluster preparation with 16 CTAs, timelines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r21_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r21_pytest.log
for ed in 0 1; do
  for c in water10k water1k protein50k; do
    ANI_B200_EPI_DIRECT=$ed timeout 300 python bench.py --config $c --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r21_ed${ed}_${c}.json 2> gpurun_out/r21_ed${ed}_${c}.err
  done
done
ANI_B200_MLP_FUSED=0 timeout 300 python bench.py --config water10k --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r21_chained_water10k.json 2> gpurun_out/r21_chained_water10k.err
ANI_B200_PREP_CLUSTER_CTAS=8 timeout 300 python bench.py --config water10k --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r21_pc8_water10k.json 2> gpurun_out/r21_pc8_water10k.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r21_*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 python tools/prep_trace.py 333 3333 5000 > gpurun_out/r21_prep_trace.log 2>&1; cat gpurun_out/r21_prep_trace.log | tail -4
ANI_B200_MLP_FUSED=1 timeout 300 python tools/gemm_trace.py > gpurun_out/r21_fused_trace.log 2>&1; head -27 gpurun_out/r21_fused_trace.log | cut -c1-250
for d in 0 8 6; do
  ANI_B200_MLP_FUSED=0 ANI_B200_GEMM_DEBUG=$d timeout 120 python tools/mlp_probe.py 2>&1 | tail -1
done | tee gpurun_out/r21_mlp_probe.log
