#!/bin/bash
# chain launch with early publication of completed units: parity, depth sweep, timeline
mkdir -p gpurun_out
ANI_B200_MLP_FUSED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r24_pytest_forced.log 2>&1; echo "pytest (chain forced everywhere) rc=$?"; tail -3 gpurun_out/r24_pytest_forced.log
for d in 2 3 4; do
  for c in water10k water1k; do
    ANI_B200_MLP_FUSED=1 ANI_B200_CHAIN_DEPTH=$d timeout 300 python bench.py --config $c --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r24_chain_d${d}_${c}.json 2> gpurun_out/r24_chain_d${d}_${c}.err
  done
done
ANI_B200_MLP_FUSED=1 ANI_B200_CHAIN_DEPTH=3 timeout 300 python bench.py --config protein50k --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r24_chain_d3_protein50k.json 2> gpurun_out/r24_chain_d3_protein50k.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r24_*_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
ANI_B200_MLP_FUSED=1 timeout 300 python tools/gemm_trace.py > gpurun_out/r24_chain_trace.log 2>&1; head -31 gpurun_out/r24_chain_trace.log | cut -c1-250
