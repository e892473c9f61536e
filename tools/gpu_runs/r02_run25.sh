#!/bin/bash
# chain launch: all chains of a CTA interleaved in equal rounds (depth chosen on the device), depth sweep
mkdir -p gpurun_out
for d in 0 5 6 8; do
  ANI_B200_MLP_FUSED=1 ANI_B200_CHAIN_DEPTH=$d timeout 300 python bench.py --config water10k --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r25_chain_d${d}_water10k.json 2> gpurun_out/r25_chain_d${d}_water10k.err
done
for c in water1k protein50k gdb256; do
  ANI_B200_MLP_FUSED=1 timeout 300 python bench.py --config $c --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r25_chain_d0_${c}.json 2> gpurun_out/r25_chain_d0_${c}.err
done
for m in 2000 5000 6666; do
  ANI_B200_MLP_FUSED=1 timeout 300 python bench.py --config water10k --molecules $m --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r25_chain_d0_w${m}.json 2> gpurun_out/r25_chain_d0_w${m}.err
  ANI_B200_MLP_CHAIN=0 timeout 300 python bench.py --config water10k --molecules $m --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r25_nochain_w${m}.json 2> gpurun_out/r25_nochain_w${m}.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r25_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
ANI_B200_MLP_FUSED=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r25_pytest_forced.log 2>&1; echo "pytest (chain forced everywhere) rc=$?"; tail -2 gpurun_out/r25_pytest_forced.log
