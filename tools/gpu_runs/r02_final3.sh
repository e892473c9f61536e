#!/bin/bash
# last run of the round (no library change): the opt-in test that times the pieces of the reference's GPU path beside
# ours, and a size sweep of the final build (data-flow vs chained MLP launches)
mkdir -p gpurun_out
ANI_B200_REF_PATH_TEST=1 timeout 600 python -m pytest tests/test_gpu_reference_cuaev.py -m gpu -q -s > gpurun_out/r02_final_reference_pieces_test.log 2>&1; echo "ref pieces rc=$?"; tail -5 gpurun_out/r02_final_reference_pieces_test.log | cut -c1-300
for m in 700 1500 2500 5000; do
  for f in 0 1; do
    ANI_B200_MLP_FUSED=$f timeout 200 python bench.py --config water10k --molecules $m --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r02_final_sweep_m${m}_f${f}.json 2> gpurun_out/r02_final_sweep_m${m}_f${f}.err
  done
done
python - <<'PY'
import json,glob
print("| atoms | launch | ms / step | e2e ms | MLP stage ms |"); print("|---|---|---|---|---|")
for m in (700,1500,2500,5000):
    for f in (0,1):
        try:
            d=json.loads(open(f'gpurun_out/r02_final_sweep_m{m}_f{f}.json').read().strip().splitlines()[-1])
            print(f"| {3*m} | {'data-flow (1 launch)' if f else 'chained (6 launches)'} | {d['ms_per_step']:.4f} | {d['e2e']['ms_per_step']:.4f} | {d['stage_ms']['mlp_forward_backward']:.4f} |")
        except Exception as e: print(m,f,'ERR',e)
PY
