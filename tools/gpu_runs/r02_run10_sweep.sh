#!/bin/bash
# size sweep: data-flow vs chained MLP launches (water boxes)
mkdir -p gpurun_out
for mol in 700 1000 1500 2000 5000 6666 10000; do
  for mode in 0 1; do
    ANI_B200_MLP_FUSED=$mode timeout 300 python bench.py --molecules $mol --steps 15 --warmup 5 --cpu-steps 0 > gpurun_out/r02_sweep_${mol}_${mode}.json 2> gpurun_out/r02_sweep_${mol}_${mode}.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_sweep_${mol}_${mode}.json").read().strip().splitlines()[-1])
    print("atoms", 3*$mol, "fused" if "$mode"=="1" else "chained", round(d["ms_per_step"],4), "mlp", round(d["stage_ms"]["mlp_forward_backward"],4))
except Exception as e:
    print("$mol $mode failed", e)
PY
  done
done
