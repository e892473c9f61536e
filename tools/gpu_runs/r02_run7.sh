#!/bin/bash
# round 2, GPU run 7: lagged completion fix; ncu of the fused MLP kernel; fallback test
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --config ${CFG:-water10k} --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r02_run7_$name.json 2> gpurun_out/r02_run7_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_run7_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms"].items()}, "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_run7_$name.err").read()[-1500:])
PY
}
run w8 X=1
run unfused ANI_B200_MLP_FUSED=0
CFG=water1k run 1k_w8 X=1
CFG=protein50k run 50k_w8 X=1
CFG=protein50k run 50k_unfused ANI_B200_MLP_FUSED=0
timeout 300 python -m pytest tests/test_gpu_api.py -q -k "fallback" --tb=short 2>&1 | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mlp_fused -c 1 -o gpurun_out/r02_ncu_mlp_fused -f python bench.py --steps 2 --warmup 1 --cpu-steps 0 > gpurun_out/r02_run7_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r02_run7_ncu.log
