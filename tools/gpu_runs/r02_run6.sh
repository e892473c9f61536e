#!/bin/bash
# round 2, GPU run 6: A/B of the fused MLP variants (8 / 16 epilogue warps, weight prefetch on / off)
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --config ${CFG:-water10k} --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r02_run6_$name.json 2> gpurun_out/r02_run6_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_run6_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms"].items()}, "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_run6_$name.err").read()[-1500:])
PY
}
run w8_pre X=1
run w8_nopre ANI_B200_PREFETCH_B=0
run w16_pre ANI_B200_EPI_WARPS=16
run w16_nopre ANI_B200_EPI_WARPS=16 ANI_B200_PREFETCH_B=0
CFG=water1k run 1k_w8_pre X=1
CFG=water1k run 1k_w8_nopre ANI_B200_PREFETCH_B=0
CFG=water1k run 1k_w16_pre ANI_B200_EPI_WARPS=16
CFG=protein50k run 50k_w8_pre X=1
CFG=protein50k run 50k_w8_nopre ANI_B200_PREFETCH_B=0
timeout 300 python -m pytest tests/test_gpu_api.py -q -k "fallback" --tb=long 2>&1 | tail -30
