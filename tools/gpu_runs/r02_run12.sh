#!/bin/bash
# round 2, GPU run 12: narrow column tiles for short lists, energy reduction beside the AEV backward, partial PBC
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rfEs --tb=short > gpurun_out/r02_run12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run12_pytest.log
grep -v "^  File\|^Extension" gpurun_out/r02_run12_pytest.log | tail -25
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --config ${CFG:-water10k} --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r02_run12_$name.json 2> gpurun_out/r02_run12_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_run12_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms"].items()})
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_run12_$name.err").read()[-1500:])
PY
}
run 10k X=1
run 10k_noverlap ANI_B200_OVERLAP_REDUCE=0
CFG=water1k run 1k X=1
CFG=water1k run 1k_wide ANI_B200_NARROW_TILES=0
CFG=water1k run 1k_fused ANI_B200_MLP_FUSED=1
CFG=gdb256 run gdb X=1
CFG=protein50k run 50k X=1
