#!/bin/bash
# round 2, GPU run 8: producer look-ahead, grace-period completion; full suite
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --config ${CFG:-water10k} --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r02_run8_$name.json 2> gpurun_out/r02_run8_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_run8_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms"].items()}, "frac", round(d["roofline"]["frac"],3))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_run8_$name.err").read()[-1500:])
PY
}
run w8 X=1
run unfused ANI_B200_MLP_FUSED=0
CFG=water1k run 1k_w8 X=1
CFG=protein50k run 50k_w8 X=1
CFG=gdb256 run gdb_w8 X=1
CFG=gdb256 run gdb_unfused ANI_B200_MLP_FUSED=0
timeout 1200 python -m pytest tests -m gpu -q -rfEs --tb=short > gpurun_out/r02_run8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run8_pytest.log
grep -v "^  File\|^Extension" gpurun_out/r02_run8_pytest.log | tail -30
