#!/bin/bash
# round 2, GPU run 15: AEV forward staging from the per-bucket species offsets (no counting pass): suite + timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rfEs --tb=short > gpurun_out/r02_run15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run15_pytest.log
grep -v "^  File\|^Extension" gpurun_out/r02_run15_pytest.log | tail -25
for cfg in water10k water1k protein50k gdb256; do
  timeout 300 python bench.py --config $cfg --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r02_run15_$cfg.json 2> gpurun_out/r02_run15_$cfg.err; echo "$cfg rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_run15_$cfg.json").read().strip().splitlines()[-1])
    print("$cfg", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms"].items()})
except Exception as e:
    print("$cfg failed", e); print(open("gpurun_out/r02_run15_$cfg.err").read()[-1500:])
PY
done
