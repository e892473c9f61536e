#!/bin/bash
# round 2, GPU run 1: full GPU suite, the reference's GPU path beside ours, every bench config
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02_run1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run1_pytest.log
tail -5 gpurun_out/r02_run1_pytest.log
timeout 600 python tools/reference_gpu_path.py > gpurun_out/r02_run1_refpath.log 2>&1; echo "refpath rc=$?"
tail -30 gpurun_out/r02_run1_refpath.log
for cfg in water10k water1k gdb256 protein50k; do
  timeout 400 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/r02_run1_bench_$cfg.json 2> gpurun_out/r02_run1_bench_$cfg.err; echo "bench $cfg rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_run1_bench_$cfg.json").read().strip().splitlines()[-1])
    print("$cfg", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["stage_ms"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["ms_per_step"], "frac", d["roofline"]["frac"])
except Exception as e:
    print("$cfg failed", e); print(open("gpurun_out/r02_run1_bench_$cfg.err").read()[-1500:])
PY
done
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_run1_bench_reference.json 2> gpurun_out/r02_run1_bench_reference.err; echo "bench reference rc=$?"
tail -c 600 gpurun_out/r02_run1_bench_reference.json
