#!/bin/bash
# 8 GPUs after the round's kernel changes: correctness on 8 ranks, strong-scaling bench lines at N = 8 and 4
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r29_check.log 2>&1; echo "check rc=$?"
grep "ok\|Error\|error\|assert" gpurun_out/r29_check.log | tail -6
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r29_bench$n.json 2> gpurun_out/r29_bench$n.err; echo "bench$n rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r29_bench$n.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=$n", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms"].items()})
except Exception as e:
    print("N=$n failed", e); print(open("gpurun_out/r29_bench$n.err").read()[-2500:])
PY
done
