#!/bin/bash
# 2 GPUs after the round's kernel changes (cluster preparation on an owned slice, aligned angular block, data-flow MLP):
# correctness of the sharded step + peer-memory reduction, then the N = 2 bench line
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r28_check.log 2>&1; echo "check rc=$?"
grep "ok\|Error\|error\|assert" gpurun_out/r28_check.log | tail -12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r28_bench2.json 2> gpurun_out/r28_bench2.err; echo "bench2 rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r28_bench2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=2", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms"].items()}, d["config"]["parallelism"])
except Exception as e:
    print("N=2 failed", e); print(open("gpurun_out/r28_bench2.err").read()[-2500:])
PY
