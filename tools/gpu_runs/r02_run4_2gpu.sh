#!/bin/bash
# round 2, GPU run 4 (2 GPUs): peer-memory reduction inside the step graph -- correctness on 2 ranks, then the
# strong-scaling bench at N = 2 with this library's reduction and with NCCL
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_run4_topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/r02_run4_check.log 2>&1; echo "check rc=$?"
tail -12 gpurun_out/r02_run4_check.log
for mode in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --reduce $mode > gpurun_out/r02_run4_bench2_$mode.json 2> gpurun_out/r02_run4_bench2_$mode.err; echo "bench2 $mode rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_run4_bench2_$mode.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$mode", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), {k: round(v,4) for k,v in d["stage_ms"].items()}, d["config"]["parallelism"])
except Exception as e:
    print("$mode failed", e); print(open("gpurun_out/r02_run4_bench2_$mode.err").read()[-2500:])
PY
done
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r02_run4_bench1.json 2> gpurun_out/r02_run4_bench1.err; echo "bench1 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r02_run4_bench1.json').read().strip().splitlines()[-1]); print('N=1', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
