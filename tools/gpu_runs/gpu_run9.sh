set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export ANI_B200_GEMM_PAIR=1
( time timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r9_pytest_pair.log 2>&1
tail -8 gpurun_out/r9_pytest_pair.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv
timeout 200 python tests/gpu_diag.py > gpurun_out/r9_diag_pair.log 2>&1
grep -E "^===|MLP|forces:|status|Error|error" gpurun_out/r9_diag_pair.log | head -40
timeout 200 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r9_bench_pair.json 2> gpurun_out/r9_bench_pair.err
tail -3 gpurun_out/r9_bench_pair.err
unset ANI_B200_GEMM_PAIR
timeout 200 python bench.py --steps 50 --warmup 10 --cpu-steps 0 > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err
python - <<'PY'
import json
for f in ("r9_bench_pair","r9_bench"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,4) for k,v in d["stage_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
